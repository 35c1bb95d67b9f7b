#!/bin/bash
# Usage: bash tools/gpu_round2_multi.sh <tag> <N> [extra bench args]   -- bench at N=1 and at N GPUs
tag=$1; N=$2; shift 2
mkdir -p gpurun_out
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu "$@" > gpurun_out/bench_${tag}_n1.json 2> gpurun_out/bench_${tag}_n1.err; echo "n1 rc=$?"
tail -3 gpurun_out/bench_${tag}_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu "$@" > gpurun_out/bench_${tag}_n$N.json 2> gpurun_out/bench_${tag}_n$N.err; echo "n$N rc=$?"
tail -5 gpurun_out/bench_${tag}_n$N.err
python - <<PY
import json
for n in (1, $N):
    try:
        d=json.load(open('gpurun_out/bench_${tag}_n%d.json'%n))
    except Exception as e:
        print(n, 'no json', e); continue
    print(n, {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('verified_bit_identical_to_device_path'), 'h2d', d['e2e'].get('h2d_ceiling_gbs_all_ranks'))
    print('   notes', d.get('notes'))
    print('   sustained', d.get('sustained'))
    print('   cfg5', d.get('cfg5'))
PY
