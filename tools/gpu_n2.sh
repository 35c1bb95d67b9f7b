#!/bin/bash
# 2-GPU check: default bench line at N=2 (as the driver launches it) and the reference arm.
tag=${1:-n2}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_${tag}_n2.json 2> gpurun_out/bench_${tag}_n2.err; echo "n2 rc=$?"
tail -5 gpurun_out/bench_${tag}_n2.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}_n2.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', d['e2e'])
    print('   notes', d.get('notes'))
    print('   sustained', d.get('sustained'))
    print('   cfg5', d.get('cfg5'))
except Exception as e:
    print('no json', e)
PY
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo "ref rc=$?"; tail -c 1200 gpurun_out/bench_${tag}_ref.json
