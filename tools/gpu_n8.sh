#!/bin/bash
# N-GPU check (default 8): the bench line as the driver launches it.
N=${1:-8}; tag=${2:-n$N}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_${tag}.json 2> gpurun_out/bench_${tag}.err; echo "n$N rc=$?"
tail -4 gpurun_out/bench_${tag}.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches','n_gpus')}, 'e2e', d['e2e']['value'], d['e2e'].get('h2d_ceiling_gbs_all_ranks'))
    print('   notes', d.get('notes'))
    print('   sustained', d.get('sustained'))
    print('   cfg5', d.get('cfg5'))
except Exception as e:
    print('no json', e)
PY
