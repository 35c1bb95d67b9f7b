#!/bin/bash
# Usage (on the GPU box, via gpurun): bash tools/gpu_check.sh [tag]
# GPU parity tests, then a bench line, written under gpurun_out/.
tag=${1:-run}
mkdir -p gpurun_out
timeout 420 python -m pytest tests -m gpu -q -x --timeout 90 2>&1 | tail -15 > gpurun_out/pytest_gpu_$tag.log; tail -4 gpurun_out/pytest_gpu_$tag.log
timeout 180 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'k1_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
PY
tail -3 gpurun_out/bench_$tag.err
