#!/bin/bash
# GPU call: parity tests, bench with the grouped mel walk (default) and the old sparse walk, K1 profile.
tag=${1:-k1}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_gpu_$tag.log; tail -3 gpurun_out/pytest_gpu_$tag.log
for f in 0 2048; do
  MMF_FLAGS=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_${tag}_f$f.json 2> gpurun_out/bench_${tag}_f$f.err; echo "bench flags=$f rc=$?"
  python - <<PY
import json
d=json.load(open('gpurun_out/bench_${tag}_f$f.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], 'k1_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
PY
done
if [ "$2" = "ncu" ]; then bash tools/gpu_ncu_kernel.sh stft_mel ${tag} 4; fi
