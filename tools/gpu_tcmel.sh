#!/bin/bash
# GPU call: the tcgen05 mel kernel -- its own tests first (short timeout: a hang must not eat the budget), then the
# bench with and without it, the whole GPU suite, and (arg 2 = ncu) a --set full capture of the kernel.
tag=${1:-tcmel}
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 120 -k "tcgen05_mel" 2>&1 | tail -25 > gpurun_out/pytest_tcmel_$tag.log
tail -5 gpurun_out/pytest_tcmel_$tag.log
if ! grep -q " passed" gpurun_out/pytest_tcmel_$tag.log || grep -q "failed" gpurun_out/pytest_tcmel_$tag.log; then echo "TC MEL TESTS NOT GREEN"; exit 1; fi
for f in 0 4096; do
  MMF_FLAGS=$f timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench_${tag}_f$f.json 2> gpurun_out/bench_${tag}_f$f.err; echo "bench flags=$f rc=$?"
  python - <<PY
import json
try:
    d=json.load(open('gpurun_out/bench_${tag}_f$f.json'))
    print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, 'e2e', d['e2e']['value'], d['e2e'].get('verified_bit_identical_to_device_path'), 'k1_ms', d['roofline']['kernel_ms'], 'frac', d['roofline']['frac'])
except Exception as e:
    print('no json', e)
PY
  tail -3 gpurun_out/bench_${tag}_f$f.err
done
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -12 > gpurun_out/pytest_gpu_$tag.log; tail -4 gpurun_out/pytest_gpu_$tag.log
if [ "$2" = "ncu" ]; then bash tools/gpu_ncu_kernel.sh stft_mel ${tag} 4; fi
