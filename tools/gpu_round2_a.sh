#!/bin/bash
# GPU call A (round 2): parity tests, bench line, launch list, K1 --set full with source counters.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -25 > gpurun_out/pytest_gpu_r2a.log; tail -5 gpurun_out/pytest_gpu_r2a.log
timeout 300 python bench.py --steps 30 --warmup 5 > gpurun_out/bench_r2a.json 2> gpurun_out/bench_r2a.err; echo "bench rc=$?"
tail -c 600 gpurun_out/bench_r2a.json
bash tools/gpu_ncu_kernel.sh stft_mel r2a_k1 4
