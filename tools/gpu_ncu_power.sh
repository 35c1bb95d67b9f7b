#!/bin/bash
# Usage (on the GPU box): bash tools/gpu_ncu_power.sh <flags> <tag>
# One `ncu --set full` capture of the mmf_stft_power kernel for a plan flag set (after a plain run exited 0).
fl=$1; tag=$2
mkdir -p gpurun_out
python tools/bench_power.py --flags $fl --iters 3 > gpurun_out/plain_$tag.log 2>&1 || exit 1
cat gpurun_out/plain_$tag.log
ncu --set full --clock-control none --import-source on -k regex:"tc_fft512|stft_mel" -s 2 -c 1 -f -o gpurun_out/prof_$tag \
  python tools/bench_power.py --flags $fl --iters 3 > gpurun_out/ncu_$tag.log 2>&1
echo rc=$?
