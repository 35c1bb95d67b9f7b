#!/bin/bash
# Usage (on the GPU box): bash tools/gpu_ncu_kernel.sh <kernel-regex> <tag> [launch-skip]
# One `ncu --set full` capture of a kernel of the bench step (after a plain run has exited 0).
k=$1; tag=$2; skip=${3:-4}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"$k" -s $skip -c 1 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
echo rc=$?
ls -la gpurun_out/prof_$tag.ncu-rep
