#!/bin/bash
# One `ncu --set full` capture of every kernel of one bench step (after a plain run has exited 0).
tag=${1:-step}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu --no-sustained --no-cfg5 > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"mfcc_pk|mfcc_kernel|change_fused|sosfiltfilt_par|modspec_tc|stft_mel" -s 15 -c 5 -f -o gpurun_out/prof_$tag \
  python bench.py --steps 3 --warmup 3 --no-cpu --no-sustained --no-cfg5 > gpurun_out/ncu_$tag.log 2>&1
echo rc=$?
ls -la gpurun_out/prof_$tag.ncu-rep
