#!/bin/bash
# Usage (on the GPU box): bash tools/gpu_ncu_cfg.sh <cfg3|gui> <tag>  -- ncu --set full of K1 of another config
cfg=${1:-cfg3}; tag=${2:-$cfg}
mkdir -p gpurun_out
python tools/bench_cfgs.py $cfg > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:"stft_mel" -s 3 -c 1 -f -o gpurun_out/prof_$tag python tools/bench_cfgs.py $cfg > gpurun_out/ncu_$tag.log 2>&1
echo rc=$?
