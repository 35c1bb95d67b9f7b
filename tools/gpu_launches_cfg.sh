#!/bin/bash
# Usage (on the GPU box): bash tools/gpu_launches_cfg.sh <cfg3|gui|cfg4> <tag>  -- ncu launch list of one pass of another config
cfg=${1:-cfg3}; tag=${2:-$cfg}
mkdir -p gpurun_out
python tools/bench_cfgs.py $cfg > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"stft_mel|mfcc|change_fused|sosfilt|modspec|fill_i32|mel_empty|delta_norm" -s 12 -c 7 --csv --log-file gpurun_out/launches_$tag.csv python tools/bench_cfgs.py $cfg > gpurun_out/ncu_$tag.log 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$tag.csv')) if len(r)>5]
h=rows[0]
for r in rows[1:]:
    if r[h.index('Metric Name')]=='gpu__time_duration.sum':
        print(r[h.index('Kernel Name')][:80], r[h.index('Metric Value')], r[h.index('Grid Size')], r[h.index('Block Size')])
PY
tail -2 gpurun_out/plain_$tag.log | cut -c1-100
