#!/bin/bash
# Usage (on the GPU box): bash tools/gpu_launches.sh <tag>  -- ncu launch list of one bench step
tag=${1:-run}
mkdir -p gpurun_out
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain_$tag.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
  -k regex:"mfcc|pcm16|sosfilt|delta_norm|modspec|stft_mel|fill_i32|change_fused" -s 20 -c 8 --csv \
  --log-file gpurun_out/launches_$tag.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu_$tag.log 2>&1
echo rc=$?
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches_$tag.csv')) if len(r)>5]
h=rows[0]
for r in rows[1:]:
    if r[h.index('Metric Name')]=='gpu__time_duration.sum':
        print(r[h.index('Kernel Name')][:70], r[h.index('Metric Value')], r[h.index('Grid Size')], r[h.index('Block Size')])
PY
