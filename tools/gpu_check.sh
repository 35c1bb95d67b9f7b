mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 300 2>&1 | tail -12 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench2.json')); print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['kernel_ms'], d['roofline']['frac'])"; tail -3 gpurun_out/bench2.err
python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/plain.log 2>&1 && ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"mfcc_kernel|sosfilt|delta_norm|modspec|stft_mel|fill_i32" -s 21 -c 7 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/ncu1.log 2>&1; echo rc=$?
