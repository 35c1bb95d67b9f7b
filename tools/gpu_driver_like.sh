#!/bin/bash
# What the driver does at round end, in one call: GPU tests, smoke(), default bench, reference arm.
tag=${1:-final}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/pytest_gpu_$tag.log; cat gpurun_out/pytest_gpu_$tag.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err; echo "bench rc=$?"; tail -2 gpurun_out/bench_$tag.err
timeout 900 python bench.py --impl reference > gpurun_out/bench_${tag}_ref.json 2> gpurun_out/bench_${tag}_ref.err; echo "ref rc=$?"
python - <<PY
import json
d=json.load(open('gpurun_out/bench_$tag.json')); r=json.load(open('gpurun_out/bench_${tag}_ref.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','steps','warmup')}, 'e2e', d['e2e']['value'], d['e2e'].get('verified_bit_identical_to_device_path'))
print('roofline', {k:d['roofline'][k] for k in ('kernel','achieved','frac','kernel_ms','traffic')})
print('tensor', d['roofline'].get('tensor_pipe'))
print('k6', d['roofline_k6']['frac'], 'step', d['roofline_step']['frac'], 'clocks', d['clocks'])
print('cpu', d['cpu_baseline'])
print('ref', r['value'], r['cpu_baseline']['cores'], 'same config', r['config']==d['config'])
print('ratio e2e', d['e2e']['value']/r['e2e']['value'], 'ratio value', d['value']/r['value'])
PY
