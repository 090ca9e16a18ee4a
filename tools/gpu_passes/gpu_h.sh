#!/bin/bash
mkdir -p gpurun_out
for k in 20 200 2000; do
timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_h_k$k.json 2> gpurun_out/bench_h_k$k.err
done
SMGX_EVENT_PATH=fused timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_h_simple_k20.json 2> gpurun_out/bench_h_simple_k20.err
for f in gpurun_out/bench_h_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'lat', round(d['latency']['device_resident_p50_us'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
( time timeout 1200 python -m pytest tests/test_gpu_feedback.py tests/test_cpp_mirror.py tests/test_gpu_scale.py -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu_h.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_h.log | tail -8
grep -B5 "Error" gpurun_out/pytest_gpu_h.log | head -60
