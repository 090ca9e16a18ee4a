#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_m_$name.json 2> gpurun_out/bench_m_$name.err; }
run split_k20 20 X=1
run split_k200 200 X=1
run split_k2000 2000 X=1
for f in gpurun_out/bench_m_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'launches', d['gpu_launches'], 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_m.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
grep -E "event_search2|hash_blocks" gpurun_out/launches_m.csv | grep -E "\(1?[0-9]+, 20, 1\)" | awk -F'","' '{print $5, $(NF-4), $NF}' | sed 's/"//g' | head -8
( time timeout 1200 python -m pytest tests/test_gpu_event_select.py tests/test_gpu_scale.py tests/test_gpu_sharded.py tests/test_gpu_config3.py -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu_m.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_m.log | tail -8
