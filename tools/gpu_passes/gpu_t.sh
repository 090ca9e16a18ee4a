#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift; shift; env "$@" timeout 120 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_t_$name.json 2> gpurun_out/bench_t_$name.err; }
run conc_k20 20 X=1
run noconc_k20 20 SMGX_SPLIT_CONCURRENT=0
run conc_k10 10 X=1
run conc_k32 32 X=1
run conc_k200 200 X=1
run conc_k2000 2000 X=1
for f in gpurun_out/bench_t_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], 'launches', d['gpu_launches'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
( time timeout 600 python -m pytest tests/test_gpu_scale.py tests/test_gpu_event_select.py -m gpu -q --maxfail=15 -k "split or select" ) > gpurun_out/pytest_gpu_t.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_t.log | tail -8
