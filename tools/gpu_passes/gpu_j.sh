#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_j_$name.json 2> gpurun_out/bench_j_$name.err; }
run split X=1
run simple SMGX_EVENT_PATH=fused
run tile16 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0
run fused4 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=0
for f in gpurun_out/bench_j_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'lat', round(d['latency']['device_resident_p50_us'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
