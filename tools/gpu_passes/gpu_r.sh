#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_r_$name.json 2> gpurun_out/bench_r_$name.err; }
run t16d4_k20 20 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=16 SMGX_TILE_DEPTH=4
run t16d8_k20 20 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=16 SMGX_TILE_DEPTH=8
run t8d4_k20 20 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=8 SMGX_TILE_DEPTH=4
run t8d8_k20 20 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=8 SMGX_TILE_DEPTH=8
run t32d4_k20 20 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=32 SMGX_TILE_DEPTH=4
run t16d4_k2000 2000 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=16 SMGX_TILE_DEPTH=4
run t8d8_k2000 2000 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 SMGX_FUSED_TILE=8 SMGX_TILE_DEPTH=8
for f in gpurun_out/bench_r_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], 'launches', d['gpu_launches'], 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tile -s 4 -c 1 -o gpurun_out/tile_r02r -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
( time timeout 1200 python -m pytest tests/test_gpu_scale.py -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu_r.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu_r.log | tail -8
