#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_i_$name.json 2> gpurun_out/bench_i_$name.err; }
run split_pipe2 SMGX_SPLIT_PIPE=2
run split_pipe3 SMGX_SPLIT_PIPE=3
run split_pipe4 SMGX_SPLIT_PIPE=4
run simple_pf1 SMGX_EVENT_PATH=fused SMGX_SIMPLE_PF=1
run simple_pf2 SMGX_EVENT_PATH=fused SMGX_SIMPLE_PF=2
run simple_pf3 SMGX_EVENT_PATH=fused SMGX_SIMPLE_PF=3
run simple_pf0 SMGX_EVENT_PATH=fused
SMGX_EVENT_PATH=fused SMGX_SIMPLE_PF=2 timeout 300 python bench.py --steps 2000 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_i_simple_pf2_k2000.json 2> /dev/null
for f in gpurun_out/bench_i_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'lat', round(d['latency']['device_resident_p50_us'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
