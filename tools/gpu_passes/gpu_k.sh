#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; env "$@" timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_k_$name.json 2> gpurun_out/bench_k_$name.err; }
run split_b2b BENCH_WARM_BACK_TO_BACK=1
run simple_b2b BENCH_WARM_BACK_TO_BACK=1 SMGX_EVENT_PATH=fused
run tile16_b2b BENCH_WARM_BACK_TO_BACK=1 SMGX_EVENT_PATH=fused SMGX_EVENT_SIMPLE=0
for f in gpurun_out/bench_k_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-500:])
PY
done
