#!/bin/bash
mkdir -p gpurun_out
for k in 1 5 10 20 32; do
timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_n_split_k$k.json 2> gpurun_out/bench_n_split_k$k.err
SMGX_EVENT_PATH=fused timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_n_simple_k$k.json 2> gpurun_out/bench_n_simple_k$k.err
done
for f in gpurun_out/bench_n_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], 'ungated', [round(x*1e3,1) for x in d['region_ms_ungated']], 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-300:])
PY
done
