#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_s_$name.json 2> gpurun_out/bench_s_$name.err; }
run rpc128_k20 20 SMGX_SEARCH_RPC=128
run rpc256_k20 20 SMGX_SEARCH_RPC=256
run rpc128_k2000 2000 SMGX_SEARCH_RPC=128
run rpc256_k2000 2000 SMGX_SEARCH_RPC=256
for f in gpurun_out/bench_s_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], 'launches', d['gpu_launches'], 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
