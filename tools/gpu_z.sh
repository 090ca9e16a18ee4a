#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_event_select.py -q -m gpu -x 2>&1 | tail -2
timeout 600 python -m pytest tests/test_gpu_scale.py -q -m gpu -x -k "split or duplicate" 2>&1 | tail -2
run() { name=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_z_$name.json 2> gpurun_out/bench_z_$name.err; }
run pdl_k20 20 X=1
run nopdl_k20 20 SMGX_PDL=0
run pdl_k2000 2000 X=1
run pdl_k200 200 X=1
run pdl_k5 5 X=1
run pdl_k1 1 X=1
for f in gpurun_out/bench_z_pdl_*.json gpurun_out/bench_z_nopdl*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], d['gpu_launches'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
