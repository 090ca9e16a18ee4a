#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_x_$name.json 2> gpurun_out/bench_x_$name.err; }
run k20 20 X=1
run k200 200 X=1
run k2000 2000 X=1
for f in gpurun_out/bench_x_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], 'launches', d['gpu_launches'], 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hash_blocks|event_search2" -s 14 -c 2 -o gpurun_out/split_r02x -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_x.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
