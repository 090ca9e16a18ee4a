#!/bin/bash
# round-2 GPU pass E: simple event kernel (one warp per request, registers only) + slow queue
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_e_default.json 2> gpurun_out/bench_e_default.err
SMGX_EVENT_SIMPLE=0 timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_e_tile16.json 2> gpurun_out/bench_e_tile16.err
for k in 200 2000; do
timeout 300 python bench.py --steps $k --warmup 10 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_e_k$k.json 2> gpurun_out/bench_e_k$k.err
done
timeout 300 python bench.py --steps 2000 --warmup 10 --lanes 1 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_e_k2000_l1.json 2> gpurun_out/bench_e_k2000_l1.err
SMGX_EVENT_SIMPLE=0 timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_e_k2000_tile16.json 2> gpurun_out/bench_e_k2000_tile16.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_simple -s 4 -c 1 -o gpurun_out/simple_r02e -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_e.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_e_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'e2e %.3g'%d['e2e']['value'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
grep -E "event_simple|event_slow" gpurun_out/launches_e.csv | awk -F'","' '{print $5, $NF}' | sed 's/"//g' | sort | uniq -c | sort -rn | head -8
