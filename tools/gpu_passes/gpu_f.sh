#!/bin/bash
# round-2 GPU pass F: split path with the balanced search kernel (event_search2_kernel)
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for k in 20 200 2000; do
SMGX_EVENT_PATH=split timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_f_split_k$k.json 2> gpurun_out/bench_f_split_k$k.err
done
SMGX_EVENT_PATH=split SMGX_SEARCH_V1=1 timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_f_splitv1_k20.json 2> gpurun_out/bench_f_splitv1_k20.err
SMGX_EVENT_PATH=split timeout 300 python bench.py --steps 2000 --warmup 5 --lanes 1 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_f_split_k2000_l1.json 2> gpurun_out/bench_f_split_k2000_l1.err
SMGX_EVENT_PATH=split timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_f.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
SMGX_EVENT_PATH=split timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_search2 -s 12 -c 2 -o gpurun_out/search2_r02f -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | tail -8
for f in gpurun_out/bench_g_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'e2e %.3g'%d['e2e']['value'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1), 'launches', d['gpu_launches'])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
grep -E "event_search|hash_blocks" gpurun_out/launches_f.csv | awk -F'","' '{print $5, $(NF-4), $NF}' | sed 's/"//g' | sort | uniq -c | sort -rn | head -12
