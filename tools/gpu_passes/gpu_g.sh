#!/bin/bash
# round-2 GPU pass G: split path default — slot prefetch from the hash kernel, paired drains, two-lane pipeline
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
for k in 20 200 2000; do
timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_g_k$k.json 2> gpurun_out/bench_g_k$k.err
done
timeout 300 python bench.py --steps 20 --warmup 5 --lanes 1 --no-text-in --no-per-request > gpurun_out/bench_g_k20_l1.json 2> gpurun_out/bench_g_k20_l1.err
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_g.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"event_search2|hash_blocks" -s 14 -c 4 -o gpurun_out/split_r02g -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
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
