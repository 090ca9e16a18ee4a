#!/bin/bash
# round 2, pass Y: the one-launch path (event_hs_kernel) — parity at config-2 scale, then A/B against the pair
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" > gpurun_out/smoke_y.log 2>&1 || { echo SMOKE FAILED; tail -20 gpurun_out/smoke_y.log; exit 1; }
tail -2 gpurun_out/smoke_y.log
timeout 300 python -m pytest tests/test_gpu_event_select.py -q -m gpu -x > gpurun_out/pytest_y0.log 2>&1 || { echo EVENT_SELECT FAILED; tail -30 gpurun_out/pytest_y0.log; exit 1; }
tail -2 gpurun_out/pytest_y0.log
timeout 600 python -m pytest tests/test_gpu_scale.py -q -m gpu --maxfail=10 -k "stream or hs or split or duplicate or event_select" > gpurun_out/pytest_y.log 2>&1
tail -5 gpurun_out/pytest_y.log
run() { name=$1; k=$2; shift; shift; env "$@" timeout 300 python bench.py --steps $k --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_y_$name.json 2> gpurun_out/bench_y_$name.err; }
run st_k20 20 SMGX_EVENT_PATH=stream
run hs_k20 20 SMGX_EVENT_PATH=hs
run split_k20 20 SMGX_EVENT_PATH=split
run st_k200 200 SMGX_EVENT_PATH=stream
run st_k2000 2000 SMGX_EVENT_PATH=stream
run split_k2000 2000 SMGX_EVENT_PATH=split
for k in 1 5 10 32; do run st_n$k $k SMGX_EVENT_PATH=stream; done
for f in gpurun_out/bench_y_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_us', [round(x*1e3,1) for x in d['region_ms']], 'launches', d['gpu_launches'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1), 'e2e %.4g'%d['e2e']['value'])
except Exception as e:
    print(sys.argv[1], 'ERR', e, open(sys.argv[1].replace('.json','.err')).read()[-400:])
PY
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"event_stream" -s 7 -c 1 -o gpurun_out/stream_r02y -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/y_under_ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_y.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/y_under_ncu.log 2>&1
grep -c event_stream gpurun_out/launches_y.csv
