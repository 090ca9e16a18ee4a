#!/bin/bash
# round-2 GPU pass C: full GPU test-suite, tiled event kernel variants at the driver's --steps 20 --warmup 5, ncu, per-request batcher
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 --no-text-in > gpurun_out/bench_c_default.json 2> gpurun_out/bench_c_default.err
for v in "8 4" "32 4" "16 3" "32 3" "0 4"; do
  set -- $v
  SMGX_FUSED_TILE=$1 SMGX_FUSED_MINB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in > gpurun_out/bench_c_t$1_m$2.json 2> gpurun_out/bench_c_t$1_m$2.err
done
SMGX_EVENT_PATH=split timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in > gpurun_out/bench_c_split.json 2> gpurun_out/bench_c_split.err
timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_c_k2000.json 2> gpurun_out/bench_c_k2000.err
SMGX_FUSED_TILE=32 timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_c_k2000_t32.json 2> gpurun_out/bench_c_k2000_t32.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_c_k200.json 2> gpurun_out/bench_c_k200.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_tile_c.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tile -s 4 -c 1 -o gpurun_out/tile_r02c -f python bench.py --steps 20 --warmup 5 --no-text-in --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
cd tests/cpp
rm -f ../../gpurun_out/batcher_c.jsonl
for a in "32 2000 1 50 1 3 0 1500" "32 2000 1 50 1 3 0 500" "32 2000 1 50 1 2 0 3000" "32 2000 1 50 0" "16 16384 512 100 1 3 0 1500" "16 16384 512 100 0" "64 4096 64 100 1 3 0 1500" "64 4096 64 100 0" "128 2048 64 200 1 3 0 1500" "8 4000 1 50 1 3 0 1500" "1 4000 1 50 1 3 0 1500"; do
  timeout 300 ./test_batcher $a >> ../../gpurun_out/batcher_c.jsonl 2>> ../../gpurun_out/batcher_c.err
done
cd ../..
tail -5 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_c_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'e2e %.3g'%d['e2e']['value'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1))
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
cat gpurun_out/batcher_c.jsonl | cut -c150-760
