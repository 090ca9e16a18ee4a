#!/bin/bash
# round-2 GPU pass B: full GPU test-suite, fused-kernel variants at the driver's --steps 20 --warmup 5, ncu, per-request batcher
mkdir -p gpurun_out
( time timeout 1800 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_b_default.json 2> gpurun_out/bench_b_default.err
for v in "4 0" "4 2" "3 1" "3 2"; do
  set -- $v
  SMGX_FUSED_MINB=$1 SMGX_FUSED_PF=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in > gpurun_out/bench_b_m$1_pf$2.json 2> gpurun_out/bench_b_m$1_pf$2.err
done
timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_b_k2000.json 2> gpurun_out/bench_b_k2000.err
timeout 300 python bench.py --steps 200 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_b_k200.json 2> gpurun_out/bench_b_k200.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fused_b.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_fused -s 6 -c 1 -o gpurun_out/fused_r02b -f python bench.py --steps 20 --warmup 5 --no-text-in --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
cd tests/cpp
for a in "32 2000 1 50 1 3 0" "32 2000 1 50 1 2 0" "32 2000 1 50 1 3 5" "32 2000 1 50 0" "16 16384 512 100 1 3 0" "16 16384 512 100 0" "64 4096 64 100 1 3 0" "64 4096 64 100 0" "128 2048 64 200 1 3 0" "8 4000 1 50 1 3 0" "1 4000 1 50 1 3 0"; do
  timeout 300 ./test_batcher $a >> ../../gpurun_out/batcher_b.jsonl 2>> ../../gpurun_out/batcher_b.err
done
cd ../..
tail -3 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_b_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'e2e %.3g'%d['e2e']['value'], 'parity', d.get('parity_checked',{}).get('mismatches'), 'lat', round(d['latency']['device_resident_p50_us'],1))
    if 'text_in' in d: print('   text_in', json.dumps(d['text_in'])[:600])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
cat gpurun_out/batcher_b.jsonl | cut -c150-700
