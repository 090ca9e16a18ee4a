#!/bin/bash
# round-2 GPU pass D: gated timing, tile kernel with L2 prefetch, feedback + tree-handle tests, tokenizer launch list
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 900 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_d_default.json 2> gpurun_out/bench_d_default.err
for v in "8 4" "32 4" "0 4"; do
  set -- $v
  SMGX_FUSED_TILE=$1 SMGX_FUSED_MINB=$2 timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_d_t$1_m$2.json 2> gpurun_out/bench_d_t$1_m$2.err
done
SMGX_EVENT_PATH=split timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request > gpurun_out/bench_d_split.json 2> gpurun_out/bench_d_split.err
for t in 8 16 32; do
SMGX_FUSED_TILE=$t timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_d_k2000_t$t.json 2> gpurun_out/bench_d_k2000_t$t.err
done
SMGX_EVENT_PATH=split timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_d_k2000_split.json 2> gpurun_out/bench_d_k2000_split.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_tile -s 4 -c 1 -o gpurun_out/tile_r02d -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_d.csv python bench.py --steps 20 --warmup 5 --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
for f in gpurun_out/bench_d_*.json; do python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1]))
    print(sys.argv[1], 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'region_ms', [round(x,4) for x in d['region_ms']], 'ungated', [round(x,4) for x in d['region_ms_ungated']], 'e2e %.3g'%d['e2e']['value'], 'parity', d.get('parity_checked',{}).get('mismatches'))
    for k in ('text_in','per_request'):
        if k in d: print('   ',k, json.dumps(d[k])[:900])
except Exception as e:
    print(sys.argv[1], 'ERR', e)
PY
done
