#!/bin/bash
# round-2 final collection: ncu of the final pair, sanitizer, sweep, wide-fleet point, secondary modes
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"hash_blocks|event_search2" -s 14 -c 2 -o gpurun_out/split_r02w -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_w.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 5 --workers 256 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/bench_w_w256.json 2> gpurun_out/bench_w_w256.err
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_event_select.py tests/test_gpu_tree_handle.py tests/test_gpu_feedback.py -m gpu -q -x -k "not 256-3" ) > gpurun_out/sanitizer_memcheck.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck.log
( time timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_scale.py -m gpu -q -x -k "random_parity or duplicate" ) > gpurun_out/sanitizer_memcheck2.log 2>&1
echo "memcheck rc=$?" >> gpurun_out/sanitizer_memcheck2.log
( time timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_event_select.py -m gpu -q -x -k "random_select_parity" ) > gpurun_out/sanitizer_racecheck.log 2>&1
echo "racecheck rc=$?" >> gpurun_out/sanitizer_racecheck.log
( time timeout 1500 python tools/sweep.py ) > gpurun_out/sweep_r02.md 2> gpurun_out/sweep_r02.err
( timeout 600 python tools/bench_modes.py tree ) > gpurun_out/mode_tree_r02.log 2>&1
( timeout 600 python tools/bench_modes.py text ) > gpurun_out/mode_text_r02.log 2>&1
( timeout 600 python tools/bench_modes.py ingest ) > gpurun_out/mode_ingest_r02.log 2>&1
tail -4 gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_memcheck2.log gpurun_out/sanitizer_racecheck.log
cat gpurun_out/sweep_r02.md
tail -2 gpurun_out/mode_tree_r02.log gpurun_out/mode_text_r02.log gpurun_out/mode_ingest_r02.log | cut -c1-700
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_w_w256.json'))
print('W=256', 'value %.4g'%d['value'], 'frac %.3f'%d['roofline']['frac'], 'e2e %.3g'%d['e2e']['value'])
PY
