#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hash_blocks|event_search2" -s 14 -c 2 -o gpurun_out/split_r02z -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/z_under_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_z.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/z_under_ncu.log 2>&1
grep -c "hash_blocks\|event_search2" gpurun_out/launches_z.csv
# sanitizer over the kernels touched this round (pair with PDL, stream, hs)
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_event_select.py tests/test_gpu_feedback.py -q -m gpu -x -k "random or feedback" > gpurun_out/sanitizer_z_pair.log 2>&1; echo "memcheck pair rc=$?"; tail -3 gpurun_out/sanitizer_z_pair.log
SMGX_EVENT_PATH=stream timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_event_select.py -q -m gpu -x -k "random" > gpurun_out/sanitizer_z_stream.log 2>&1; echo "memcheck stream rc=$?"; tail -3 gpurun_out/sanitizer_z_stream.log
SMGX_EVENT_PATH=hs timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_event_select.py -q -m gpu -x -k "random" > gpurun_out/sanitizer_z_hs.log 2>&1; echo "memcheck hs rc=$?"; tail -3 gpurun_out/sanitizer_z_hs.log
SMGX_EVENT_PATH=stream timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_event_select.py -q -m gpu -x -k "random" > gpurun_out/sanitizer_z_stream_race.log 2>&1; echo "racecheck stream rc=$?"; tail -3 gpurun_out/sanitizer_z_stream_race.log
