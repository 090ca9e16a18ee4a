#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu --maxfail=10 -x -q > gpurun_out/pytest_full_z.log 2>&1; tail -5 gpurun_out/pytest_full_z.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"hash_blocks|event_search2" -s 14 -c 2 -o gpurun_out/split_r02z -f python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/z_under_ncu_full.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_z.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request --no-cpu-baseline > gpurun_out/z_under_ncu.log 2>&1
grep -c "hash_blocks\|event_search2" gpurun_out/launches_z.csv
