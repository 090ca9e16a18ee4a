#!/bin/bash
# round-2 GPU pass A: full GPU test-suite, bench variants at the driver's --steps 20 --warmup 5, launch list + one full ncu capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_f4.json 2> gpurun_out/bench_f4.err
SMGX_FUSED_MINB=3 timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in > gpurun_out/bench_f3.json 2> gpurun_out/bench_f3.err
SMGX_EVENT_PATH=split timeout 300 python bench.py --steps 20 --warmup 5 --no-text-in > gpurun_out/bench_split.json 2> gpurun_out/bench_split.err
timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_f4_k2000.json 2> gpurun_out/bench_f4_k2000.err
SMGX_FUSED_MINB=3 timeout 300 python bench.py --steps 2000 --warmup 10 --no-text-in --no-cpu-baseline > gpurun_out/bench_f3_k2000.json 2> gpurun_out/bench_f3_k2000.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_fused.csv python bench.py --steps 20 --warmup 5 --no-text-in --no-cpu-baseline > gpurun_out/b_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:event_fused -s 6 -c 2 -o gpurun_out/fused_r02 -f python bench.py --steps 20 --warmup 5 --no-text-in --no-cpu-baseline > gpurun_out/b_under_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
tail -3 gpurun_out/pytest_gpu.log
cat gpurun_out/bench_f4.json | head -c 3000
