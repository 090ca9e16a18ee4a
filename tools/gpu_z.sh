#!/bin/bash
mkdir -p gpurun_out
SMGX_STREAM_DBG=72 SMGX_EVENT_PATH=stream timeout 300 python bench.py --steps 20 --warmup 5 --regions 1 --no-text-in --no-per-request --no-cpu-baseline 2>&1 | grep "^cta" > gpurun_out/trace_k20.txt
grep -v " n 7:" gpurun_out/trace_k20.txt | tail -33
