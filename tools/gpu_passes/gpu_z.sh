#!/bin/bash
mkdir -p gpurun_out
for i in 1 2 3 4 5 6 7 8; do timeout 100 tests/cpp/test_batcher 32 400 1 50 0 > gpurun_out/tb_$i.out 2> gpurun_out/tb_$i.err; echo "run $i rc=$?"; tail -c 300 gpurun_out/tb_$i.err; done
timeout 100 tests/cpp/test_batcher 3 2000 700 200 0 > gpurun_out/tb_w.out 2> gpurun_out/tb_w.err; echo "wide rc=$?"; tail -c 300 gpurun_out/tb_w.err
timeout 100 tests/cpp/test_batcher 8 1500 64 100 0 > gpurun_out/tb_p.out 2> gpurun_out/tb_p.err; echo "pool rc=$?"; tail -c 300 gpurun_out/tb_p.err
timeout 300 python -m pytest tests/test_cpp_mirror.py -q -m gpu 2>&1 | tail -2
