#!/bin/bash
# round-2 GPU pass P: full test-suite + the driver's two bench arms as the driver runs them
mkdir -p gpurun_out
( time timeout 2400 python -m pytest tests -m gpu -q --maxfail=15 ) > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
( time timeout 900 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_p_reference.json 2> gpurun_out/bench_p_reference.err
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 5 ) > gpurun_out/bench_p_smgx.json 2> gpurun_out/bench_p_smgx.err
( time timeout 900 python bench.py ) > gpurun_out/bench_p_default.json 2> gpurun_out/bench_p_default.err
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
grep -E "^FAILED|passed|failed" gpurun_out/pytest_gpu.log | tail -8
tail -2 gpurun_out/smoke.log
for f in gpurun_out/bench_p_*.json; do echo "== $f"; head -c 2500 $f; echo; tail -4 ${f%.json}.err; done
