#!/bin/bash
# last sanity of HEAD: smoke, the default path's parity tests, power_of_two, the driver's bench command
mkdir -p gpurun_out
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 200 python -m pytest tests/test_gpu_event_select.py tests/test_gpu_power_of_two.py tests/test_gpu_feedback.py -q -m gpu -x 2>&1 | tail -2
timeout 200 python bench.py --steps 20 --warmup 5 --no-text-in --no-per-request > gpurun_out/head_k20.json 2> gpurun_out/head_k20.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/head_k20.json').read().strip().splitlines()[-1])
    print('value %.4g'%d['value'],'frac %.3f'%d['roofline']['frac'],'parity',d['parity_checked']['mismatches'],'e2e %.4g'%d['e2e']['value'],'launches',d['gpu_launches'])
except Exception as e:
    print('ERR',e,open('gpurun_out/head_k20.err').read()[-500:])
PY
