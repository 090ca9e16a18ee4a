#!/bin/bash
# final validation of the round on one GPU: smoke, the whole GPU suite, the driver's two bench commands and the default run
mkdir -p gpurun_out
timeout 180 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE OK')" 2>&1 | tail -1
timeout 900 python -m pytest tests -q -m gpu -x -q > gpurun_out/pytest_final.log 2>&1; tail -2 gpurun_out/pytest_final.log
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/final_k20.json 2> gpurun_out/final_k20.err
timeout 400 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/final_ref.json 2> gpurun_out/final_ref.err
timeout 400 python bench.py > gpurun_out/final_k2000.json 2> gpurun_out/final_k2000.err
python - <<'PY'
import json
for f in ('final_k20','final_k2000','final_ref'):
    try:
        d=json.loads(open(f'gpurun_out/{f}.json').read().strip().splitlines()[-1])
        print(f, 'value %.4g'%d['value'], 'e2e %.4g'%d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'), 'parity', d.get('parity_checked',{}).get('mismatches'), 'launches', d.get('gpu_launches'),
              'text_in', (d.get('text_in') or {}).get('value'), 'per_request', {k:v for k,v in (d.get('per_request') or {}).items() if k in ('value','p50_us','p99_us')}, 'cpu', (d.get('cpu_baseline') or {}).get('value'), d.get('clocks'))
    except Exception as e:
        print(f, 'ERR', e, open(f'gpurun_out/{f}.err').read()[-600:])
PY
