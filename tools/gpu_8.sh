#!/bin/bash
# 8-GPU validation: the driver's launch line at N = 8 and N = 4 (replicas + the sharded config-4 leg), short
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4; do
timeout 420 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $n --steps 20 --warmup 5 > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
python - $n <<'PY'
import json,sys
n=sys.argv[1]
try:
    d=json.loads(open(f'gpurun_out/bench_n{n}.json').read().strip().splitlines()[-1])
    print('N',n,'value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'],'parity',d.get('parity_checked'),'numa',d['config'].get('numa_node'))
    print('  sharded', {k:d['sharded'][k] for k in d.get('sharded',{}) if k in ('value','ms_per_step','parity_checked','workers')} if 'sharded' in d else None)
except Exception as e:
    print('N',n,'ERR',e, open(f'gpurun_out/bench_n{n}.err').read()[-1500:])
PY
done
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --steps 2000 --warmup 5 --no-sharded > gpurun_out/bench_n8_k2000.json 2> gpurun_out/bench_n8_k2000.err
python - <<'PY'
import json
try:
    d=json.loads(open('gpurun_out/bench_n8_k2000.json').read().strip().splitlines()[-1])
    print('N 8 K2000 value %.4g'%d['value'],'e2e %.4g'%d['e2e']['value'],'frac %.3f'%d['roofline']['frac'])
except Exception as e:
    print('ERR',e, open('gpurun_out/bench_n8_k2000.err').read()[-800:])
PY
