#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo8.txt 2>&1
for n in 8 4; do
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2951$n bench.py --gpus $n --steps 20 --warmup 5 ) > gpurun_out/bench_n$n.json 2> gpurun_out/bench_n$n.err
done
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29528 bench.py --gpus 8 --steps 2000 --warmup 5 ) > gpurun_out/bench_n8_k2000.json 2> gpurun_out/bench_n8_k2000.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29529 bench.py --impl reference --gpus 8 --steps 20 --warmup 5 ) > gpurun_out/bench_n8_ref.json 2> gpurun_out/bench_n8_ref.err
for f in gpurun_out/bench_n8.json gpurun_out/bench_n4.json gpurun_out/bench_n8_k2000.json gpurun_out/bench_n8_ref.json; do echo "== $f"; grep "^{" $f | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','n_gpus','ms_per_step') if k in d}, 'e2e', d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'), 'parity', (d.get('parity_checked') or {}).get('mismatches'), 'numa', d.get('config',{}).get('numa_node'))
print('sharded', json.dumps(d.get('sharded'))[:900])
"; tail -3 ${f%.json}.err; done
