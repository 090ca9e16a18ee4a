#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo.txt 2>&1
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_q_n2.json 2> gpurun_out/bench_q_n2.err
( time timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 20 --warmup 5 ) > gpurun_out/bench_q_n2_ref.json 2> gpurun_out/bench_q_n2_ref.err
( time timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 2000 --warmup 5 ) > gpurun_out/bench_q_n2_k2000.json 2> gpurun_out/bench_q_n2_k2000.err
timeout 600 python -m pytest tests/test_gpu_feedback.py tests/test_cpp_mirror.py -m gpu -q > gpurun_out/pytest_q.log 2>&1
tail -3 gpurun_out/pytest_q.log
for f in gpurun_out/bench_q_n2.json gpurun_out/bench_q_n2_k2000.json gpurun_out/bench_q_n2_ref.json; do echo "== $f"; grep -v "^$" $f | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print({k:d[k] for k in ('value','n_gpus','ms_per_step') if k in d}, 'e2e', d['e2e']['value'], 'frac', d.get('roofline',{}).get('frac'), 'parity', d.get('parity_checked'))
print('sharded', json.dumps(d.get('sharded'))[:1200])
"; tail -5 ${f%.json}.err; done
