#!/bin/bash
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29500 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/bench_r02_8gpu.json 2> gpurun_out/bench_r02_8gpu.err
echo "bench8 rc=$?"; python -c "
import json; d=json.loads(open('gpurun_out/bench_r02_8gpu.json').read().strip().splitlines()[-1]); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/eval_sharded.py 4000 8 > gpurun_out/eval_sharded_8gpu.json 2> gpurun_out/eval_sharded_8gpu.err
echo "eval8 rc=$?"; tail -1 gpurun_out/eval_sharded_8gpu.json; tail -3 gpurun_out/eval_sharded_8gpu.err
