#!/bin/bash
mkdir -p gpurun_out
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/eval_sharded.py 4000 8 > gpurun_out/eval_sharded_8gpu.json 2> gpurun_out/eval_sharded_8gpu.err
echo "eval8 rc=$?"; tail -1 gpurun_out/eval_sharded_8gpu.json
