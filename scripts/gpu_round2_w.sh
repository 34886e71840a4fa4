#!/bin/bash
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29621"
(timeout 600 python scripts/e2e_host_overhead.py --n 32 --steps 400 --warmup 3 2>&1 | tail -60) > gpurun_out/r2w_host1.log 2>&1
(timeout 600 $TR scripts/e2e_host_overhead.py --edge 32 --gpus 2 --steps 400 --warmup 3 2>&1 | tail -60) > gpurun_out/r2w_host2.log 2>&1
cat gpurun_out/r2w_host1.log; cat gpurun_out/r2w_host2.log
