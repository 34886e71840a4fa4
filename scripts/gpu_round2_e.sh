#!/bin/bash
# 2 GPUs: fused-exchange path after the double-buffer change; distributed checks vs the one-GPU result at n = 48
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --edge 48 --steps 5 --check 2>&1 | tail -3) > gpurun_out/r2e_n48_g2.log 2>&1
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --edge 48 --steps 5 --check --exchange nccl 2>&1 | tail -3) > gpurun_out/r2e_n48_g2_nccl.log 2>&1
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --check 2>&1 | tail -3) > gpurun_out/r2e_n200_g2.log 2>&1
for v in 5 6 8; do (FDK_HEAT_MINB=$v timeout 600 python bench.py --config heat_tet4 --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2e_heat_minb$v.log 2>&1; done
for f in gpurun_out/r2e_n48_g2.log gpurun_out/r2e_n48_g2_nccl.log gpurun_out/r2e_n200_g2.log gpurun_out/r2e_heat_minb5.log gpurun_out/r2e_heat_minb6.log gpurun_out/r2e_heat_minb8.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
