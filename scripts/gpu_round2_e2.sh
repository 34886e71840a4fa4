#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --edge 48 --steps 5 --check --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2e_n48_g2.log 2>&1
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --edge 48 --steps 5 --check --exchange nccl --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2e_n48_g2_nccl.log 2>&1
for f in gpurun_out/r2e_n48_g2.log gpurun_out/r2e_n48_g2_nccl.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
