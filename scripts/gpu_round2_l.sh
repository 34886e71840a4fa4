#!/bin/bash
mkdir -p gpurun_out
for n in 8 4; do
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n bench.py --gpus $n --steps 10 --warmup 3 --check 2>&1 | tail -3) > gpurun_out/r2l_hex8_g$n.log 2>&1
done
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 8 --config tet10 --steps 5 --check 2>&1 | tail -3) > gpurun_out/r2l_tet10_g8.log 2>&1
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 bench.py --gpus 8 --impl reference --steps 2 --warmup 1 2>&1 | tail -2) > gpurun_out/r2l_ref_g8.log 2>&1
for f in gpurun_out/r2l_hex8_g8.log gpurun_out/r2l_hex8_g4.log gpurun_out/r2l_tet10_g8.log gpurun_out/r2l_ref_g8.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d.get("checks"), d.get("impl"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
