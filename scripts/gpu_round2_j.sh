#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --config tet10 --scale 0.05 --steps 3 --check 2>&1 | tail -4) > gpurun_out/r2j_tet10_small_g2.log 2>&1
(timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 2 --config tet10 --steps 5 --check 2>&1 | tail -4) > gpurun_out/r2j_tet10_g2.log 2>&1
for f in gpurun_out/r2j_tet10_small_g2.log gpurun_out/r2j_tet10_g2.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], d.get("checks")); print(d["config"]["ranks"])
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
