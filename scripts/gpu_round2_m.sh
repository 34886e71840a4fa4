#!/bin/bash
mkdir -p gpurun_out
N=${NG:-2}
for ex in copy peer; do
(timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --edge 48 --steps 5 --check --no-cpu-baseline --exchange $ex 2>&1 | tail -3) > gpurun_out/r2m_n48_g${N}_$ex.log 2>&1
(timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29552 bench.py --gpus $N --steps 10 --check --exchange $ex 2>&1 | tail -3) > gpurun_out/r2m_n200_g${N}_$ex.log 2>&1
done
for f in gpurun_out/r2m_n48_g${N}_copy.log gpurun_out/r2m_n48_g${N}_peer.log gpurun_out/r2m_n200_g${N}_copy.log gpurun_out/r2m_n200_g${N}_peer.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("n_gpus", d["n_gpus"], "ms/step", d["ms_per_step"], "value", d["value"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
