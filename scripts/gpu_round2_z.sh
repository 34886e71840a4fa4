#!/bin/bash
# 8 GPUs: the three unstructured-partition configurations (RCB) + the hex8 headline once more
mkdir -p gpurun_out
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29631"
for c in heat_tet4 j2_plate tet10; do
  (timeout 900 $TR bench.py --config $c --gpus $N --steps 10 --warmup 3 --check 2>&1 | tail -2) > gpurun_out/r2z_${c}_n$N.json 2>&1
done
(timeout 900 $TR bench.py --gpus $N --steps 20 --warmup 3 --check 2>&1 | tail -2) > gpurun_out/r2z_hex8_n$N.json 2>&1
for f in gpurun_out/r2z_*_n$N.json; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    r = d["config"].get("ranks")
    print("  ms/step", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "kernel_ms", d["roofline"].get("kernel_ms"), d.get("checks"), (r[0] if r else None))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
