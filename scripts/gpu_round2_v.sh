#!/bin/bash
# 2 GPUs: heat_tet4 and j2_plate sharded by RCB (small: gathered D against one GPU; full size: timing); adapter tests on GPU 0
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611"
(timeout 600 python -m pytest tests/test_adapter_gpu.py -m gpu -q -x 2>&1 | tail -5) > gpurun_out/r2v_tests.log 2>&1
for c in heat_tet4 j2_plate; do
  (timeout 600 $TR bench.py --config $c --gpus 2 --steps 3 --warmup 3 --scale 0.015 --check 2>&1 | tail -3) > gpurun_out/r2v_${c}_small.json 2>&1
  (timeout 900 $TR bench.py --config $c --gpus 2 --steps 10 --warmup 3 --check 2>&1 | tail -3) > gpurun_out/r2v_${c}_n2.json 2>&1
done
cat gpurun_out/r2v_tests.log
for f in gpurun_out/r2v_*_small.json gpurun_out/r2v_*_n2.json; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("  ms/step", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "kernel_ms", d["roofline"]["kernel_ms"], d.get("checks"), d["config"].get("ranks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
