#!/bin/bash
# structured-tangent kernel with the elastic-warp shortcut: parity + j2_plate timing
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -6) > gpurun_out/r2y_tests.log 2>&1
(timeout 900 python bench.py --config j2_plate --steps 10 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2y_j2.json 2>&1
cat gpurun_out/r2y_tests.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2y_j2.json").read().strip().split("\n") if l.startswith("{")][-1])
print("j2_plate ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d.get("checks"))
PY
