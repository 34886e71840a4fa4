#!/bin/bash
# generic cluster kernel: parallel residual reduction + 4-chain heavy pre-reduction -> parity, tet10 / quad4 timing
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -8) > gpurun_out/r2x_tests.log 2>&1
(timeout 900 python bench.py --config tet10 --steps 10 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2x_tet10.json 2>&1
cat gpurun_out/r2x_tests.log
python - <<'PY'
import json
d = json.loads([l for l in open("gpurun_out/r2x_tet10.json").read().strip().split("\n") if l.startswith("{")][-1])
print("tet10 ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d.get("checks"))
PY
