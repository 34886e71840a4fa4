#!/bin/bash
# hex8 balanced kernel with the monomial-basis geometry phase (FDK_HEX_MONO): parity + timing
mkdir -p gpurun_out
(timeout 900 python bench.py --steps 20 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2ah_hex8.json 2>&1
(timeout 900 python bench.py --jitter 1 --steps 10 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2ah_hex8j.json 2>&1
(timeout 900 python bench.py --config j2_plate --steps 10 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2ah_j2.json 2>&1
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -12) > gpurun_out/r2ah_tests.log 2>&1
cat gpurun_out/r2ah_tests.log
python - <<'PY'
import json
for f in ("r2ah_hex8", "r2ah_hex8j", "r2ah_j2"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().split("\n") if l.startswith("{")][-1])
        print(f, "ms/step", d["ms_per_step"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], d.get("checks"))
    except Exception as e:
        print(f, "parse error", e, open(f"gpurun_out/{f}.json").read()[-1500:])
PY
