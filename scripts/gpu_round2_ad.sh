#!/bin/bash
# tet10: asynchronous local-connectivity fetch; heat: 5 resident CTAs (default) against 4; full GPU test suite
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) > gpurun_out/r2ad_tests.log 2>&1
(timeout 900 python bench.py --config tet10 --steps 10 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2ad_tet10.json 2>&1
for v in 5 4; do
  (FDK_HEAT_MINB=$v timeout 600 python bench.py --config heat_tet4 --steps 20 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2ad_heat_minb$v.json 2>&1
done
cat gpurun_out/r2ad_tests.log
python - <<'PY'
import json
for f in ("r2ad_tet10", "r2ad_heat_minb5", "r2ad_heat_minb4"):
    try:
        d = json.loads([l for l in open(f"gpurun_out/{f}.json").read().strip().split("\n") if l.startswith("{")][-1])
        print(f, "ms/step", d["ms_per_step"], "kernel_ms (K only)", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
    except Exception as e:
        print(f, "parse error", e); print(open(f"gpurun_out/{f}.json").read()[-1500:])
PY
