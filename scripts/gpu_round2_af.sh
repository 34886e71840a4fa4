#!/bin/bash
# hex8 residual-only kernel in the monomial basis: parity (full GPU suite) + R-alone timing at 8 M elements for the register variants
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -4) > gpurun_out/r2af_tests.log 2>&1
for v in 0 2 3 4; do
  (FDK_ELEM_FORCE_HEX8=$v timeout 600 python scripts/variants_bench.py --n 200 2>&1 | tail -2 | head -1) > gpurun_out/r2af_var$v.json 2>&1
done
(timeout 600 python bench.py --config j2_plate --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2af_j2.json 2>&1
cat gpurun_out/r2af_tests.log
python - <<'PY'
import json
for v in (0, 2, 3, 4):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2af_var{v}.json").read().strip().split("\n") if l.startswith("{")][-1])
        print("FDK_ELEM_FORCE_HEX8 =", v, {k: d[k] for k in d if "ms" in k or "mesh" in k})
    except Exception as e:
        print(v, "parse error", e, open(f"gpurun_out/r2af_var{v}.json").read()[-800:])
d = json.loads([l for l in open("gpurun_out/r2af_j2.json").read().strip().split("\n") if l.startswith("{")][-1])
print("j2_plate ms/step", d["ms_per_step"])
PY
