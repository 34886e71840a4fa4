#!/bin/bash
mkdir -p gpurun_out
(timeout 600 python -m pytest tests -m gpu -q -k "heat or thermal" 2>&1 | tail -4) > gpurun_out/r2t_tests.log 2>&1
for v in 5 6; do (FDK_HEAT_MINB=$v timeout 600 python bench.py --config heat_tet4 --check --steps 10 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2t_heat_minb$v.log 2>&1; done
cat gpurun_out/r2t_tests.log
for f in gpurun_out/r2t_heat_minb5.log gpurun_out/r2t_heat_minb6.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
