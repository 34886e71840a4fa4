#!/bin/bash
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2f_tests.log 2>&1
(timeout 600 python bench.py --config j2_plate --check --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2f_bench_j2.log 2>&1
tail -15 gpurun_out/r2f_tests.log
python - gpurun_out/r2f_bench_j2.log <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "step_frac", d["roofline"]["step_frac"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
