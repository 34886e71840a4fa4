#!/bin/bash
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -40) > gpurun_out/r2d_tests.log 2>&1
(timeout 300 python scripts/phase_clocks.py --n 100 2>&1 | tail -9) > gpurun_out/r2d_clocks.log 2>&1
(timeout 600 python bench.py --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2d_bench_hex8.log 2>&1
tail -12 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_clocks.log
python - gpurun_out/r2d_bench_hex8.log <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-2000:])
PY
