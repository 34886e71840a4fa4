#!/bin/bash
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2s_tests.log 2>&1
(timeout 600 python bench.py --config j2_plate --check --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2s_bench_j2.log 2>&1
(timeout 600 python scripts/variants_bench.py 2>&1 | tail -8) > gpurun_out/r2s_variants.log 2>&1
(FDK_ELEM_FORCE_LOOP=1 timeout 600 python scripts/variants_bench.py 2>&1 | tail -8) > gpurun_out/r2s_variants_loop.log 2>&1
tail -6 gpurun_out/r2s_tests.log
python - gpurun_out/r2s_bench_j2.log <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
echo "--- variants (GP-parallel residual)"; cat gpurun_out/r2s_variants.log; echo "--- variants (element loop)"; cat gpurun_out/r2s_variants_loop.log
