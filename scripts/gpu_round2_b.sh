#!/bin/bash
# GPU round trip B: all GPU tests (no -x), heat config with the row-owner kernel vs the cluster kernel
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -60) > gpurun_out/r2b_tests.log 2>&1
(timeout 600 python bench.py --config heat_tet4 --check --steps 5 2>&1 | tail -1) > gpurun_out/r2b_bench_heat_rows.log 2>&1
(FDK_HEAT_TET4_ROWS=0 timeout 600 python bench.py --config heat_tet4 --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2b_bench_heat_cluster.log 2>&1
tail -40 gpurun_out/r2b_tests.log
for f in gpurun_out/r2b_bench_heat_*.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "step_frac", d["roofline"]["step_frac"], "e2e", d["e2e"]["value"], d.get("checks"), "first", d["config"]["first_call_s"])
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-2000:])
PY
done
