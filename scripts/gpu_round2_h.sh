#!/bin/bash
mkdir -p gpurun_out
(timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "j2 or plastic or octet" 2>&1 | tail -5) > gpurun_out/r2h_tests_j2.log 2>&1
(FDK_TET10_BIG=1 timeout 900 python -m pytest tests -m gpu -q -k "tet10" 2>&1 | tail -8) > gpurun_out/r2h_tests_tet10.log 2>&1
(timeout 600 python bench.py --config j2_plate --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2h_bench_j2.log 2>&1
(timeout 600 python bench.py --config tet10 --check --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2h_bench_tet10_generic.log 2>&1
(FDK_TET10_BIG=1 timeout 600 python bench.py --config tet10 --check --steps 5 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2h_bench_tet10_big.log 2>&1
cat gpurun_out/r2h_tests_j2.log gpurun_out/r2h_tests_tet10.log
for f in gpurun_out/r2h_bench_j2.log gpurun_out/r2h_bench_tet10_generic.log gpurun_out/r2h_bench_tet10_big.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"), d["config"].get("first_call_s"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
