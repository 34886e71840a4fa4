#!/bin/bash
mkdir -p gpurun_out
for th in 1024 768; do
(FDK_TET10_BIG=1 FDK_BIG_THREADS=$th timeout 900 python -m pytest tests -m gpu -q -k "tet10" 2>&1 | tail -4) > gpurun_out/r2p_tests_tet10_$th.log 2>&1
(FDK_TET10_BIG=1 FDK_BIG_THREADS=$th timeout 600 python bench.py --config tet10 --check --steps 5 --no-cpu-baseline 2>&1 | tail -3) > gpurun_out/r2p_bench_tet10_$th.log 2>&1
done
(timeout 600 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2p_bench_hex8.log 2>&1
cat gpurun_out/r2p_tests_tet10_1024.log gpurun_out/r2p_tests_tet10_768.log
for f in gpurun_out/r2p_bench_tet10_1024.log gpurun_out/r2p_bench_tet10_768.log gpurun_out/r2p_bench_hex8.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
