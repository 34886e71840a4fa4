#!/bin/bash
mkdir -p gpurun_out
for sh in 4 5; do
(FDK_PLAN_SHIFT=$sh timeout 600 python bench.py --config tet10 --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2i_bench_tet10_shift$sh.log 2>&1
done
for f in gpurun_out/r2i_bench_tet10_shift4.log gpurun_out/r2i_bench_tet10_shift5.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"])
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
