#!/bin/bash
# GPU round trip C: all GPU tests, phase clocks + bench of the hex8 kernel (reflected geometry), heat rows kernel, ncu
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -80) > gpurun_out/r2c_tests.log 2>&1
(timeout 300 python scripts/phase_clocks.py --n 100 2>&1 | tail -14) > gpurun_out/r2c_clocks.log 2>&1
(timeout 600 python bench.py --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2c_bench_hex8.log 2>&1
(timeout 600 python bench.py --config heat_tet4 --check --steps 5 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2c_bench_heat.log 2>&1
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_assemble_iso -s 3 -c 1 -o gpurun_out/r2c_ncu_iso python bench.py --n 100 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu.log 2>&1)
(timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_heat_tet4 -s 2 -c 1 -o gpurun_out/r2c_ncu_heat python bench.py --config heat_tet4 --scale 0.1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2c_ncu_heat.log 2>&1)
tail -30 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_clocks.log
for f in gpurun_out/r2c_bench_hex8.log gpurun_out/r2c_bench_heat.log; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads(open(sys.argv[1]).read().strip().split("\n")[-1])
    print("ms/step", d["ms_per_step"], "value", d["value"], "kernel_ms", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-2000:])
PY
done
ls -la gpurun_out/*.ncu-rep
