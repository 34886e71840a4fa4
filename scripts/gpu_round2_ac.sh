#!/bin/bash
# heat kernel: temperature loads hoisted next to the coordinate loads; 5 or 6 resident CTAs; + ncu of the tet10 generic kernel
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -q -x -k "heat or thermal or delaunay or smallest" 2>&1 | tail -3) > gpurun_out/r2ac_tests.log 2>&1
for v in 6 5; do
  (FDK_HEAT_MINB=$v timeout 600 python bench.py --config heat_tet4 --steps 20 --warmup 3 --check --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2ac_heat_minb$v.json 2>&1
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:^k_assemble$ -s 2 -c 1 -o gpurun_out/r2ac_ncu_tet10 python bench.py --config tet10 --scale 0.1 --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r2ac_ncu_tet10.log 2>&1
cat gpurun_out/r2ac_tests.log
python - <<'PY'
import json
for v in (6, 5):
    try:
        d = json.loads([l for l in open(f"gpurun_out/r2ac_heat_minb{v}.json").read().strip().split("\n") if l.startswith("{")][-1])
        print("minb", v, "ms/step", d["ms_per_step"], "kernel_ms (K only)", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], d.get("checks"))
    except Exception as e:
        print(v, "parse error", e)
PY
ls -la gpurun_out/r2ac_ncu_tet10*
