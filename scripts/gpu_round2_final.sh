#!/bin/bash
# final single-GPU evidence of round 2: GPU tests, smoke, the four bench lines (with property checks), the reference arm
mkdir -p gpurun_out
(timeout 1800 python -m pytest tests -m gpu -q 2>&1 | tail -12) > gpurun_out/r2_gpu_tests_final.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2) >> gpurun_out/r2_gpu_tests_final.log 2>&1
(timeout 900 python bench.py --steps 20 --warmup 3 --check 2>&1 | tail -1) > gpurun_out/r2_bench_hex8.json 2>&1
for c in heat_tet4 j2_plate tet10; do
  (timeout 900 python bench.py --config $c --steps 10 --warmup 3 --check 2>&1 | tail -1) > gpurun_out/r2_bench_$c.json 2>&1
done
(timeout 900 python bench.py --impl reference --steps 5 --warmup 1 2>&1 | tail -1) > gpurun_out/r2_bench_reference_hex8.json 2>&1
(timeout 900 python bench.py --jitter 1 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1) > gpurun_out/r2_bench_hex8_jitter.json 2>&1
cat gpurun_out/r2_gpu_tests_final.log
for f in gpurun_out/r2_bench_hex8.json gpurun_out/r2_bench_hex8_jitter.json gpurun_out/r2_bench_heat_tet4.json gpurun_out/r2_bench_j2_plate.json gpurun_out/r2_bench_tet10.json gpurun_out/r2_bench_reference_hex8.json; do echo $f; python - "$f" <<'PY'
import sys, json
try:
    d = json.loads([l for l in open(sys.argv[1]).read().strip().split("\n") if l.startswith("{")][-1])
    r = d.get("roofline") or {}
    print("  ms/step", round(d["ms_per_step"],3), "value", round(d["value"],2), "e2e", round(d["e2e"]["value"],2), "kernel_ms", r.get("kernel_ms"), "frac", r.get("frac"), "traffic", r.get("traffic"), "cpu", (d.get("cpu_baseline") or {}).get("value"), (d.get("cpu_baseline") or {}).get("kind"), d.get("checks"), d.get("clocks"))
except Exception as e:
    print("parse error", e); print(open(sys.argv[1]).read()[-3000:])
PY
done
