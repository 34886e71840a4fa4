#!/bin/bash
# GPU round trip A: all GPU tests, smoke, headline bench, other configs at full size
mkdir -p gpurun_out
(timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25) > gpurun_out/r2a_tests.log 2>&1
(timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) > gpurun_out/r2a_smoke.log 2>&1
(timeout 600 python bench.py --check 2>&1 | tail -2) > gpurun_out/r2a_bench_hex8.log 2>&1
for c in heat_tet4 j2_plate tet10; do
  (timeout 600 python bench.py --config $c --check --steps 5 2>&1 | tail -2) > gpurun_out/r2a_bench_$c.log 2>&1
done
tail -5 gpurun_out/r2a_tests.log; cat gpurun_out/r2a_smoke.log; for f in gpurun_out/r2a_bench_*.log; do echo $f; tail -c 1500 $f; done
