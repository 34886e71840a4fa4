#!/bin/bash
# One GPU-box round trip: parity tests, phase clocks of the cluster kernel, headline bench (+ property checks).
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -3)
python scripts/phase_clocks.py --n 100 ${CLK_ARGS:-}
timeout 300 python bench.py --no-cpu-baseline --check ${BENCH_ARGS:-} 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('ms/step', d['ms_per_step'], 'Melem/s', d['value'], 'frac', d['roofline']['frac'], 'e2e', d['e2e']['value'], d.get('checks'))"
