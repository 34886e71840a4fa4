#!/usr/bin/env python
"""Host-side cost of one end-to-end step: bench.py's hex8 run on a SMALL box (kernel time negligible) under cProfile.
   python scripts/e2e_host_overhead.py --n 32 --steps 300          (also under torchrun, rank 0 prints)"""
import cProfile
import io
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402

if __name__ == "__main__":
    if "--no-cpu-baseline" not in sys.argv:
        sys.argv.append("--no-cpu-baseline")
    args = bench.parse()
    pr = cProfile.Profile()
    pr.enable()
    bench.run_hex8(args)
    pr.disable()
    if int(os.environ.get("RANK", "0")) == 0:
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(28)
        print(s.getvalue()[:6000])
