#!/usr/bin/env python
"""Diagnostic: where do the SM cycles of the cluster kernel go?

Builds an instrumented copy of the library (-DFDK_PHASE_CLOCKS -> fedoo_b200/_fdk_clk.so), assembles the
hex8 box a few times and prints, per phase, the cycles between consecutive barriers summed over all CTAs
(thread 0 of each CTA), as clocks per cluster and as a share of the kernel.  Run on the GPU box:

    python scripts/phase_clocks.py [--n 100] [--mma 1] [--build-only]
"""

import argparse
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
CLK_LIB = os.environ.get("FDK_CLK_LIB") or os.path.join(ROOT, "fedoo_b200", "_fdk_clk.so")  # FDK_CLK_LIB: a library built by hand
os.environ["FDK_LIB"] = CLK_LIB  # before fedoo_b200 is imported (fedoo_b200/_lib.py reads it at import time)

NAMES = ["-", "phase 0 wait", "phase 1 (geometry)", "phase 2 (blocks)", "staging / phase 2m", "phase 3a heavy",
         "phase 3b gather+store", "D tail + end", "-", "-"]  # fmt: skip


def build():
    from fedoo_b200 import build as b

    if os.environ.get("FDK_CLK_LIB"):
        return
    cmd = b.nvcc_command(out=CLK_LIB)
    cmd.insert(1, "-DFDK_PHASE_CLOCKS")
    if not os.path.exists(CLK_LIB) or any(os.path.getmtime(CLK_LIB) < os.path.getmtime(d) for d in b.DEPS):
        print(" ".join(cmd), flush=True)
        subprocess.run(cmd, check=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--mma", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--compute", default="all", choices=["all", "matrix", "vector"])
    ap.add_argument("--build-only", action="store_true")
    a = ap.parse_args()
    build()
    if a.build_only:
        return
    import numpy as np
    import torch

    import fedoo_b200 as fd
    from fedoo_b200 import _lib

    _lib.set_option("mma", a.mma)
    n = a.n
    fd.ModelingSpace("3D")
    nodes, elements = fd.meshgen.box_hex8(n + 1, n + 1, n + 1)
    mesh = fd.Mesh(nodes, elements, "hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="ElasticLaw")
    fd.weakform.StressEquilibrium("ElasticLaw", name="weakform")
    asm = fd.Assembly.create("weakform", "Domain", "hex8", name="Assembling", reuse_buffers=True)
    pb = fd.problem.Linear("Assembling")
    pb.set_X(np.random.default_rng(0).standard_normal(pb.n_dof) * 1e-3)
    asm.update(pb, compute="all")
    asm.vector_on_device = True
    plan = asm._plan(asm._saved_bloc_structure)
    lib = _lib.load()
    out = (C.c_ulonglong * 16)()
    _lib.check(lib.fdk_debug_phase_clocks(out, 16, 1), "phase clocks")
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        asm.assemble_global_mat(a.compute)
    ev1.record()
    torch.cuda.synchronize()
    ms = ev0.elapsed_time(ev1) / a.steps
    _lib.check(lib.fdk_debug_phase_clocks(out, 16, 1), "phase clocks")
    tot = sum(out[:10])
    ncl = plan.n_clusters * a.steps
    print(f"n={n} mma={a.mma}: {ms:.3f} ms/step (instrumented), {plan.n_clusters} clusters, caps {plan.caps}")
    print(f"total {tot / ncl:.0f} clk/cluster = {tot / (n**3 * a.steps):.0f} clk/element")
    for i in range(10):
        if out[i]:
            print(f"  [{i}] {NAMES[i]:24s} {out[i] / ncl:9.0f} clk/cluster  {100.0 * out[i] / tot:5.1f} %")


if __name__ == "__main__":
    main()
