#!/usr/bin/env python
"""Diagnostic: SM cycles per phase of the generic cluster kernel on the tet10 configuration (instrumented library)."""
import ctypes as C
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ["FDK_LIB"] = os.path.join(ROOT, "fedoo_b200", "_fdk_clk.so")
import numpy as np
import torch

import bench
import fedoo_b200 as fd
from fedoo_b200 import _lib

scale = float(sys.argv[1]) if len(sys.argv) > 1 else 0.2
nodes, elements, elm, _ = bench.make_inputs("tet10", scale=scale)
fd.ModelingSpace("3D")
fd.Mesh(nodes, elements, elm, name="Domain")
law = fd.constitutivelaw.ElasticIsotrop(1e5, 0.3, name="law")
fd.weakform.StressEquilibrium(law, name="wf")
a = fd.Assembly.create("wf", "Domain", elm, name="A", reuse_buffers=True, vector_on_device=True)
pb = fd.problem.Linear("A")
pb.set_X(torch.from_numpy(np.random.default_rng(0).standard_normal(3 * len(nodes)) * 1e-3).cuda())
a.update(pb, compute="all")
torch.cuda.synchronize()
out = (C.c_ulonglong * 16)()
_lib.check(_lib.load().fdk_debug_phase_clocks(out, 16, 1), "reset")
steps = 3
for _ in range(steps):
    a.assemble_global_mat("all")
torch.cuda.synchronize()
_lib.check(_lib.load().fdk_debug_phase_clocks(out, 16, 0), "read")
plan = a._plan(a._saved_bloc_structure)
tot = sum(out[:10])
print(f"tet10 {len(elements)} elements, {plan.n_clusters} clusters, caps {plan.caps}, stats {plan.stats}")
print(f"total {tot / steps / plan.n_clusters:.0f} clk/cluster = {tot / steps / len(elements):.0f} clk/element")
names = ["-", "phase 0 wait", "phase 1 (geometry)", "phase 2 (blocks)", "staging", "phase 3a heavy", "phase 3b gather+store", "D tail + end", "-", "-"]
for i in range(10):
    if out[i]:
        print(f"  [{i}] {names[i]:26s} {out[i] / steps / plan.n_clusters:8.0f} clk/cluster  {100.0 * out[i] / tot:5.1f} %")
