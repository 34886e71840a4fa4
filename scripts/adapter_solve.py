#!/usr/bin/env python
"""A whole linear solve of the REAL fedoo (cantilever box: clamp one face, push the other), reference path against
fedoo_b200.install(fedoo) + pb.set_solver(fedoo_b200.solver.pcg):   python scripts/adapter_solve.py [--n 40] [--ref 1]"""
import argparse
import json
import os
import sys
import time
import warnings

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "_ref"))
warnings.simplefilter("ignore")
import numpy as np  # noqa: E402
import torch  # noqa: E402

import fedoo as fd  # noqa: E402
import fedoo_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=40)
ap.add_argument("--ref", type=int, default=1)
a = ap.parse_args()


def run(tag, device):
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    t0 = time.perf_counter()
    mesh = fd.mesh.box_mesh(nx=a.n + 1, ny=a.n + 1, nz=a.n + 1, elm_type="hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    fd.Assembly.create("wf", "Domain", "hex8", name="A")
    pb = fd.problem.Linear("A")
    if device:
        pb.set_solver(fedoo_b200.solver.pcg, rtol=1e-8)
    else:
        pb.set_solver("cg", rtol=1e-8)  # scipy CG with the Jacobi preconditioner (fedoo/core/base.py:482-537)
    pb.bc.add("Dirichlet", mesh.find_nodes("X", 0), "Disp", 0)
    pb.bc.add("Dirichlet", mesh.find_nodes("X", 1), "DispY", -0.01)
    pb.apply_boundary_conditions()
    t1 = time.perf_counter()
    pb.solve()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    U = pb.get_dof_solution()
    return {"impl": tag, "n_elements": a.n**3, "n_dof": int(pb.n_dof), "setup_s": t1 - t0, "solve_s (assemble + solve + state)": t2 - t1,
            "U_absmax": float(np.abs(U).max()), "U_sum": float(np.abs(U).sum())}  # fmt: skip


out = []
if a.ref:
    out.append(run("reference: NumPy / SciPy assembly + scipy cg", False))
fedoo_b200.install(fd)
out.append(run("install(fedoo) + set_solver(fedoo_b200.solver.pcg)", True))
out[-1]["pcg"] = dict(fedoo_b200.solver.info)
for o in out:
    print(json.dumps(o))
if len(out) == 2:
    k = "solve_s (assemble + solve + state)"
    print(json.dumps({"speedup": out[0][k] / out[1][k], "U_sum_rel_diff": abs(out[0]["U_sum"] - out[1]["U_sum"]) / out[0]["U_sum"]}))
