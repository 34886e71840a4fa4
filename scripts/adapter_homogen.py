#!/usr/bin/env python
"""fedoo.homogen.get_homogenized_stiffness of the REAL fedoo on a hex8 cell with a stiff spherical inclusion: the reference's
path (NumPy / SciPy assembly, scipy direct solver) against install(fedoo) + the device PCG as the perturbation problem's
solver (six periodic load cases on the matrix in HBM).   python scripts/adapter_homogen.py [--n 24] [--ref 1]"""
import argparse
import functools
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
from scipy.sparse.linalg import spsolve  # noqa: E402

import fedoo as fd  # noqa: E402
import fedoo_b200  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=24)
ap.add_argument("--ref", type=int, default=1)
a = ap.parse_args()


def run(tag, solver):
    fd.Assembly.delete_memory()
    fd.Problem.get_all().pop("_perturbation", None)
    fd.ModelingSpace("3D")
    t0 = time.perf_counter()
    mesh = fd.mesh.box_mesh(nx=a.n + 1, ny=a.n + 1, nz=a.n + 1, elm_type="hex8", name="Domain")
    ctr = mesh.nodes[mesh.elements].mean(axis=1)
    fd.constitutivelaw.ElasticIsotrop(np.where(np.linalg.norm(ctr - 0.5, axis=1) < 0.3, 1e6, 1e5), 0.3, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    asm = fd.Assembly.create("wf", "Domain", "hex8", name="A")
    C = np.array(fd.homogen.get_homogenized_stiffness(asm, solver=solver))
    torch.cuda.synchronize()
    return {"impl": tag, "n_elements": a.n**3, "n_dof": 3 * mesh.n_nodes + 6, "total_s": time.perf_counter() - t0,
            "C11": C[0, 0], "C12": C[0, 1], "C44": C[3, 3]}, C  # fmt: skip


out = []
if a.ref:
    out.append(run("reference: NumPy / SciPy assembly + scipy spsolve", lambda A, B, **k: spsolve(A, B)))
fedoo_b200.install(fd)
out.append(run("install(fedoo) + solver = fedoo_b200.solver.pcg", functools.partial(fedoo_b200.solver.pcg, rtol=1e-10)))
for o, _ in out:
    print(json.dumps(o))
if len(out) == 2:
    print(json.dumps({"speedup": out[0][0]["total_s"] / out[1][0]["total_s"],
                      "C_rel_diff": float(np.abs(out[0][1] - out[1][1]).max() / np.abs(out[0][1]).max())}))
