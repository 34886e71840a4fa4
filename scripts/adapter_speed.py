#!/usr/bin/env python
"""What a user of the REAL fedoo sees after fedoo_b200.install(fedoo): wall time of Assembly.update(pb, "all") on a hex8
box, the reference's own NumPy / SciPy path against the same call routed to the kernels (K handed back as a host scipy
matrix, as the reference's solvers expect).   python scripts/adapter_speed.py [--n 100]"""
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
ap.add_argument("--n", type=int, default=100)
a = ap.parse_args()


def run(tag):
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    mesh = fd.mesh.box_mesh(nx=a.n + 1, ny=a.n + 1, nz=a.n + 1, elm_type="hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    asm = fd.Assembly.create("wf", "Domain", "hex8", name="A")
    pb = fd.problem.Linear("A")
    pb.set_X(np.random.default_rng(0).standard_normal(pb.n_dof) * 1e-3)
    times = []
    for _ in range(3):
        t0 = time.perf_counter()
        asm.update(pb, compute="all")
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    K = asm.get_global_matrix()
    return {"impl": tag, "n_elements": a.n**3, "first_call_s": times[0], "warm_call_s": min(times[1:]), "nnz": int(K.nnz),
            "K_fro": float(np.sqrt((K.data**2).sum())), "D_absmax": float(np.abs(asm.get_global_vector()).max())}  # fmt: skip


ref = run("reference (NumPy / SciPy)")
fedoo_b200.install(fd)
gpu = run("fedoo_b200.install(fedoo)")
print(json.dumps(ref))
print(json.dumps(gpu))
print(json.dumps({"speedup_first_call": ref["first_call_s"] / gpu["first_call_s"], "speedup_warm_call": ref["warm_call_s"] / gpu["warm_call_s"],
                  "K_fro_rel_diff": abs(ref["K_fro"] - gpu["K_fro"]) / ref["K_fro"]}))  # fmt: skip
