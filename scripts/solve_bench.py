#!/usr/bin/env python
"""Device solve path at scale (out of the headline metric, timed separately as the north star
asks): SpMV bandwidth on the assembled K and the cost of one Jacobi-PCG iteration on the hex8 box.

    python scripts/solve_bench.py [--n 200] [--iters 50] [--solve-n 64]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def setup(fd, n):
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    nodes, elements = fd.meshgen.box_hex8(n + 1, n + 1, n + 1)
    mesh = fd.Mesh(nodes, elements, "hex8", node_sets=fd.meshgen.box_node_sets(n + 1, n + 1, n + 1), name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    a = fd.Assembly.create("wf", "Domain", "hex8", name="A")
    pb = fd.problem.Linear("A")
    for var in ("DispX", "DispY", "DispZ"):
        pb.bc.add("Dirichlet", mesh.node_sets["left"], var, 0)
    pb.bc.add("Dirichlet", mesh.node_sets["right"], "DispY", -0.01)
    pb.apply_boundary_conditions()
    return mesh, a, pb


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=200)
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--solve-n", type=int, default=64)
    args = ap.parse_args()
    import torch

    import fedoo_b200 as fd

    try:
        peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        peak = 6650.0
    mesh, a, pb = setup(fd, args.n)
    a.assemble_global_mat("matrix")
    K = a.get_global_matrix()
    n = K.shape[0]
    x = torch.randn(n, dtype=torch.float64, device="cuda")
    y = torch.empty_like(x)
    free = torch.ones(n, dtype=torch.uint8, device="cuda")
    free[torch.from_numpy(pb._dirichlet[0]).cuda()] = 0
    out = {"n_elems": args.n**3, "n_dof": n, "nnz": K.nnz, "kernel": "tiled (block pattern)" if K.block is not None else "generic CSR"}
    for label, m in (("spmv", None), ("spmv_masked", free)):
        for _ in range(3):
            K.matvec(x, free_mask=m, out=y)
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(10):
            K.matvec(x, free_mask=m, out=y)
        ev1.record()
        torch.cuda.synchronize()
        ms = ev0.elapsed_time(ev1) / 10
        if K.block is not None:  # tiled kernel: one column list per node row
            byt = 8 * K.nnz + 4 * int(K.block[1].numel()) + 8 * (K.block[3] + 1) + 16 * n
        else:
            byt = K.nnz * (8 + K.indices.element_size()) + n * (8 + 8 + K.indptr.element_size())
        out[label] = {"ms": ms, "algorithmic_gb": byt / 1e9, "gbs": byt / ms / 1e6, "hbm_frac": byt / ms / 1e6 / peak}
    # fixed number of PCG iterations (rtol = 0 never stops early)
    b = torch.randn(n, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _, it, rel = K.pcg(b, free_mask=free, rtol=0.0, maxiter=args.iters, check_every=args.iters)
    torch.cuda.synchronize()
    out["pcg"] = {"iterations": it, "ms_per_iteration": (time.perf_counter() - t0) * 1e3 / max(it, 1)}
    del K, x, y, b
    # a real solve on a smaller box: iterations to rtol 1e-8
    mesh, a, pb = setup(fd, args.solve_n)
    pb.set_solver("cg", rtol=1e-8)
    t0 = time.perf_counter()
    pb.solve()
    torch.cuda.synchronize()
    out["solve"] = {"n_elems": args.solve_n**3, "n_dof": pb.n_dof, "seconds_incl_assembly": time.perf_counter() - t0, **pb.solver_info}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
