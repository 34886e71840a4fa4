"""SURVEY 8f rank 3 at scale: homogenised stiffness of a periodic hex8 cell with a stiff spherical inclusion
(n^3 elements, per-Gauss-point tangent), K assembled once by the cluster kernel and kept in HBM, six constrained
Jacobi-PCG solves on the device (fd.homogen.get_homogenized_stiffness).  Prints one JSON line.

    python scripts/homogen_bench.py --n 100
"""

import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fedoo_b200 as fd  # noqa: E402


def iso_H_gp(E_gp, nu):
    H = np.zeros((6, 6, len(E_gp)))
    lam = E_gp * nu / ((1 + nu) * (1 - 2 * nu))
    mu = 0.5 * E_gp / (1 + nu)
    for i in range(3):
        for j in range(3):
            H[i, j] = lam
        H[i, i] = lam + 2 * mu
        H[3 + i, 3 + i] = mu
    return H


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=100)
    ap.add_argument("--rtol", type=float, default=1e-8)
    ap.add_argument("--contrast", type=float, default=20.0)
    ap.add_argument("--sequential", action="store_true", help="one load case after the other instead of lockstep")
    ap.add_argument("--device-tangent", action="store_true", help="build the per-Gauss-point tangent on the device only")
    ap.add_argument("--elem", default="hex8", choices=["hex8", "tet10"], help="tet10: every cell split into 6 tetrahedra (15 GP)")
    a = ap.parse_args()
    n = a.n
    fd.ModelingSpace("3D")
    nodes, elements = fd.meshgen.box_hex8(n + 1, n + 1, n + 1)
    n_gp = 8
    if a.elem == "tet10":
        nodes, elements = fd.meshgen.tet4_to_tet10(nodes, fd.meshgen.hex8_to_tet4(elements))
        n_gp = 15
    mesh = fd.Mesh(nodes, elements, a.elem, name="Domain")
    cen = nodes[elements[:, : (4 if a.elem == "tet10" else 8)]].mean(axis=1)
    inside = np.linalg.norm(cen - 0.5, axis=1) < 0.3
    E_el = np.where(inside, 1.0e5 * a.contrast, 1.0e5)
    if a.device_tangent:  # build the (6,6,N) tangent in HBM only (21 GB at 5 M tet10 elements): 36 contiguous doubles per GP
        E_gp = torch.from_numpy(E_el).cuda().repeat(n_gp)
        lam, mu = E_gp * 0.3 / (1.3 * 0.4), 0.5 * E_gp / 1.3
        H = torch.zeros((E_gp.numel(), 36), dtype=torch.float64, device="cuda")
        for i in range(3):
            for j in range(3):
                H[:, i + 6 * j] = lam
            H[:, 7 * i] = lam + 2 * mu
            H[:, 7 * (3 + i)] = mu
        del lam, mu, E_gp
        law = fd.constitutivelaw.ElasticAnisotropic(H.reshape(-1), name="law")
    else:
        law = fd.constitutivelaw.ElasticAnisotropic(iso_H_gp(np.tile(E_el, n_gp), 0.3), name="law")
    fd.weakform.StressEquilibrium(law, name="wf")
    assemb = fd.Assembly.create("wf", "Domain", a.elem, name="A")
    fd.problem.Linear(assemb, name="main")  # initialises the assembly (sv["TangentMatrix"]) like any fedoo script
    assemb.assemble_global_mat("matrix")  # symbolic phase + first launch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    assemb.assemble_global_mat("matrix")
    torch.cuda.synchronize()
    t_asm = time.perf_counter() - t0
    t0 = time.perf_counter()
    C = fd.homogen.get_homogenized_stiffness(assemb, rtol=a.rtol, lockstep=not a.sequential)
    torch.cuda.synchronize()
    t_hom = time.perf_counter() - t0
    t0 = time.perf_counter()  # second call: the perturbation problem (node pairing, constraint map) is reused
    C = fd.homogen.get_homogenized_stiffness(assemb, rtol=a.rtol, lockstep=not a.sequential)
    torch.cuda.synchronize()
    t_again = time.perf_counter() - t0
    info = fd.Problem["_perturbation"].load_case_info
    f = float(inside.mean())
    H0, H1 = iso_H_gp(np.array([1.0e5]), 0.3)[:, :, 0], iso_H_gp(np.array([1.0e5 * a.contrast]), 0.3)[:, :, 0]
    voigt = (1 - f) * H0 + f * H1
    reuss = np.linalg.inv((1 - f) * np.linalg.inv(H0) + f * np.linalg.inv(H1))
    ev = lambda M: np.linalg.eigvalsh(0.5 * (M + M.T))  # noqa: E731
    checks = {
        "symmetric": bool(np.abs(C - C.T).max() <= 1e-6 * np.abs(C).max()),
        "cubic": a.elem != "hex8" or bool(
            np.ptp([C[0, 0], C[1, 1], C[2, 2]]) <= 1e-6 * C[0, 0]
            and np.ptp([C[0, 1], C[0, 2], C[1, 2]]) <= 1e-6 * C[0, 0]
            and np.ptp([C[3, 3], C[4, 4], C[5, 5]]) <= 1e-6 * C[0, 0]
        ),
        "reuss_le_C_le_voigt": bool(ev(voigt - C).min() >= -1e-6 * C[0, 0] and ev(C - reuss).min() >= -1e-6 * C[0, 0]),
    }
    print(json.dumps({
        "workload": f"periodic {a.elem} cell {n}^3 cells = {mesh.n_elements} elements, inclusion volume fraction {f:.4f}, contrast {a.contrast}",
        "n_dof": int(3 * mesh.n_nodes + 6), "assemble_ms": round(1e3 * t_asm, 2), "homogenisation_s": round(t_hom, 3), "homogenisation_again_s": round(t_again, 3),
        "iterations": [i["iterations"] for i in info], "relative_residual": [float(f"{i['relative_residual']:.2e}") for i in info],
        "mode": "sequential" if a.sequential else "lockstep (6 right-hand sides per K read)",
        "ms_per_iteration": round(1e3 * t_again / max(sum(i["iterations"] for i in info) if a.sequential else info[0]["iterations"], 1), 4),
        "C11_C12_C44": [round(float(C[0, 0]), 3), round(float(C[0, 1]), 3), round(float(C[3, 3]), 3)], "checks": checks,
    }))  # fmt: skip


if __name__ == "__main__":
    main()
