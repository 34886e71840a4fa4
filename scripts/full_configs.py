#!/usr/bin/env python
"""BASELINE.json configs [2], [3], [4] at (about) full size on one B200: assembly time of the device path and
the size-independent properties no oracle is needed for.  One JSON line per configuration;
`--out` appends them to a file (profiles/r1_full_configs.jsonl).

  thermal : tet4 HeatEquation, 6-tet split of a jittered box (150^3 cells -> 20.25 M elements): K, D, sv
  plastic : hex8 plate with a hole (8 x 100 x 100 x 50 = 4 M elements), Simcoon("EPICP")-protocol J2 update,
            K with the per-Gauss-point consistent tangent + D = -int B^T sigma (one Newton iteration's work)
  rve     : tet10 (15 Gauss points) elastic K + D on a 6-tet split box with curved edges (94^3 cells -> 4.98 M elements)

    python scripts/full_configs.py [--scale 1.0] [--only thermal,plastic,rve] [--out profiles/r1_full_configs.jsonl]
"""

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def peaks():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def csr(K):
    import torch

    crow = K.indptr.to(torch.int64)
    return torch.sparse_csr_tensor(crow, K.indices.to(torch.int64), K.data, size=K.shape)


def timed(fn, steps=3):
    import torch

    fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    return float(np.median([a.elapsed_time(b) for a, b in ev]))


def thermal(scale):
    import torch

    import fedoo_b200 as fd

    n = max(4, int(round(150 * scale)))
    nodes, hexes = fd.meshgen.box_hex8(n + 1, n + 1, n + 1)
    nodes = fd.meshgen.jitter_nodes(nodes, n + 1, n + 1, n + 1, seed=2)
    elements = fd.meshgen.hex8_to_tet4(hexes)
    del hexes
    k, c, rho, dt = 500.0, 0.5, 7800.0, 10.0 / 3.0  # tests/test_thermal3D.py:28-31,75
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    fd.Mesh(nodes, elements, "tet4", name="Domain")
    fd.constitutivelaw.ThermalProperties(k, c, rho, name="ThermalLaw")
    fd.weakform.HeatEquation("ThermalLaw")
    a = fd.Assembly.create("ThermalLaw", "Domain", name="A")
    pb = fd.problem.NonLinear("A")
    pb.dtime = dt
    nn = len(nodes)
    T0 = np.random.default_rng(3).uniform(0, 3, nn)
    T = T0 + np.random.default_rng(4).uniform(-0.5, 0.5, nn)
    pb._U = T0.copy()
    pb._dU = 0
    t0 = time.perf_counter()
    pb.initialize()
    a.set_start(pb)
    pb._dU = T - T0
    a.vector_on_device = True
    a.update(pb, "all")
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    ms = timed(lambda: a.assemble_global_mat("all"))
    import fedoo_b200.assembly as asm_mod

    ms_res = timed(lambda: a.assemble_global_mat("vector"))  # residual alone: fdk_residual_heat
    D_fast = a.global_vector.clone()
    asm_mod._RESIDUAL_KERNEL = False
    ms_res_cluster = timed(lambda: a.assemble_global_mat("vector"))  # the cluster kernel's path
    res_paths_rel = float((a.global_vector - D_fast).abs().max() / D_fast.abs().max())
    asm_mod._RESIDUAL_KERNEL = True
    a.assemble_global_mat("all")
    K = a.get_global_matrix()
    A = csr(K)
    D = a.global_vector
    one = torch.ones(nn, dtype=torch.float64, device="cuda")
    m = A @ one  # conduction rows sum to zero: what is left is the lumped capacity
    vol = 1.0
    Td, T0d = torch.from_numpy(T).cuda(), torch.from_numpy(T0).cuda()
    # the capacity term of D interpolates T - T0 consistently while K lumps it, so the exact identity is the one
    # without it: at T = T0, D = -K_cond T0 with K_cond = K - diag(m)
    D_T = D.clone()
    pb._dU = np.zeros(nn)
    a.update(pb, "all")
    D0 = a.global_vector
    resid = D0 + (A @ T0d) - m * T0d
    # and the capacity part of D sums to -(rho c / dt) int (T - T0) dV, by partition of unity (trapezoid-free check:
    # compare with the lumped estimate, second order in h)
    cap_sum = float((D_T - D0 + (A @ (Td - T0d)) - m * (Td - T0d)).sum())
    cap_lumped = -float((m * (Td - T0d)).sum())
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(nn, dtype=torch.float64, device="cuda", generator=g)
    sxy, syx = float(x @ (A @ y)), float(y @ (A @ x))
    nnz = int(K.data.numel())
    algo = 8 * nnz + 4 * 4 * len(elements) + 8 * 3 * nn + 16 * nn
    return dict(
        config="configs[2]: tet4 HeatEquation (K=500, c=0.5, rho=7800, dt=10/3), 6-tet split of a jittered box",
        n_elems=len(elements), n_nodes=nn, nnz=nnz, first_call_s=first, ms_per_assembly=ms,
        ms_residual_only=ms_res, ms_residual_only_cluster_kernel=ms_res_cluster,
        melem_per_s=len(elements) / ms / 1e3, algorithmic_gb=algo / 1e9, hbm_frac=algo / (ms * 1e-3) / 1e9 / peaks(),
        checks=dict(
            residual_paths_rel=res_paths_rel,
            lumped_capacity_total_rel=abs(float(m.sum()) - rho * c / dt * vol) / (rho * c / dt * vol),
            capacity_positive=bool((m > 0).all()),
            D_conduction_balance_rel=float(resid.abs().max() / D0.abs().max()),
            D_capacity_sum_vs_lumped_rel=abs(cap_sum - cap_lumped) / abs(cap_lumped),
            symmetry_rel=abs(sxy - syx) / abs(sxy),
        ),
    )  # fmt: skip


def plastic(scale):
    import torch

    import fedoo_b200 as fd

    nr = nt = max(3, int(round(100 * scale ** (1 / 3))) + 1)
    layers = max(2, int(round(50 * scale ** (1 / 3))))
    n2, quads = fd.meshgen.hole_plate_quad4(nr, nt, 100.0, 100.0, 20.0)
    nodes, elements = fd.meshgen.extrude_quad4_to_hex8(n2, quads, 25.0, layers)
    props = [200e3, 0.3, 1e-5, 300.0, 1000.0, 0.3]  # examples/plasticity/plastic_bending_3D.py:27-58
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    fd.Mesh(nodes, elements, "hex8", name="Domain")
    law = fd.constitutivelaw.Simcoon("EPICP", props, name="law")
    fd.weakform.StressEquilibrium(law, name="wf")
    a = fd.Assembly.create("wf", "Domain", "hex8", name="A")
    pb = fd.problem.Linear("A")
    nn = len(nodes)
    # tension along x growing with y: the upper half of the plate yields (sigma_Y / E = 1.5e-3)
    e0 = 2.5e-3 * (0.5 + nodes[:, 1] / 100.0)
    U = np.concatenate([e0 * nodes[:, 0], -0.3 * e0 * nodes[:, 1], -0.3 * e0 * nodes[:, 2]])
    U += np.random.default_rng(0).standard_normal(3 * nn) * 1e-5
    pb.set_X(torch.from_numpy(U).cuda())  # the iterate lives in HBM between Newton iterations (no 100 MB pageable H2D per update)
    a.vector_on_device = True
    t0 = time.perf_counter()
    a.update(pb, compute="all")
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    ms_update = timed(lambda: a.update(pb, compute="all"))  # strain + J2 update + K (per-GP tangent) + B^T sigma
    ms_asm = timed(lambda: a.assemble_global_mat("all"))    # K + D alone, state already updated
    K = a.get_global_matrix()
    A = csr(K)
    D = a.global_vector
    scale_k = float(K.data.abs().max())
    t = torch.zeros(3 * nn, dtype=torch.float64, device="cuda")
    t[:nn] = 1.0
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(3 * nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(3 * nn, dtype=torch.float64, device="cuda", generator=g)
    sxy, syx = float(x @ (A @ y)), float(y @ (A @ x))
    sv = a.sv["Statev"]
    p = sv[:, 1] if sv.shape[1] == 8 else sv[1]
    nnz = int(K.data.numel())
    n_gp = 8 * len(elements)
    # J2 fused update + K + R: SURVEY 8d (statev in/out, stress out, tangent written then read, K, conn, coords, U, D)
    algo = 8 * nnz + 32 * len(elements) + 24 * nn + 48 * nn + n_gp * 8 * (8 + 8 + 6 + 6 + 36 + 36)
    return dict(
        config="configs[3]: hex8 plate with a hole, J2 (EPICP props [200e3, 0.3, 1e-5, 300, 1000, 0.3]): Gauss-point "
        "radial return + consistent tangent, K with the per-GP tangent + D = -int B^T sigma",
        n_elems=len(elements), n_nodes=nn, nnz=nnz, first_call_s=first, ms_per_update=ms_update, ms_per_assembly=ms_asm,
        melem_per_s=len(elements) / ms_update / 1e3, algorithmic_gb=algo / 1e9, hbm_frac=algo / (ms_update * 1e-3) / 1e9 / peaks(),
        checks=dict(
            yielded_fraction=float((p > 0).double().mean()),
            rigid_translation_rel=float((A @ t).abs().max() / scale_k),
            internal_force_sum_rel=float(torch.stack([D[v * nn:(v + 1) * nn].sum() for v in range(3)]).abs().max() / D.abs().max()),
            symmetry_rel=abs(sxy - syx) / abs(sxy),
        ),
    )  # fmt: skip


def rve(scale):
    import torch

    import fedoo_b200 as fd

    n = max(3, int(round(94 * scale ** (1 / 3))))
    nodes, hexes = fd.meshgen.box_hex8(n + 1, n + 1, n + 1)
    t4 = fd.meshgen.hex8_to_tet4(hexes)
    del hexes
    nodes, elements = fd.meshgen.tet4_to_tet10(nodes, t4, bulge=0.02)
    del t4
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    fd.Mesh(nodes, elements, "tet10", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(1e5, 0.3, name="law")  # examples/homogenization/periodic_homogenization.py:37
    fd.weakform.StressEquilibrium("law", name="wf")
    a = fd.Assembly.create("wf", "Domain", "tet10", name="A")
    pb = fd.problem.Linear("A")
    nn = len(nodes)
    U = np.random.default_rng(0).standard_normal(3 * nn) * 1e-3
    pb.set_X(U)
    a.vector_on_device = True
    t0 = time.perf_counter()
    a.update(pb, compute="all")
    torch.cuda.synchronize()
    first = time.perf_counter() - t0
    ms = timed(lambda: a.assemble_global_mat("all"))
    K = a.get_global_matrix()
    A = csr(K)
    D = a.global_vector
    scale_k = float(K.data.abs().max())
    Ud = torch.from_numpy(U).cuda()
    t = torch.zeros(3 * nn, dtype=torch.float64, device="cuda")
    t[nn:2 * nn] = 1.0
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(3 * nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(3 * nn, dtype=torch.float64, device="cuda", generator=g)
    sxy, syx = float(x @ (A @ y)), float(y @ (A @ x))
    nnz = int(K.data.numel())
    algo = 8 * nnz + 4 * 10 * len(elements) + 24 * nn + 48 * nn
    return dict(
        config="configs[4]: tet10 (15 Gauss points) ElasticIsotrop E=1e5 nu=0.3, K + D, 6-tet split box with curved edges "
        "(one load case; the 6 load cases of the homogenisation reuse K)",
        n_elems=len(elements), n_nodes=nn, nnz=nnz, first_call_s=first, ms_per_assembly=ms,
        melem_per_s=len(elements) / ms / 1e3, algorithmic_gb=algo / 1e9, hbm_frac=algo / (ms * 1e-3) / 1e9 / peaks(),
        checks=dict(
            D_plus_KU_rel=float((D + A @ Ud).abs().max() / D.abs().max()),
            rigid_translation_rel=float((A @ t).abs().max() / scale_k),
            symmetry_rel=abs(sxy - syx) / abs(sxy),
        ),
    )  # fmt: skip


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the full element count")
    ap.add_argument("--only", default="thermal,plastic,rve")
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch

    torch.cuda.set_device(0)
    for name in args.only.split(","):
        fn = {"thermal": thermal, "plastic": plastic, "rve": rve}[name]
        t0 = time.perf_counter()
        try:
            line = fn(args.scale if name != "thermal" else args.scale ** (1 / 3))
        except Exception as e:  # report and go on with the next configuration
            line = dict(config=name, error=f"{type(e).__name__}: {e}")
        if "clk" in os.environ.get("FDK_LIB", ""):  # diagnostic build: SM cycles per phase of the cluster kernel
            import ctypes as C

            from fedoo_b200 import _lib

            out = (C.c_ulonglong * 16)()
            _lib.check(_lib.load().fdk_debug_phase_clocks(out, 16, 1), "phase clocks")
            tot = max(1, sum(out[:10]))
            line["phase_clock_shares"] = {str(i): round(out[i] / tot, 3) for i in range(10) if out[i]}
        line["wall_s"] = time.perf_counter() - t0
        line["scale"] = args.scale
        print(json.dumps(line), flush=True)
        if args.out:
            with open(args.out, "a") as f:
                f.write(json.dumps(line) + "\n")
        import fedoo_b200 as fd

        fd.Assembly.delete_memory()
        import gc

        gc.collect()
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
