#!/usr/bin/env python
"""Benchmarks of the assembly hot path.  Default = the headline: global K + residual assembly, hex8 linear elastic,
200^3-element box (BASELINE.json configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--config hex8|heat_tet4|j2_plate|tet10] [--impl reference]
                    [--n 200] [--scale 1.0] [--jitter 0|1] [--check]

One "step" = one pass of the hot path over the whole mesh.
  hex8      (configs[1]) the fused cluster kernel writes every CSR value of K and every entry of D = -int B^T sigma(U).
            N > 1 (torchrun, one rank per GPU): nodes in z-slabs, each rank assembles the rows it owns (strong scaling:
            the global mesh is fixed); the residual exchange is fused into the kernel (NVLink stores) or NCCL all-gather.
  heat_tet4 (configs[2]) tet4 HeatEquation K (conduction + lumped capacity) + residual, 6-tet split of a jittered box.
  j2_plate  (configs[3]) hex8 plate with a hole: strain, J2 radial return + tangent at every Gauss point, K with the
            per-Gauss-point tangent, D = -int B^T sigma (one Newton iteration's work).
  tet10     (configs[4]) tet10 (15 Gauss points) elastic K + D (one load case of the homogenisation).
            heat_tet4 / j2_plate / tet10 with N > 1: nodes partitioned by recursive coordinate bisection, owner-computes
            rows, the owned slices of D all-gathered over NCCL.
--impl reference: the UNMODIFIED reference (oracle/_ref, copied by oracle/make_ref.py) timed on the host cores on a
bounded sample of the same workload (the J2 law lives in simcoon, absent: that config times the NumPy port instead).

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for the definitions.
"""

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRICS = {
    "hex8": "Global K+R assembly Melem/s (hex8, 8M elems)",
    "heat_tet4": "Global K+R assembly Melem/s (tet4 HeatEquation, 20M elems)",
    "j2_plate": "J2 update + K+R assembly Melem/s (hex8 plate with hole, 4M elems)",
    "tet10": "Global K+R assembly Melem/s (tet10 elastic, 5M elems)",
}
CPU_SAMPLE_EDGE = {"hex8": 40, "heat_tet4": 24, "j2_plate": 12, "tet10": 16}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", default="hex8", choices=list(METRICS))
    ap.add_argument("--n", "--edge", dest="n", type=int, default=200,
                    help="hex8: elements per box edge (200 -> 8 M hex8); --edge: the spelling torchrun does not mistake for its own")
    ap.add_argument("--scale", type=float, default=1.0, help="other configs: fraction of the full element count")
    ap.add_argument("--jitter", type=int, default=0, help="hex8: 1 = displace interior nodes (no identical elements)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-n", type=int, default=0, help="edge of the bounded CPU sample (0: per-config default)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="copy", choices=["copy", "peer", "nccl"],
                    help="N > 1, residual exchange: copy = one coalesced copy of the owned slices to the NVLink multicast "
                    "address after the kernel; peer = stores fused into the assembly kernel; nccl = pack + all-gather + unpack")
    ap.add_argument("--check", action="store_true", help="verify size-independent properties of the assembled K, D")
    return ap.parse_args()


def workload_label(args):
    if args.config == "hex8":
        return (f"hex8 box {args.n}x{args.n}x{args.n} ({args.n**3} elements), ElasticIsotrop E=200e3 nu=0.3, "
                "K + residual assembly (configs[1])")  # fmt: skip
    return {
        "heat_tet4": "tet4 HeatEquation (K=500, c=0.5, rho=7800, dt=10/3), 6-tet split of a jittered 150^3-cell box "
        "(20.25 M elements), K + residual (configs[2])",
        "j2_plate": "hex8 plate with a hole 8 x 100 x 100 x 50 (4 M elements), J2 EPICP [200e3, 0.3, 1e-5, 300, 1000, 0.3]: "
        "strain + radial return + tangent + K (per-GP tangent) + B^T sigma (configs[3])",
        "tet10": "tet10 (15 GP) ElasticIsotrop E=1e5 nu=0.3, 6-tet split of a 94^3-cell box with curved edges "
        "(4.98 M elements), K + residual (configs[4])",
    }[args.config] + ("" if args.scale == 1.0 else f" [scale {args.scale}]")


# ----------------------------------------------------------------------------------------------
# synthetic inputs (host NumPy; shared by both arms)
# ----------------------------------------------------------------------------------------------
def make_inputs(config, edge=None, scale=1.0):
    """(nodes, elements, elm_type, extras) of a configuration; ``edge`` overrides the full-size cell count."""
    from fedoo_b200 import meshgen

    if config == "hex8":
        n = edge
        nodes, elements = meshgen.box_hex8(n + 1, n + 1, n + 1)
        return nodes, elements, "hex8", {}
    if config == "heat_tet4":
        n = edge or max(4, int(round(150 * scale ** (1 / 3))))
        nodes, hexes = meshgen.box_hex8(n + 1, n + 1, n + 1)
        nodes = meshgen.jitter_nodes(nodes, n + 1, n + 1, n + 1, seed=2)
        return nodes, meshgen.hex8_to_tet4(hexes), "tet4", {}
    if config == "j2_plate":
        if edge:
            nr, layers = edge + 1, max(2, edge // 2)
        else:
            nr, layers = max(3, int(round(100 * scale ** (1 / 3))) + 1), max(2, int(round(50 * scale ** (1 / 3))))
        n2, quads = meshgen.hole_plate_quad4(nr, nr, 100.0, 100.0, 20.0)
        nodes, elements = meshgen.extrude_quad4_to_hex8(n2, quads, 25.0, layers)
        return nodes, elements, "hex8", {}
    if config == "tet10":
        n = edge or max(3, int(round(94 * scale ** (1 / 3))))
        nodes, hexes = meshgen.box_hex8(n + 1, n + 1, n + 1)
        t4 = meshgen.hex8_to_tet4(hexes)
        nodes, elements = meshgen.tet4_to_tet10(nodes, t4, bulge=0.02)
        return nodes, elements, "tet10", {}
    raise ValueError(config)


def dof_vector(config, nodes):
    nn = len(nodes)
    if config == "heat_tet4":
        T0 = np.random.default_rng(3).uniform(0, 3, nn)
        return T0, T0 + np.random.default_rng(4).uniform(-0.5, 0.5, nn)
    if config == "j2_plate":
        e0 = 2.5e-3 * (0.5 + nodes[:, 1] / 100.0)  # tension along x growing with y: the upper half yields
        U = np.concatenate([e0 * nodes[:, 0], -0.3 * e0 * nodes[:, 1], -0.3 * e0 * nodes[:, 2]])
        return None, U + np.random.default_rng(0).standard_normal(3 * nn) * 1e-5
    return None, np.random.default_rng(0).standard_normal(3 * nn) * 1e-3


J2_PROPS = [200e3, 0.3, 1e-5, 300.0, 1000.0, 0.3]  # examples/plasticity/plastic_bending_3D.py:27-58


# ----------------------------------------------------------------------------------------------
# CPU arm: the reference itself (oracle/_ref) on the host cores
# ----------------------------------------------------------------------------------------------
def reference_step_factory(config, edge):
    """Returns (step(), n_elements, kind, description): one steady-state pass of the reference's own path
    (Assembly.update(pb, "all") on a warm assembly: operators and CSR structure cached, as its 2nd call)."""
    nodes, elements, elm, _ = make_inputs(config, edge=edge)
    T0, U = dof_vector(config, nodes)
    from oracle import make_ref

    have_ref = make_ref.ref_path() is not None
    if config == "j2_plate" or not have_ref:
        return _port_step_factory(config, nodes, elements, elm, T0, U, have_ref)
    import warnings

    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        fedoo = make_ref.import_fedoo()
    fedoo.Assembly.delete_memory()
    fedoo.ModelingSpace("3D")
    fedoo.Mesh(nodes, elements.astype(np.int64), elm, name="Domain")
    if config == "heat_tet4":
        fedoo.constitutivelaw.ThermalProperties(500.0, 0.5, 7800.0, name="ThermalLaw")
        fedoo.weakform.HeatEquation("ThermalLaw")
        a = fedoo.Assembly.create("ThermalLaw", "Domain", name="A")
        pb = fedoo.problem.NonLinear("A")
        pb.dtime = 10.0 / 3.0
        pb._U, pb._dU = T0.copy(), 0
        pb.initialize()
        a.set_start(pb)
        pb._dU = U - T0
    else:
        E = 200e3 if config == "hex8" else 1e5
        fedoo.constitutivelaw.ElasticIsotrop(E, 0.3, name="law")
        fedoo.weakform.StressEquilibrium("law", name="wf")
        a = fedoo.Assembly.create("wf", "Domain", elm, name="A")
        pb = fedoo.problem.Linear("A")
        pb.set_X(U)

    def step():
        a.update(pb, compute="all")
        return a.global_matrix, a.global_vector

    import scipy

    desc = (f"UNMODIFIED reference (fedoo {fedoo.__version__}, oracle/_ref): Assembly.update(pb, 'all') = state update + "
            f"assemble_global_mat on a warm assembly; numpy {np.__version__}, scipy {scipy.__version__}")  # fmt: skip
    return step, len(elements), "reference", desc


def _port_step_factory(config, nodes, elements, elm, T0, U, have_ref):
    """NumPy port of the reference path (oracle/fedoo_oracle.py): used for the J2 configuration (the reference's law is
    simcoon's, absent) and wherever oracle/_ref has not been built."""
    from oracle import fedoo_oracle as fo

    nn = len(nodes)
    G, wdet = fo.geometry(nodes, elements, elm)
    if config == "heat_tet4":
        pat = fo.Pattern(elements, nn, 1)
        pat.indices

        def step():
            K = fo.assemble_heat(nodes, elements, elm, 500.0, 7800.0 * 0.5, 10.0 / 3.0, pattern=pat, geom=(G, wdet))
            D = fo.residual_heat(G, wdet, elements, elm, 500.0, 7800.0 * 0.5, 10.0 / 3.0, U, T0, nn)
            return K, D
    else:
        pat = fo.Pattern(elements, nn, 3)
        pat.indices
        H = fo.elastic_isotropic_H(200e3 if config != "tet10" else 1e5, 0.3)
        sv0 = np.zeros((8, wdet.size))

        def step():
            eps = fo.strain_gp(G, elements, U, nn, 3)
            if config == "j2_plate":
                sig, _sv, Ct = fo.j2_radial_return(eps, sv0, J2_PROPS)
            else:
                sig, Ct = fo.stress_gp(H, eps), H
            data = fo.stiffness_blocks(G, wdet, Ct, 3)
            blocks = [[pat.block_values(data[a][b]) for b in range(3)] for a in range(3)]
            K = pat.csr(pat.tile_values(blocks))
            D = fo.residual(G, wdet, elements, sig, nn, 3)
            return K, D

    why = ("the reference's J2 law computes in simcoon (absent): NumPy port of the same path (oracle/fedoo_oracle.py)"
           if have_ref else "oracle/_ref not built here: NumPy port of the reference path (oracle/fedoo_oracle.py)")  # fmt: skip
    return step, len(elements), "port", why + f"; numpy {np.__version__}"


def cpu_baseline(config, edge, steps, warmup):
    step, n_el, kind, desc = reference_step_factory(config, edge)
    for _ in range(max(1, warmup)):
        step()
    c0, t0 = time.process_time(), time.perf_counter()
    for _ in range(steps):
        step()
    wall = time.perf_counter() - t0
    cpu = time.process_time() - c0
    return dict(
        value=n_el * steps / wall / 1e6,
        unit="Melem/s",
        cores=max(1, int(round(cpu / wall))),
        kind=kind,
        sample=f"{config}: {n_el} elements (edge {edge}), {steps} steady-state steps after {max(1, warmup)} warm-up; {desc}; "
        f"host logical cpus {os.cpu_count()}, OMP/OPENBLAS_NUM_THREADS="
        f"{os.environ.get('OMP_NUM_THREADS', 'unset')}/{os.environ.get('OPENBLAS_NUM_THREADS', 'unset')} "
        "(the path is single-thread bound: scipy.sparse products and Python term loops)",
        ms_per_step=wall / steps * 1e3,
        n_elements=n_el,
    )


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    edge = args.cpu_n or CPU_SAMPLE_EDGE[args.config]
    cb = cpu_baseline(args.config, edge, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": METRICS[args.config],
        "value": cb["value"],
        "unit": "Melem/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": max(1, args.warmup),
        "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong" if args.config == "hex8" else "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload_label(args),
            "sample": f"bounded CPU sample actually run: {cb['n_elements']} elements of the same mesh family (edge {edge}); "
            "throughput per element is flat or falling with size for the reference (SURVEY 6: 0.057 Melem/s at 64 k, "
            "0.046 at 1 M)",
            "jitter": bool(args.jitter),
        },
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------------
# clocks sampler (pynvml), DURING the timed region
# ----------------------------------------------------------------------------------------------
class Clocks:
    def __init__(self, index):
        self.samples, self.reasons, self.stop = [], set(), False
        try:
            import pynvml

            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None
            self.max = None
        self.t = threading.Thread(target=self.loop, daemon=True)

    def loop(self):
        nv = self.nv
        names = {
            "hw_slowdown": 0x8, "sw_power_cap": 0x4, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40,
            "hw_power_brake": 0x80, "sync_boost": 0x10,
        }  # fmt: skip
        while not self.stop:
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:
                pass
            time.sleep(0.02)

    def __enter__(self):
        if self.nv:
            self.t.start()
        return self

    def __exit__(self, *a):
        self.stop = True
        if self.nv:
            self.t.join()

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.max, "reasons": []}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.max, "reasons": sorted(self.reasons),
                "samples": len(self.samples)}  # fmt: skip


def hbm_peak():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return float(p["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy figure)"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


def measured_traffic(key):
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    try:
        return json.load(open(tp)).get(key)
    except Exception:
        return None


def time_loop(torch, fn, steps):
    """CUDA-event time of ``steps`` calls of fn on the current stream: (total ms, mean per-call ms)."""
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for a, b in ev:
        a.record()
        fn()
        b.record()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1), float(np.mean([a.elapsed_time(b) for a, b in ev]))


# ----------------------------------------------------------------------------------------------
# headline: hex8, 1..N GPUs
# ----------------------------------------------------------------------------------------------
def run_hex8(args):
    import torch
    import torch.distributed as dist

    import fedoo_b200 as fd
    from fedoo_b200 import dist as fdist

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n = args.n
    nn = n + 1
    n_elems_global = n**3
    n_nodes_global = nn**3

    # ---- rank-local mesh (never builds the global 8 M-element arrays on N > 1) ----
    loc = fdist.box_local_slab(n, rank, world, jitter=bool(args.jitter))
    fd.ModelingSpace("3D")
    fd.Mesh(loc.nodes, loc.elements, "hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="ElasticLaw")
    fd.weakform.StressEquilibrium("ElasticLaw", name="weakform")
    asm = fd.Assembly.create("weakform", "Domain", "hex8", name="Assembling", reuse_buffers=True,
                             owned_nodes=(loc.owned if world > 1 else None))  # fmt: skip
    pb = fd.problem.Linear("Assembling")
    n_loc = len(loc.nodes)
    # U = default_rng(0).standard_normal(n_dof) * 1e-3 of the GLOBAL problem, restricted to the local nodes
    Ug = np.random.default_rng(0).standard_normal(3 * n_nodes_global) * 1e-3
    U_host = torch.empty(3 * n_loc, dtype=torch.float64, pin_memory=True)
    for v in range(3):
        U_host[v * n_loc : (v + 1) * n_loc] = torch.from_numpy(Ug[v * n_nodes_global + loc.node_gid])
    del Ug
    pb.set_X(U_host)

    t0 = time.perf_counter()
    asm.update(pb, compute="all")  # builds pattern + plan (one-time symbolic) and assembles once
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    entry = asm._saved_bloc_structure
    plan, pattern = asm._plan(entry), entry["pattern"]
    # N > 1: the exchange of the residual.  Default: fused into the assembly kernel (stores over NVLink into a
    # symmetric, multicast-mapped, double-buffered global vector, fedoo_b200.dist.PeerVector); --exchange nccl, or a
    # failed symmetric-memory rendezvous: pack + NCCL all-gather + unpack (fedoo_b200.dist.VectorExchange)
    exch = peer = D_nccl = None
    if world > 1:
        if args.exchange in ("peer", "copy"):
            try:
                peer = fdist.PeerVector(loc, 3, mode=("fused" if args.exchange == "peer" else "copy"))
                asm.peer_vector = peer
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); NCCL all-gather", file=sys.stderr)
                peer = None
        if peer is None:
            exch = fdist.VectorExchange(loc, 3)
            D_nccl = torch.zeros(3 * n_nodes_global, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    asm.vector_on_device = True

    def step_device():
        asm.assemble_global_mat("all")
        if exch is not None:
            exch.allgather(asm.global_vector, D_nccl)

    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    torch.cuda.profiler.start()  # ncu --profile-from-start off: the launch list of the timed region only
    with Clocks(local_rank) as clk:
        ev0.record()
        for i in range(args.steps):
            kev[i][0].record()
            asm.assemble_global_mat("all")
            kev[i][1].record()
            if exch is not None:
                exch.allgather(asm.global_vector, D_nccl)
        ev1.record()
        barrier()
    torch.cuda.profiler.stop()
    ms_total = ev0.elapsed_time(ev1)
    ms_kernel = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    t = torch.tensor([ms_total, ms_kernel], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_kernel_max = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    value = n_elems_global / (ms_step * 1e-3) / 1e6

    # ---- end-to-end through the public API with HOST buffers: `e2e` ----
    # every step: pinned dof vector -> HBM, fused K + R kernel, residual -> pinned host.  The copies ride two copy
    # streams with rotating buffers (Assembly(async_copies=True)): step i + 1's H2D and step i's D2H overlap the kernels;
    # the result of step i is read (synchronised) after step i + 1 has been queued.
    asm.vector_on_device = False
    asm.async_copies = True

    def e2e_loop(k):
        prev = out = None
        for _ in range(k):
            pb.set_X(U_host)
            asm.update(pb, compute="all")
            cur = asm.get_global_vector()
            if prev is not None:
                out = prev.result()
            prev = cur
        return prev.result() if prev is not None else out

    e2e_loop(3)
    barrier()
    t0 = time.perf_counter()
    Dh = e2e_loop(args.steps)
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_elems_global / (float(t[0]) / args.steps) / 1e6
    h2d = U_host.numel() * 8
    d2h = int(np.asarray(Dh).size) * 8
    asm.async_copies = False

    # ---- roofline of the dominant kernel (this rank's share) ----
    n_own_nodes = plan.n_owned
    n_own_elems = n_elems_global / world  # owner-computes: a rank's algorithmic share of the elements
    nnz_local = 9 * int(plan.t["cl_slot_ptr"][-1])
    algo_bytes = 8 * nnz_local + 4 * 8 * n_own_elems + 8 * 3 * n_own_nodes + 16 * 3 * n_own_nodes
    peak, peak_src = hbm_peak()
    achieved = algo_bytes / (ms_kernel_max * 1e-3) / 1e9

    from fedoo_b200 import _lib as _fdk_lib

    kernel_name = ("fdk::k_assemble_iso<Hex8, 1024 threads, 4 per incidence>" if _fdk_lib.get_option("iso4")
                   else "fdk::k_assemble<Hex8, PHYS_ISO>")  # fmt: skip
    checks = None
    if args.check and world == 1:
        checks = property_checks(asm, pb, U_host, n)
    elif args.check:
        checks = distributed_checks(asm, pb, loc, peer, exch, D_nccl, n, world, rank, bool(args.jitter))

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline("hex8", args.cpu_n or CPU_SAMPLE_EDGE["hex8"], 3, 1)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": METRICS["hex8"],
            "value": value,
            "unit": "Melem/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": max(args.warmup, 3),
            "ms_per_step": ms_step,
            "higher_is_better": True,
            "scaling": "strong",
            "vs_baseline": None,
            "dtype": "f64",
            "data": "synthetic",
            "config": {
                "workload": workload_label(args),
                "jitter": bool(args.jitter),
                "l2": "working set (>= 16 GB written per step at n=200) far exceeds the 126 MB L2; no flush needed",
                "partition": f"z-slabs of nodes over {world} GPU(s), owner-computes rows, residual exchange: "
                + ("none (1 GPU)" if world == 1 else
                   ((("fused into the assembly kernel: stores" if peer.mode == "fused" else
                      "one coalesced segment copy of the owned slices after the kernel")
                     + " over NVLink into a symmetric, double-buffered global vector ("
                     + ("NVSwitch multicast" if peer.multicast else "peer addresses") + ") + device barrier")
                    if peer is not None else "pack + NCCL all-gather of D + unpack")),
                "nnz": 9 * pattern.blk_nnz if world == 1 else None,
                "nnz_per_s": (9 * (3 * n + 1) ** 3) / (ms_step * 1e-3),
                "clusters": plan.n_clusters,
                "geometry_redundancy": plan.stats["redundancy"] if world == 1 else None,
                "plan_metadata_bytes": plan.metadata_bytes(),
                "first_call_s": t_first,
            },
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": measured_traffic(f"n{n}_g{world}"),
                "peak_source": peak_src,
                "kernel": kernel_name,
                "binding_resource": "shared FP64 / shared-memory issue path of the SM (ncu: LSU data pipe + FP64 pipe ~ 100 % "
                "busy, DRAM < 20 %); see DESIGN.md section 7",
                "kernel_ms": ms_kernel_max,
                "algorithmic_bytes_per_launch": algo_bytes,
            },
            "cpu_baseline": cb,
            "e2e": {
                "value": e2e_value,
                "unit": "Melem/s",
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "note": "every step: pinned host U -> HBM, fused K+R kernel, residual -> pinned host, through "
                "Problem.set_X / Assembly.update / get_global_vector; copies pipelined on two copy streams with rotating "
                "buffers (async_copies); K stays in HBM (DeviceCSR, materialised to scipy only on demand)",
            },
            # the assembly kernel, plus pack / unpack of the residual exchange (NCCL's own kernels not counted)
            "gpu_launches": args.steps * ((1 if peer is None or peer.mode == "fused" else 2) if exch is None
                                          else (3 if exch.seg_pack is not None else 6)),
            "clocks": clk.summary(),
        }
        if checks is not None:
            line["checks"] = checks
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def property_checks(asm, pb, U_host, n):
    """Size-independent properties at full size, on the device (no oracle can run at 8 M elements):
    D == -K U, K t = 0 for rigid translations, K symmetric (via x^T K y == y^T K x)."""
    import torch

    asm.vector_on_device = False
    asm.assemble_global_mat("all")
    K = asm.get_global_matrix()
    crow = K.indptr.to(torch.int64) if K.indptr.dtype != torch.int64 else K.indptr
    A = torch.sparse_csr_tensor(crow, K.indices.to(torch.int64), K.data, size=K.shape)
    U = U_host.cuda()
    D = torch.from_numpy(np.asarray(asm.get_global_vector())).cuda()
    scale = float(K.data.abs().max())
    KU = A @ U
    out = {"D_plus_KU_rel": float((D + KU).abs().max() / D.abs().max())}
    nn = (n + 1) ** 3
    t = torch.zeros(3 * nn, dtype=torch.float64, device="cuda")
    t[:nn] = 1.0
    out["rigid_translation_rel"] = float((A @ t).abs().max() / scale)
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(3 * nn, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(3 * nn, dtype=torch.float64, device="cuda", generator=g)
    a, b = float(x @ (A @ y)), float(y @ (A @ x))
    out["symmetry_rel"] = abs(a - b) / max(abs(a), 1e-300)
    return out


def distributed_checks(asm, pb, loc, peer, exch, D_nccl, n, world, rank, jitter):
    """N > 1 against the ONE-GPU result, on a mesh small enough to redo on every rank: each rank reassembles the GLOBAL
    problem alone (no partition) and compares (a) the gathered residual it received with the single-GPU D, entry by
    entry, and (b) its owned K rows with the same rows of the single-GPU K.  Also at full size: the gathered residual is
    identical on all ranks and self-equilibrated."""
    import torch
    import torch.distributed as dist

    import fedoo_b200 as fd
    from fedoo_b200 import dist as fdist

    nn = (n + 1) ** 3
    asm.vector_on_device = True
    asm.assemble_global_mat("all")
    if exch is not None:
        exch.allgather(asm.global_vector, D_nccl)
    D_global = peer.tensor if peer is not None else D_nccl
    torch.cuda.synchronize()
    s3 = torch.stack([D_global[v * nn : (v + 1) * nn].sum() for v in range(3)])
    spread = torch.stack([D_global.sum(), -D_global.sum()])
    dist.all_reduce(spread, op=dist.ReduceOp.MAX)
    out = {
        "D_sum_rel": float(s3.abs().max() / D_global.abs().max()),
        "D_identical_on_all_ranks": bool(float(spread[0] + spread[1]) == 0.0),
        "D_nonzero_fraction": float((D_global != 0).double().mean()),
    }
    if n <= 64:
        D_multi = D_global.clone()
        K_loc = asm.get_global_matrix()
        kd, kp = K_loc.data.clone(), K_loc.indptr.clone()
        whole = fdist.box_local_slab(n, 0, 1, jitter=jitter)
        fd.Assembly.delete_memory()
        fd.ModelingSpace("3D")
        fd.Mesh(whole.nodes, whole.elements, "hex8", name="Whole")
        a1 = fd.Assembly.create("weakform", "Whole", "hex8", name="One", vector_on_device=True)
        pb1 = fd.problem.Linear("One")
        Ug = np.random.default_rng(0).standard_normal(3 * nn) * 1e-3
        pb1.set_X(Ug)
        a1.update(pb1, compute="all")
        D1 = a1.global_vector
        out["D_vs_single_gpu_rel"] = float((D_multi - D1).abs().max() / D1.abs().max())
        # owned rows of K: local row v * n_loc + l  <->  global row v * nn + gid[l]; same column count, same values
        K1 = a1.get_global_matrix()
        n_loc = len(loc.nodes)
        own = torch.from_numpy(np.flatnonzero(loc.owned)).cuda()
        gid = torch.from_numpy(np.asarray(loc.node_gid)).cuda()[own]
        worst = 0.0
        for v in range(3):
            for l, g in zip(own[:: max(1, len(own) // 200)].tolist(), gid[:: max(1, len(own) // 200)].tolist()):
                r0, r1 = int(kp[v * n_loc + l]), int(kp[v * n_loc + l + 1])
                q0, q1 = int(K1.indptr[v * nn + g]), int(K1.indptr[v * nn + g + 1])
                if r1 - r0 != q1 - q0:
                    worst = float("inf")
                    break
                worst = max(worst, float((kd[r0:r1] - K1.data[q0:q1]).abs().max()))
        out["K_owned_rows_vs_single_gpu_abs_over_max"] = worst / float(K1.data.abs().max())
    t = torch.tensor([v for v in out.values() if isinstance(v, float)], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    keys = [k for k, v in out.items() if isinstance(v, float)]
    out.update({k: float(x) for k, x in zip(keys, t)})
    return out


# ----------------------------------------------------------------------------------------------
# configs [2]-[4]: one GPU, or (torchrun) the unstructured partition of fedoo_b200.dist over N GPUs
# ----------------------------------------------------------------------------------------------
def run_other(args):
    """configs [2]-[4].  Under torchrun the nodes are partitioned by recursive coordinate bisection, every rank
    assembles the complete rows of the nodes it owns from its local mesh (one halo layer of elements, whose Gauss-point
    state it updates itself; no matrix exchange) and the owned slices of D are all-gathered over NCCL (fedoo_b200.dist:
    partition_rcb / extract_local / VectorExchange) -- strong scaling, the global mesh is fixed."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    cfg = args.config
    sharded = world > 1
    import torch
    import torch.distributed as dist

    import fedoo_b200 as fd
    from fedoo_b200 import dist as fdist

    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    if sharded:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    nodes, elements, elm, _ = make_inputs(cfg, scale=args.scale)
    T0, U = dof_vector(cfg, nodes)
    nn_global, n_el = len(nodes), len(elements)
    nvar = 1 if cfg == "heat_tet4" else 3
    loc = exch = D_global = None
    part_stats = None
    glob = None
    if sharded:
        if args.check and n_el <= 400_000:
            glob = (nodes, elements, U, T0)  # small enough to redo on one GPU for the comparison
        part = fdist.partition_rcb(nodes, world)  # deterministic: every rank computes the same partition
        loc = fdist.extract_local(nodes, elements, part, rank)
        U = np.concatenate([U[v * nn_global + loc.node_gid] for v in range(nvar)])
        if T0 is not None:
            T0 = T0[loc.node_gid]
        part_stats = dict(owned_nodes=int(loc.owned.sum()), local_nodes=len(loc.nodes), local_elements=len(loc.elements),
                          halo_element_share=float(len(loc.elements) * world / n_el - 1.0))  # fmt: skip
        nodes, elements = loc.nodes, loc.elements.astype(np.int32)
    nn = len(nodes)
    U_host = torch.empty(U.size, dtype=torch.float64, pin_memory=True)
    U_host.copy_(torch.from_numpy(U))

    def build(nodes, elements, T_start, **kargs):
        """Mesh, law, weak form, assembly and problem of this configuration -> (assembly, problem, set_state)."""
        fd.Assembly.delete_memory()
        fd.ModelingSpace("3D")
        fd.Mesh(nodes, elements, elm, name="Domain")
        if cfg == "heat_tet4":
            fd.constitutivelaw.ThermalProperties(500.0, 0.5, 7800.0, name="ThermalLaw")
            fd.weakform.HeatEquation("ThermalLaw")
            a = fd.Assembly.create("ThermalLaw", "Domain", name="A", **kargs)
            pb = fd.problem.NonLinear("A")
            pb.dtime = 10.0 / 3.0
            pb._U, pb._dU = T_start.copy(), 0
            pb.initialize()
            a.set_start(pb)

            def set_state(x):  # the iterate; the start temperature stays what set_start recorded
                pb._U, pb._dU = x, 0
        else:
            if cfg == "j2_plate":
                law = fd.constitutivelaw.Simcoon("EPICP", J2_PROPS, name="law")
            else:
                law = fd.constitutivelaw.ElasticIsotrop(1e5, 0.3, name="law")
            fd.weakform.StressEquilibrium(law, name="wf")
            a = fd.Assembly.create("wf", "Domain", elm, name="A", **kargs)
            pb = fd.problem.Linear("A")

            def set_state(x):
                pb.set_X(x)

        return a, pb, set_state

    a, pb, set_state = build(nodes, elements, T0, reuse_buffers=True, owned_nodes=(loc.owned if sharded else None))

    U_dev = U_host.cuda()

    def set_state_device():
        set_state(U_dev)

    def barrier():
        if sharded:
            dist.barrier()
        torch.cuda.synchronize()

    a.vector_on_device = True
    set_state_device()
    t0 = time.perf_counter()
    a.update(pb, compute="all")
    torch.cuda.synchronize()
    t_first = time.perf_counter() - t0
    if sharded:
        exch = fdist.VectorExchange(loc, nvar)
        D_global = torch.zeros(nvar * nn_global, dtype=torch.float64, device="cuda")
    # the step: state update (J2: strain + radial return + tangent; elastic / heat: lazy, nothing materialised) + K + D
    full_update = cfg == "j2_plate"

    def step_device():
        if full_update:
            a.update(pb, compute="all")
        else:
            a.assemble_global_mat("all")
        if exch is not None:
            exch.allgather(a.global_vector, D_global)

    W = max(args.warmup, 3)
    for _ in range(W):
        step_device()
    barrier()
    torch.cuda.profiler.start()
    with Clocks(torch.cuda.current_device()) as clk:
        ms_total, _ = time_loop(torch, step_device, args.steps)
        barrier()
    torch.cuda.profiler.stop()
    # the dominant kernel alone: the matrix assembly
    _, ms_matrix = time_loop(torch, lambda: a.assemble_global_mat("matrix"), max(3, args.steps // 2))
    t = torch.tensor([ms_total, ms_matrix], dtype=torch.float64, device="cuda")
    if sharded:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_total, ms_matrix = float(t[0]), float(t[1])
    ms_step = ms_total / args.steps
    K = a.get_global_matrix()
    if sharded:  # the blocks of the rows this rank owns
        ip = a._saved_bloc_structure["pattern"].blk_indptr
        nnz = nvar * nvar * int((ip[1:] - ip[:-1])[torch.from_numpy(loc.owned).to(ip.device)].sum())
    else:
        nnz = int(K.data.numel())
    nne = elements.shape[1]
    n_own_nodes = int(loc.owned.sum()) if sharded else nn
    n_own_elems = n_el / world
    if cfg == "heat_tet4":
        algo_k = 8 * nnz + 4 * nne * n_own_elems + 8 * 3 * n_own_nodes
        algo_step = algo_k + 16 * n_own_nodes
    elif cfg == "tet10":
        algo_k = 8 * nnz + 4 * nne * n_own_elems + 8 * 3 * n_own_nodes
        algo_step = algo_k + 16 * 3 * n_own_nodes
    else:  # SURVEY 8d, fused J2: statev in/out, stress out, K, conn, coords, U, D (+ the structured tangent: 80 B per GP)
        n_gp = 8 * n_own_elems
        algo_k = 8 * nnz + 4 * nne * n_own_elems + 8 * 3 * n_own_nodes + n_gp * 8 * 10  # the matrix kernel reads the structured tangent
        algo_step = 8 * nnz + 4 * nne * n_own_elems + 8 * 3 * n_own_nodes + 16 * 3 * n_own_nodes + n_gp * 8 * (8 + 8 + 6)
    peak, peak_src = hbm_peak()

    # ---- e2e: pinned host dof vector in, residual out, every step ----
    a.vector_on_device = False
    a.async_copies = True

    def e2e_loop(k):
        prev = out = None
        for _ in range(k):
            set_state(U_host)
            a.update(pb, compute="all")
            cur = a.get_global_vector()
            if prev is not None:
                out = prev.result()
            prev = cur
        return prev.result() if prev is not None else out

    e2e_loop(2)
    barrier()
    t0 = time.perf_counter()
    Dh = e2e_loop(args.steps)
    barrier()
    e2e_s = (time.perf_counter() - t0) / args.steps
    if sharded:
        t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_s = float(t[0])
    a.async_copies = False

    checks = None
    if args.check and not sharded:
        checks = other_checks(cfg, a, pb, U, T0, nn)
    elif args.check:
        a.vector_on_device = True
        step_device()
        torch.cuda.synchronize()
        s3 = torch.stack([D_global[v * nn_global : (v + 1) * nn_global].sum() for v in range(nvar)])
        spread = torch.stack([D_global.sum(), -D_global.sum()])
        dist.all_reduce(spread, op=dist.ReduceOp.MAX)
        checks = {"D_sum_rel": float(s3.abs().max() / D_global.abs().max()) if cfg != "heat_tet4" else None,
                  "D_identical_on_all_ranks": bool(float(spread[0] + spread[1]) == 0.0),
                  "D_nonzero_fraction": float((D_global != 0).double().mean())}  # fmt: skip
        if glob is not None:  # the whole mesh on this GPU alone: the gathered residual must be the one-GPU residual
            D_multi = D_global.clone()
            a1, pb1, set1 = build(glob[0], glob[1], glob[3], vector_on_device=True)
            set1(glob[2])
            a1.update(pb1, compute="all")
            err = (D_multi - a1.global_vector).abs().max() / a1.global_vector.abs().max()
            dist.all_reduce(err, op=dist.ReduceOp.MAX)
            checks["D_vs_single_gpu_rel"] = float(err)
    if sharded:
        stats_all = [None] * world
        ent = a._saved_bloc_structure  # no plan at all: the row-owner kernel needs none
        plan = next((ent[k] for k in ("plan", "plan_big", "plan_small") if ent.get(k) is not None), None)
        dist.all_gather_object(stats_all, dict(part_stats, clusters=None if plan is None else int(plan.n_clusters),
                                               heavy_nodes=None if plan is None else int(plan.heavy_nodes.numel())))  # fmt: skip
    if rank != 0:
        dist.destroy_process_group()
        return
    cb = None
    if not args.no_cpu_baseline and not sharded:
        cb = cpu_baseline(cfg, args.cpu_n or CPU_SAMPLE_EDGE[cfg], 3, 1)
        cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
    kernels = {"heat_tet4": "fdk::k_heat_tet4_rows (row-owner kernel, K + D in one launch)",
               "tet10": "fdk::k_assemble<Tet10, PHYS_ISO>", "j2_plate": "fdk::k_assemble_iso<Hex8, 512, 4, PHYS_R1>"}  # fmt: skip
    line = {
        "metric": METRICS[cfg],
        "value": n_el / (ms_step * 1e-3) / 1e6,
        "unit": "Melem/s",
        "n_gpus": world if sharded else 1,
        "steps": args.steps,
        "warmup": W,
        "ms_per_step": ms_step,
        "higher_is_better": True,
        "scaling": "strong" if sharded else "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": workload_label(args),
            "n_elements": n_el, "n_nodes": nn_global, "nnz": nnz if not sharded else None,
            "l2": f"{algo_step / 1e9:.2f} GB of compulsory traffic per step and GPU against a 126 MB L2; no flush needed",
            "partition": ("one GPU" if not sharded else
                          f"recursive coordinate bisection of the nodes over {world} GPUs, owner-computes rows, one halo layer of "
                          "elements per rank, residual exchange: pack + NCCL all-gather of D + unpack"),
            "ranks": stats_all if sharded else None,
            "first_call_s": t_first,
            "step": "Assembly.update(pb, 'all')" if full_update else "Assembly.assemble_global_mat('all')",
        },
        "roofline": {
            "bound": "hbm",
            "achieved": algo_k / (ms_matrix * 1e-3) / 1e9,
            "peak": peak,
            "unit": "GB/s",
            "frac": algo_k / (ms_matrix * 1e-3) / 1e9 / peak,
            "traffic": measured_traffic(cfg),
            "peak_source": peak_src,
            "kernel": kernels[cfg],
            "kernel_ms": ms_matrix,
            "algorithmic_bytes_per_launch": algo_k,
            "step_algorithmic_bytes": algo_step,
            "step_frac": algo_step / (ms_step * 1e-3) / 1e9 / peak,
        },
        "cpu_baseline": cb,
        "e2e": {
            "value": n_el / e2e_s / 1e6,
            "unit": "Melem/s",
            "h2d_bytes_per_step": U_host.numel() * 8,
            "d2h_bytes_per_step": int(np.asarray(Dh).size) * 8,
            "note": "every step: pinned host dof vector -> HBM, state update + K + D, residual -> pinned host (async_copies); "
            "K and the Gauss-point state stay in HBM",
        },
        "gpu_launches": args.steps * ({"heat_tet4": 1, "tet10": 1, "j2_plate": 5}[cfg] + (2 if sharded else 0)),
        "clocks": clk.summary(),
    }
    if checks is not None:
        line["checks"] = checks
    print(json.dumps(line), flush=True)
    if sharded:
        dist.destroy_process_group()


def other_checks(cfg, a, pb, U, T0, nn):
    """Size-independent properties of configs [2]-[4] on the device."""
    import torch

    a.vector_on_device = True
    a.assemble_global_mat("all")
    K = a.get_global_matrix()
    A = torch.sparse_csr_tensor(K.indptr.to(torch.int64), K.indices.to(torch.int64), K.data, size=K.shape)
    D = a.global_vector
    g = torch.Generator(device="cuda").manual_seed(0)
    nd = K.shape[0]
    x = torch.randn(nd, dtype=torch.float64, device="cuda", generator=g)
    y = torch.randn(nd, dtype=torch.float64, device="cuda", generator=g)
    sxy, syx = float(x @ (A @ y)), float(y @ (A @ x))
    out = {"symmetry_rel": abs(sxy - syx) / abs(sxy)}
    if cfg == "heat_tet4":
        one = torch.ones(nn, dtype=torch.float64, device="cuda")
        m = A @ one  # conduction rows sum to zero: what is left is the lumped capacity
        out["lumped_capacity_total_rel"] = abs(float(m.sum()) - 7800 * 0.5 / (10 / 3)) / (7800 * 0.5 / (10 / 3))
        out["capacity_positive"] = bool((m > 0).all())
    else:
        t = torch.zeros(nd, dtype=torch.float64, device="cuda")
        t[nn : 2 * nn] = 1.0
        out["rigid_translation_rel"] = float((A @ t).abs().max() / float(K.data.abs().max()))
        if cfg == "tet10":
            out["D_plus_KU_rel"] = float((D + A @ torch.from_numpy(U).cuda()).abs().max() / D.abs().max())
        else:
            sv = a.sv["Statev"]
            out["yielded_fraction"] = float((sv[:, 1] > 0).double().mean())
            out["internal_force_sum_rel"] = float(
                torch.stack([D[v * nn : (v + 1) * nn].sum() for v in range(3)]).abs().max() / D.abs().max()
            )
    return out


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.config == "hex8":
        run_hex8(a)
    else:
        run_other(a)
