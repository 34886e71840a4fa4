#!/usr/bin/env python
"""Headline benchmark: global K + residual assembly, hex8 linear elastic, 200^3-element box.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--n 200] [--jitter 0|1] [--impl reference]

One "step" = one pass of the hot path over the whole mesh: the fused cluster kernel writes
every CSR value of K and every entry of the residual D = -int B^T sigma(U) (sigma recomputed
from U on the fly).  N > 1: launched by torchrun, one rank per GPU; nodes are partitioned in
z-slabs, each rank assembles the rows it owns from its local mesh (strong scaling: the global
mesh is fixed), then the owned slices of D are all-gathered over NCCL.

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

ALGO_BYTES_PER_ELEM = None  # filled from the actual mesh: 8 nnz + 4 nne n_el + 8 ndim n_nodes + 16 nvar n_nodes


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n", type=int, default=200, help="elements per box edge (200 -> 8 M hex8)")
    ap.add_argument("--jitter", type=int, default=0, help="1: displace interior nodes (no identical elements)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-n", type=int, default=40, help="edge of the bounded CPU-baseline sample (40 -> 64 k hex8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--exchange", default="peer", choices=["peer", "nccl"],
                    help="N > 1: residual exchange fused into the kernel over NVLink peer memory, or NCCL all-gather")
    ap.add_argument("--check", action="store_true", help="verify size-independent properties of the assembled K, D")
    return ap.parse_args()


# ----------------------------------------------------------------------------------------------
# CPU baseline: the oracle port of the reference path, timed on the host cores
# ----------------------------------------------------------------------------------------------
def cpu_baseline(n, steps, warmup):
    """Steady-state K + D assembly of the reference algorithm (oracle port) on an n^3-element box:
    geometry operators and the symbolic pattern are cached, as in the reference's 2nd call
    ('steady state (2nd call)'); timed: term-by-term batched matmuls, COO->CSR gather,
    3x3 block tiling, strain/stress update and the residual."""
    from fedoo_b200 import meshgen
    from oracle import fedoo_oracle as fo

    nodes, elements = meshgen.box_hex8(n + 1, n + 1, n + 1)
    H = fo.elastic_isotropic_H(200e3, 0.3)
    pat = fo.Pattern(elements, len(nodes), 3)
    pat.indices  # symbolic part, one-time
    G, wdet = fo.geometry(nodes, elements, "hex8")
    U = np.random.default_rng(0).standard_normal(3 * len(nodes)) * 1e-3

    def step():
        data = fo.stiffness_blocks(G, wdet, H, 3)
        blocks = [[pat.block_values(data[a][b]) for b in range(3)] for a in range(3)]
        K = pat.csr(pat.tile_values(blocks))
        sig = fo.stress_gp(H, fo.strain_gp(G, elements, U, len(nodes), 3))
        D = fo.residual(G, wdet, elements, sig, len(nodes), 3)
        return K, D

    for _ in range(warmup):
        step()
    c0, t0 = time.process_time(), time.perf_counter()
    for _ in range(steps):
        step()
    wall = time.perf_counter() - t0
    cpu = time.process_time() - c0
    melem = len(elements) * steps / wall / 1e6
    return dict(
        value=melem,
        unit="Melem/s",
        cores=max(1, int(round(cpu / wall))),
        kind="port",
        sample=f"hex8 box {n}^3 = {len(elements)} elements, {steps} steady-state K+D steps after {warmup} warm-up "
        f"(geometry operators + CSR structure cached, as the reference's 2nd call); numpy {np.__version__}, "
        f"host logical cpus {os.cpu_count()}, OMP/OPENBLAS_NUM_THREADS="
        f"{os.environ.get('OMP_NUM_THREADS', 'unset')}/{os.environ.get('OPENBLAS_NUM_THREADS', 'unset')}",
        ms_per_step=wall / steps * 1e3,
    )


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cb = cpu_baseline(args.cpu_n, args.steps, args.warmup)
    line = {
        "impl": "reference",
        "metric": "Global K+R assembly Melem/s (hex8, 8M elems)",
        "value": cb["value"],
        "unit": "Melem/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": cb["ms_per_step"],
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": workload_config(args, sample=cb["sample"]),
        "cpu_baseline": {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": cb["value"], "unit": "Melem/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, **extra):
    cfg = {
        "workload": f"hex8 box {args.n}x{args.n}x{args.n} ({args.n**3} elements), ElasticIsotrop E=200e3 nu=0.3, "
        "K + residual assembly (configs[1])",
        "jitter": bool(args.jitter),
        "l2": "working set (>= 16 GB written per step at n=200) far exceeds the 126 MB L2; no flush needed",
        "partition": f"z-slabs of nodes over {args.gpus} GPU(s), owner-computes rows, residual exchange: "
        + ("none (1 GPU)" if args.gpus == 1 else extra.pop("exchange", "NCCL all-gather of D")),
    }
    cfg.update(extra)
    return cfg


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


# ----------------------------------------------------------------------------------------------
def run_b200(args):
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
    mesh = fd.Mesh(loc.nodes, loc.elements, "hex8", name="Domain")
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
    plan, pattern = entry["plan"], entry["pattern"]
    # N > 1: the exchange of the residual.  Default: fused into the assembly kernel (stores over NVLink into a
    # symmetric, multicast-mapped global vector, fedoo_b200.dist.PeerVector); --exchange nccl, or a failed
    # symmetric-memory rendezvous: pack + NCCL all-gather + unpack (fedoo_b200.dist.VectorExchange)
    exch = peer = D_global = None
    if world > 1:
        if args.exchange == "peer":
            try:
                peer = fdist.PeerVector(loc, 3)
                asm.peer_vector = peer
                D_global = peer.tensor
            except Exception as e:  # noqa: BLE001
                if rank == 0:
                    print(f"[bench] symmetric memory unavailable ({type(e).__name__}: {e}); NCCL all-gather", file=sys.stderr)
                peer = None
        if peer is None:
            exch = fdist.VectorExchange(loc, 3)
            D_global = torch.zeros(3 * n_nodes_global, dtype=torch.float64, device="cuda")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident timing: `value` ----
    asm.vector_on_device = True

    def step_device():
        asm.assemble_global_mat("all")
        if exch is not None:
            exch.allgather(asm.global_vector, D_global)

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
                exch.allgather(asm.global_vector, D_global)
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

    # ---- end-to-end timing through the public API with HOST buffers: `e2e` ----
    asm.vector_on_device = False

    def step_e2e():
        pb.set_X(U_host)  # pinned host dof vector -> H2D inside update()
        asm.update(pb, compute="all")  # ... kernel ... D2H of the residual into a pinned buffer
        return asm.get_global_vector()

    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        Dh = step_e2e()
    barrier()
    e2e_s = time.perf_counter() - t0
    t = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    e2e_value = n_elems_global / (float(t[0]) / args.steps) / 1e6
    h2d = U_host.numel() * 8
    d2h = int(np.asarray(Dh).size) * 8

    # ---- roofline of the dominant kernel (this rank's share) ----
    n_own_nodes = plan.n_owned
    n_own_elems = n_elems_global / world  # owner-computes: a rank's algorithmic share of the elements
    nnz_local = 9 * int(plan.t["cl_slot_ptr"][-1])
    algo_bytes = 8 * nnz_local + 4 * 8 * n_own_elems + 8 * 3 * n_own_nodes + 16 * 3 * n_own_nodes
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0))
    achieved = algo_bytes / (ms_kernel_max * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        try:
            traffic = json.load(open(tp)).get(f"n{n}_g{world}")
        except Exception:
            traffic = None

    from fedoo_b200 import _lib as _fdk_lib

    kernel_name = ("fdk::k_assemble_iso<Hex8, 1024 threads, 4 per incidence>" if _fdk_lib.get_option("iso4")
                   else "fdk::k_assemble<Hex8, PHYS_ISO>")  # fmt: skip
    checks = None
    if args.check and world == 1:
        checks = property_checks(asm, pb, U_host, n)
    elif args.check:
        # the gathered residual: identical on every rank, and self-equilibrated (sum over the nodes = 0 per component),
        # which fails if any rank's run of owned entries did not land
        asm.vector_on_device = True
        step_device()
        s3 = torch.stack([D_global[v * n_nodes_global : (v + 1) * n_nodes_global].sum() for v in range(3)])
        spread = torch.stack([D_global.sum(), -D_global.sum()])
        dist.all_reduce(spread, op=dist.ReduceOp.MAX)
        checks = {
            "D_sum_rel": float(s3.abs().max() / D_global.abs().max()),
            "D_identical_on_all_ranks": bool(float(spread[0] + spread[1]) == 0.0),
            "D_nonzero_fraction": float((D_global != 0).double().mean()),
        }

    if rank == 0:
        cb = None
        if not args.no_cpu_baseline and world == 1:
            cb = cpu_baseline(args.cpu_n, 3, 1)
            cb = {k: cb[k] for k in ("value", "unit", "cores", "kind", "sample")}
        line = {
            "metric": "Global K+R assembly Melem/s (hex8, 8M elems)",
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
            "config": workload_config(
                args,
                exchange=("fused into the assembly kernel: stores over NVLink into a symmetric global vector ("
                          + ("NVSwitch multicast" if peer is not None and peer.multicast else "peer addresses") + ") + device barrier"
                          if peer is not None else "pack + NCCL all-gather of D + unpack"),
                nnz=9 * pattern.blk_nnz if world == 1 else None,
                nnz_per_s=(9 * (3 * n + 1) ** 3) / (ms_step * 1e-3),
                clusters=plan.n_clusters,
                geometry_redundancy=plan.stats["redundancy"] if world == 1 else None,
                plan_metadata_bytes=plan.metadata_bytes(),
                first_call_s=t_first,
            ),
            "roofline": {
                "bound": "hbm",
                "achieved": achieved,
                "peak": peak,
                "unit": "GB/s",
                "frac": achieved / peak,
                "traffic": traffic,
                "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6.65 TB/s",
                "kernel": kernel_name,
                "binding_resource": "shared FP64 / shared-memory issue path of the SM (ncu: LSU data pipe 75 % + FP64 pipe "
                "24 % busy, DRAM 13 %); see DESIGN.md section 7",
                "kernel_ms": ms_kernel_max,
                "algorithmic_bytes_per_launch": algo_bytes,
            },
            "cpu_baseline": cb,
            "e2e": {
                "value": e2e_value,
                "unit": "Melem/s",
                "h2d_bytes_per_step": h2d,
                "d2h_bytes_per_step": d2h,
                "note": "pinned host U -> HBM, fused K+R kernel, residual -> pinned host; K stays in HBM "
                "(DeviceCSR, materialised to scipy only on demand)",
            },
            # the assembly kernel, plus pack / unpack of the residual exchange (NCCL's own kernels not counted)
            "gpu_launches": args.steps * (1 if exch is None else (3 if exch.seg_pack is not None else 6)),
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


if __name__ == "__main__":
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_b200(a)
