"""Multi-GPU sharding of the assembly path: one process per GPU, owner-computes.

The reference is single-process (no partitioner, no collective anywhere).  Here the
NODES are partitioned over the ranks; a rank owns the CSR rows of its nodes and assembles them
completely from its local mesh = every element touching an owned node (one layer of halo
elements, recomputed redundantly -- exactly the scheme the CTAs use inside one GPU).  Hence the
matrix needs NO exchange: K stays row-distributed in HBM and the rows a rank holds are
bit-identical to the same rows of the single-GPU matrix.  The only exchange step of the path is
the global vector: every rank all-gathers the owned slices of D over NCCL (NVLink/NVSwitch).

* ``partition_slabs``    contiguous slabs along the slowest axis of a structured numbering
* ``partition_rcb``      recursive coordinate bisection on node coordinates (unstructured)
* ``extract_local``      local mesh of a rank (monotone renumbering keeps column order)
* ``VectorExchange``     pack owned entries -> all_gather -> scatter into the global vector
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib


def partition_slabs(n_nodes: int, n_parts: int, plane: int = 1) -> np.ndarray:
    """part[I] for contiguous, plane-aligned node ranges (box_mesh numbers nodes plane by plane)."""
    n_planes = (n_nodes + plane - 1) // plane
    bounds = np.round(np.linspace(0, n_planes, n_parts + 1)).astype(np.int64) * plane
    part = np.zeros(n_nodes, dtype=np.int32)
    for r in range(n_parts):
        part[bounds[r] : min(bounds[r + 1], n_nodes)] = r
    return part


def partition_rcb(coords: np.ndarray, n_parts: int) -> np.ndarray:
    """Recursive coordinate bisection (weighted median splits along the longest axis)."""
    coords = np.asarray(coords, dtype=float)
    part = np.zeros(len(coords), dtype=np.int32)

    def rec(idx, p0, np_):
        if np_ == 1:
            part[idx] = p0
            return
        left = np_ // 2
        x = coords[idx]
        axis = int(np.argmax(x.max(axis=0) - x.min(axis=0)))
        k = int(round(len(idx) * left / np_))
        o = np.argsort(x[:, axis], kind="stable")
        rec(idx[o[:k]], p0, left)
        rec(idx[o[k:]], p0 + left, np_ - left)

    rec(np.arange(len(coords)), 0, n_parts)
    return part


class LocalMesh:
    """What one rank needs: local nodes/elements, the owned mask and the maps back to global."""

    def __init__(self, nodes, elements, owned, node_gid, elem_gid, n_global_nodes):
        self.nodes, self.elements, self.owned = nodes, elements, owned
        self.node_gid, self.elem_gid, self.n_global_nodes = node_gid, elem_gid, n_global_nodes

    @property
    def owned_gid(self):
        return self.node_gid[self.owned]


def extract_local(nodes, elements, part, rank) -> LocalMesh:
    """Elements touching a node owned by ``rank`` and their nodes, renumbered monotonically
    (local order == global order), so local block rows list columns in the global order."""
    elements = np.asarray(elements)
    own_node = np.asarray(part) == rank
    touch = own_node[elements].any(axis=1)
    elem_gid = np.nonzero(touch)[0]
    el = elements[elem_gid]
    node_gid = np.unique(el)
    lut = np.full(len(nodes), -1, dtype=np.int64)
    lut[node_gid] = np.arange(len(node_gid))
    # isolated owned nodes (no element) still belong to the rank
    extra = np.nonzero(own_node & (lut < 0))[0]
    if len(extra):
        node_gid = np.union1d(node_gid, extra)
        lut[:] = -1
        lut[node_gid] = np.arange(len(node_gid))
    return LocalMesh(
        np.ascontiguousarray(np.asarray(nodes)[node_gid]), lut[el], own_node[node_gid], node_gid, elem_gid, len(nodes)
    )


def box_local_slab(n, rank, world, jitter=False, seed=1):
    """Rank-local part of the (n x n x n element) unit box without ever building the global mesh:
    the rank owns node planes [z0, z1) and holds element layers [z0-1, z1-1] clipped to the box."""
    from . import meshgen

    nn = n + 1
    zb = np.round(np.linspace(0, nn, world + 1)).astype(np.int64)
    z0, z1 = int(zb[rank]), int(zb[rank + 1])
    k0, k1 = max(z0 - 1, 0), min(z1 - 1, n - 1)  # element layers, inclusive
    p0, p1 = k0, k1 + 1  # local node planes, inclusive
    h = 1.0 / n
    nodes, elements = meshgen.box_hex8(nn, nn, p1 - p0 + 1, 0.0, 1.0, 0.0, 1.0, p0 * h, p1 * h)
    plane = nn * nn
    node_gid = np.arange(p0 * plane, (p1 + 1) * plane, dtype=np.int64)
    if jitter:
        # jitter must be a function of the GLOBAL node id so that ranks agree on shared nodes
        nodes = _jitter_global(nodes, node_gid, nn, h, seed)
    owned = (node_gid >= z0 * plane) & (node_gid < z1 * plane)
    elem_gid = np.arange(k0 * n * n, (k1 + 1) * n * n, dtype=np.int64)
    return LocalMesh(nodes, elements, owned, node_gid, elem_gid, nn**3)


def _jitter_global(nodes, node_gid, nn, h, seed, amplitude=0.2):
    """Deterministic per-node displacement from a hash of the global id (interior nodes only)."""
    g = node_gid.astype(np.uint64)
    ix, iy, iz = g % nn, (g // nn) % nn, g // (nn * nn)
    interior = (ix > 0) & (ix < nn - 1) & (iy > 0) & (iy < nn - 1) & (iz > 0) & (iz < nn - 1)
    out = np.array(nodes, dtype=float, copy=True)
    for d in range(3):
        x = (g * np.uint64(3) + np.uint64(d) + np.uint64(seed) * np.uint64(0x9E3779B97F4A7C15)) & np.uint64(0xFFFFFFFFFFFFFFFF)
        x ^= x >> np.uint64(30)
        x *= np.uint64(0xBF58476D1CE4E5B9)
        x ^= x >> np.uint64(27)
        x *= np.uint64(0x94D049BB133111EB)
        x ^= x >> np.uint64(31)
        u = (x >> np.uint64(11)).astype(np.float64) / float(1 << 53)  # [0, 1)
        out[:, d] += np.where(interior, (2 * u - 1) * amplitude * h, 0.0)
    return out


class VectorExchange:
    """All-gather of the owned entries of the global vector (the one exchange step of the path).

    Each rank packs D_local[var * n_local + owned] into a fixed-size send buffer
    (``fdk_gather_f64``), ``all_gather_into_tensor`` moves the buffers over NCCL, and
    ``fdk_scatter_add_f64``-free plain indexing writes them at var * n_global + gid."""

    def __init__(self, local: LocalMesh, nvar: int, group=None):
        import torch.distributed as dist

        self.dist, self.group, self.nvar = dist, group, nvar
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        dev = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
        self.dev = dev
        n_loc = len(local.nodes)
        own_l = np.nonzero(local.owned)[0]
        own_g = local.node_gid[own_l]
        n_own = torch.tensor([len(own_l)], dtype=torch.int64, device=dev)
        counts = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(counts, n_own, group=group)
        else:
            counts = [n_own]
        self.counts = [int(c) for c in counts]
        self.max_own = max(self.counts)
        self.n_global = local.n_global_nodes
        src = np.concatenate([v * n_loc + own_l for v in range(nvar)])
        self.src_index = torch.from_numpy(src.astype(np.int64)).to(dev)
        self.n_send = nvar * self.max_own
        # destination index of every received entry (-> global dof), gathered once
        gid_pad = np.full(self.max_own, -1, dtype=np.int64)
        gid_pad[: len(own_g)] = own_g
        gid_t = torch.from_numpy(gid_pad).to(dev)
        all_gid = [torch.empty_like(gid_t) for _ in range(self.world)]
        if self.world > 1:
            dist.all_gather(all_gid, gid_t, group=group)
        else:
            all_gid = [gid_t]
        self.all_gid = torch.stack(all_gid)  # (world, max_own)
        self.send = torch.zeros(self.n_send, dtype=torch.float64, device=dev)
        self.recv = torch.zeros(self.world * self.n_send, dtype=torch.float64, device=dev)
        # flat (src position in recv, dst global dof) for valid entries
        valid = self.all_gid >= 0  # (world, max_own)
        pos, dst = [], []
        for v in range(nvar):
            base = torch.arange(self.world, device=dev)[:, None] * self.n_send + v * self.max_own
            p = base + torch.arange(self.max_own, device=dev)[None, :]
            pos.append(p[valid])
            dst.append(v * self.n_global + self.all_gid[valid])
        self.recv_pos = torch.cat(pos).contiguous()
        self.dst_index = torch.cat(dst).contiguous()
        self.n_own = len(own_l)
        # contiguous runs (slab partitions: the owned nodes of a rank are one range of local AND of global ids):
        # pack and unpack become one segment-copy launch each, without index arrays or zero-fill
        self.seg_pack = self.seg_unpack = None
        gid_all = self.all_gid.cpu().numpy()
        runs_ok = len(own_l) > 0 and np.array_equal(own_l, np.arange(own_l[0], own_l[0] + len(own_l)))
        for r in range(self.world):
            c = self.counts[r]
            runs_ok = runs_ok and c > 0 and np.array_equal(gid_all[r, :c], np.arange(gid_all[r, 0], gid_all[r, 0] + c))
        if runs_ok and dev.type == "cuda":
            def seg(rows):
                a = np.asarray(rows, dtype=np.int64).reshape(-1, 3)
                return tuple(torch.from_numpy(np.ascontiguousarray(a[:, k])).to(dev) for k in range(3)) + (int(a[:, 2].max()),)

            self.seg_pack = seg([(v * n_loc + own_l[0], v * self.max_own, len(own_l)) for v in range(nvar)])
            self.seg_unpack = seg([(r * self.n_send + v * self.max_own, v * self.n_global + gid_all[r, 0], self.counts[r])
                                   for r in range(self.world) for v in range(nvar)])  # fmt: skip

    def allgather(self, D_local: torch.Tensor, D_global: torch.Tensor | None = None) -> torch.Tensor:
        if D_global is None:
            D_global = torch.zeros(self.nvar * self.n_global, dtype=torch.float64, device=self.dev)
        segs = D_local.is_cuda and self.seg_pack is not None
        if segs:
            lib = _lib.load()
            stream = _lib.current_stream()
            ss, sd, sl, mx = self.seg_pack
            _lib.check(lib.fdk_copy_segments(len(ss), _lib.ptr(ss), _lib.ptr(sd), _lib.ptr(sl), mx, _lib.ptr(D_local),
                                             _lib.ptr(self.send), stream), "fdk_copy_segments")  # fmt: skip
        elif D_local.is_cuda:
            lib = _lib.load()
            stream = _lib.current_stream()
            # pack: send[v * max_own + k] = D_local[v * n_loc + own_l[k]]
            for v in range(self.nvar):
                _lib.check(
                    lib.fdk_gather_f64(
                        self.n_own, C.c_void_p(self.src_index.data_ptr() + 8 * v * self.n_own), _lib.ptr(D_local),
                        C.c_void_p(self.send.data_ptr() + 8 * v * self.max_own), stream,
                    ),
                    "fdk_gather_f64",
                )  # fmt: skip
        else:  # gloo / CPU tests of the host logic
            for v in range(self.nvar):
                self.send[v * self.max_own : v * self.max_own + self.n_own] = D_local[
                    self.src_index[v * self.n_own : (v + 1) * self.n_own]
                ]
        if self.world > 1:
            self.dist.all_gather_into_tensor(self.recv, self.send, group=self.group)
        else:
            self.recv.copy_(self.send)
        if segs:  # the runs tile the whole global vector: nothing to zero
            ss, sd, sl, mx = self.seg_unpack
            _lib.check(lib.fdk_copy_segments(len(ss), _lib.ptr(ss), _lib.ptr(sd), _lib.ptr(sl), mx, _lib.ptr(self.recv),
                                             _lib.ptr(D_global), _lib.current_stream()), "fdk_copy_segments")  # fmt: skip
        elif D_local.is_cuda:
            lib = _lib.load()
            tmp = torch.empty(self.recv_pos.numel(), dtype=torch.float64, device=self.dev)
            _lib.check(lib.fdk_gather_f64(self.recv_pos.numel(), _lib.ptr(self.recv_pos), _lib.ptr(self.recv), _lib.ptr(tmp),
                                          _lib.current_stream()), "fdk_gather_f64")  # fmt: skip
            D_global.zero_()
            _lib.check(lib.fdk_scatter_add_f64(self.dst_index.numel(), _lib.ptr(self.dst_index), _lib.ptr(tmp),
                                               _lib.ptr(D_global), _lib.current_stream()), "fdk_scatter_add_f64")  # fmt: skip
        else:
            D_global.zero_()
            D_global[self.dst_index] = self.recv[self.recv_pos]
        return D_global


class PeerVector:
    """The global vector replicated on every rank in SYMMETRIC memory (torch.distributed._symmetric_memory), so that
    the assembly kernel itself can deliver the owned entries of the residual into every rank's copy over NVLink --
    one store to the allocation's multicast address, which NVSwitch replicates to all GPUs, or stores to the peers'
    own addresses when no multicast object could be created.  Replaces pack + NCCL all-gather + unpack
    (``VectorExchange``) by ``fdk_assemble_elastic_iso_dist`` + a device-side barrier."""

    def __init__(self, local: LocalMesh, nvar: int, group=None, mode="copy"):
        """``mode``: "fused" -- the assembly kernel itself stores every owned entry to the multicast / peer addresses as
        soon as its cluster is done (8-byte stores scattered in Morton order: the transfer overlaps the assembly, but
        every store is its own NVLink packet); "copy" -- the kernel writes the rank-local vector only and ONE
        segment-copy launch then streams the owned slices (contiguous for slab partitions) to the multicast address with
        coalesced stores.  Measured on 8 B200 (round 2, 200^3 hex8): see DESIGN.md section 5."""
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm_mem

        self.group = group or dist.group.WORLD
        self.mode = mode
        self.nvar, self.n_global = nvar, local.n_global_nodes
        dev = torch.device("cuda", torch.cuda.current_device())
        # DOUBLE-BUFFERED: step k stores into half k % 2.  A rank can only launch step k + 1 after it has passed the
        # barrier of step k, which every rank enters after its own step-k kernel -- i.e. after everything it had
        # queued on the stream to read the result of step k - 1 (the half step k + 1 overwrites).  One barrier per
        # step therefore covers both the read-after-write and the write-after-read hazard.
        n = nvar * self.n_global
        self._n = n
        self._both = symm_mem.empty(2 * n, dtype=torch.float64, device=dev)
        self._both.zero_()
        self.handle = symm_mem.rendezvous(self._both, self.group)
        mc = int(self.handle.multicast_ptr or 0)
        base = [mc] if mc else [int(x) for x in self.handle.buffer_ptrs]
        assert len(base) <= 8, "peer-address fallback is limited to 8 ranks"
        self.multicast = bool(mc)
        self._dst = [[b + 8 * n * h for b in base] for h in (0, 1)]
        self._arrays = [(C.c_void_p * len(d))(*d) for d in self._dst]
        self._half = 1  # half written by the LAST step (begin_step flips it)
        self.node_gid = torch.from_numpy(np.ascontiguousarray(local.node_gid, dtype=np.int64)).to(dev)
        # "copy" mode: the owned nodes must be one run of local AND of global ids (slab partitions)
        own_l = np.nonzero(local.owned)[0]
        self._seg = None
        if len(own_l) and np.array_equal(own_l, np.arange(own_l[0], own_l[0] + len(own_l))):
            g = np.asarray(local.node_gid)[own_l]
            if np.array_equal(g, np.arange(g[0], g[0] + len(g))):
                n_loc = len(local.nodes)
                rows = np.array([(v * n_loc + own_l[0], v * self.n_global + g[0], len(own_l)) for v in range(nvar)], dtype=np.int64)
                self._seg = tuple(torch.from_numpy(np.ascontiguousarray(rows[:, k])).to(dev) for k in range(3)) + (len(own_l),)
        if mode == "copy" and self._seg is None:
            raise ValueError("PeerVector(mode='copy') needs owned nodes that are contiguous in local and global numbering")
        torch.cuda.synchronize()
        self.handle.barrier()

    def publish(self, D_local):
        """"copy" mode: stream the owned slices of the rank-local vector into this step's half of every rank's copy."""
        lib = _lib.load()
        self._half ^= 1
        ss, sd, sl, mx = self._seg
        for dst in self._dst[self._half]:
            _lib.check(lib.fdk_copy_segments(self.nvar, _lib.ptr(ss), _lib.ptr(sd), _lib.ptr(sl), mx, _lib.ptr(D_local),
                                             C.c_void_p(dst), _lib.current_stream()), "fdk_copy_segments")  # fmt: skip

    @property
    def tensor(self):
        """The global vector of the last completed step (valid after ``barrier()``, until the step after next)."""
        return self._both[self._half * self._n : (self._half + 1) * self._n]

    @property
    def dst_ptrs(self):
        return self._dst[self._half]

    def begin_step(self):
        """Flip to the other half; returns (ctypes pointer array, count) of the destinations of this step."""
        self._half ^= 1
        return self._arrays[self._half], len(self._dst[self._half])

    def barrier(self):
        """Device-side barrier on the current stream: after it every rank's stores have landed everywhere."""
        self.handle.barrier()
