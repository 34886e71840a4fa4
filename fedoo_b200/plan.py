"""Assembly plan: node clusters + per-cluster metadata consumed by the cluster kernels.

The numeric kernels (csrc/fdk_assemble.cuh) are *owner-computes*: a CTA owns a compact
cluster of nodes (hence their CSR rows) and needs, per cluster,
  * the owned nodes, their block-row start and length,
  * the touched elements (all elements incident to an owned node) with a cluster-local
    connectivity, and the touched nodes,
  * the incidences (owned node, touched element, local node index), one per thread,
  * the gather lists: for every CSR block slot (I, J) of an owned row the list of
    (incidence, local column node) whose element block contributes to it -- the
    cluster-local analogue of the reference's ``Matrix_convertCOOtoCSR``
    (fedoo/core/_sparsematrix.py:256-274), i.e. the transposed element->nnz-slot map.

Everything here is one-time symbolic work expressed with torch tensor ops (sort / unique /
searchsorted / cumsum), so it runs on the GPU in production and on the CPU in the unit
tests.  Clusters are Morton blocks of a rank-bucketed node lattice, refined until each one
fits the kernel's shared-memory / thread capacities.
"""

from __future__ import annotations

import dataclasses

import torch

from . import _lib

# per element type: CTA threads (= max incidences), max touched elements, first Morton shift
# (log2 of lattice cells per cluster).  te_max keeps te_max * ESTR * 8 B <= ~160 KB.
_CAPS = {
    "hex8": dict(nne=8, ngp=8, dim=3, inc_max=256, te_max=80, shift=5),
    "tet4": dict(nne=4, ngp=4, dim=3, inc_max=256, te_max=256, shift=4),
    "tet10": dict(nne=10, ngp=15, dim=3, inc_max=128, te_max=36, shift=5),
    "quad4": dict(nne=4, ngp=4, dim=2, inc_max=256, te_max=400, shift=6),
}
# CTA sizes the kernels are instantiated for (csrc/fdk_common.cuh ElemTraits::THREADS and half of it):
# two threads per incidence (half a block row each), so threads = 2 * inc_max
_THREADS = {"hex8": 512, "tet4": 512, "tet10": 256, "quad4": 512}
# "small" variant: half-size clusters / CTAs, two resident per SM (phases of different clusters overlap)
_CAPS_SMALL = {
    "hex8": dict(inc_max=128, te_max=45, shift=4),
    "tet4": dict(inc_max=128, te_max=128, shift=3),
    "quad4": dict(inc_max=128, te_max=200, shift=5),
}
HEAVY_T = 4  # slots with more contributions are pre-reduced by a balanced pass (csrc/fdk_assemble.cuh)
# "big" variant (tet10 + isotropic law): clusters for the balanced 1024-thread kernel, 5 threads per incidence
_CAPS_BIG = {"tet10": dict(inc_max=204, te_max=80, shift=7)}  # 5 Gauss points per geometry chunk: 1.4 KB per element
TN_MAX = 255
ENT_MAX = 65535


def _expand_ranges(starts, counts):
    """Concatenate arange(starts[i], starts[i]+counts[i])."""
    total = int(counts.sum())
    if total == 0:
        return torch.zeros(0, dtype=torch.int64, device=starts.device)
    ends = torch.cumsum(counts, 0)
    seg = torch.repeat_interleave(torch.arange(len(counts), device=starts.device), counts)
    within = torch.arange(total, device=starts.device) - (ends - counts)[seg]
    return starts[seg] + within


def _part1by2(x):
    x = x & 0x1FFFFF
    x = (x | (x << 32)) & 0x1F00000000FFFF
    x = (x | (x << 16)) & 0x1F0000FF0000FF
    x = (x | (x << 8)) & 0x100F00F00F00F00F
    x = (x | (x << 4)) & 0x10C30C30C30C30C3
    x = (x | (x << 2)) & 0x1249249249249249
    return x


def _part1by1(x):
    x = x & 0xFFFFFFFF
    x = (x | (x << 16)) & 0x0000FFFF0000FFFF
    x = (x | (x << 8)) & 0x00FF00FF00FF00FF
    x = (x | (x << 4)) & 0x0F0F0F0F0F0F0F0F
    x = (x | (x << 2)) & 0x3333333333333333
    x = (x | (x << 1)) & 0x5555555555555555
    return x


def morton_order(coords):
    """Locality-preserving node order: per-axis buckets -> Morton key -> stable sort.

    The number of buckets of an axis follows the bounding box (extent / mean node spacing), so a slab
    that is half as thick as it is wide gets half as many buckets across: Morton blocks stay compact
    bricks of nodes.  An axis whose sorted coordinates fall into well separated planes is bucketed by plane
    (the lattice planes of a structured, even jittered, box), any other axis by equal-count rank buckets."""
    n, dim = coords.shape
    if n == 0:
        e = torch.zeros(0, dtype=torch.int64, device=coords.device)
        return e, e
    lo, hi = coords.min(0).values, coords.max(0).values
    ext = (hi - lo).to(torch.float64)
    live = ext > 0
    n_live = max(1, int(live.sum()))
    vol = float(torch.prod(torch.where(live, ext, torch.ones_like(ext))))
    h = (vol / n) ** (1.0 / n_live)  # mean node spacing
    comps = []
    for d in range(dim):
        x = coords[:, d]
        nbd = max(1, int(round(float(ext[d]) / h))) if bool(live[d]) else 1
        # lattice planes: runs of the sorted coordinates separated by gaps of a good fraction of the mean
        # spacing (exact for a structured axis, robust to a jitter of +-0.2 h)
        xs, perm = torch.sort(x, stable=True)
        big = torch.zeros(n, dtype=torch.int64, device=coords.device)
        if n > 1:
            big[1:] = ((xs[1:] - xs[:-1]) > 0.25 * h).to(torch.int64)
        plane = torch.cumsum(big, 0)
        n_planes = int(plane[-1]) + 1
        if nbd // 2 <= n_planes <= 2 * nbd + 1:
            comp = torch.empty(n, dtype=torch.int64, device=coords.device)
            comp[perm] = plane
        else:  # no plane structure: equal-count buckets by rank
            comp = torch.empty(n, dtype=torch.int64, device=coords.device)
            comp[perm] = torch.arange(n, device=coords.device) * nbd // n
        comps.append(comp)
    if dim == 3:
        key = _part1by2(comps[0]) | (_part1by2(comps[1]) << 1) | (_part1by2(comps[2]) << 2)
    else:
        key = _part1by1(comps[0]) | (_part1by1(comps[1]) << 1)
    order = torch.argsort(key, stable=True)
    return order, key[order]


@dataclasses.dataclass
class Pattern:
    """Block (node-node) CSR pattern and its sorted unique keys I*n_nodes+J."""

    n_nodes: int
    blk_indptr: torch.Tensor  # int64 [n_nodes+1]
    blk_indices: torch.Tensor  # int32 [blk_nnz]
    keys: torch.Tensor  # int64 [blk_nnz] sorted unique

    @property
    def blk_nnz(self):
        return int(self.blk_indices.numel())


class Plan:
    """Cluster plan for one (mesh connectivity, element type).  Holds the device tensors and
    the C struct handed to the kernels."""

    def __init__(self, elem_type, coords, conn, pattern, caps=None, verbose=False, owned=None, small=None, big=False):
        """``owned``: optional bool mask (n_nodes,) -- only these nodes get clusters (their rows are
        assembled); used by the multi-GPU partition where a rank's local mesh carries halo nodes.
        ``small``: half-size clusters / CTAs, two resident per SM (default: env FDK_SMALL_CTA, else on)."""
        import os

        if small is None:
            # measured on B200: hex8 runs best with 32-node clusters / 512-thread
            # persistent CTAs (tensor-core producer); the other elements keep the half-size variant
            env = os.environ.get("FDK_SMALL_CTA")
            small = (elem_type != "hex8") if env is None else (env != "0")
        cap = dict(_CAPS[elem_type])
        self.threads = _THREADS[elem_type]
        if big and elem_type in _CAPS_BIG:
            cap.update(_CAPS_BIG[elem_type])
            self.threads = int(os.environ.get("FDK_BIG_THREADS", "1024"))  # 1024 or 768 threads, 5 per incidence
            cap["inc_max"] = min(cap["inc_max"], self.threads // 5)
        elif small and elem_type in _CAPS_SMALL:  # tet10: one vertex node alone can touch > 18 elements
            cap.update(_CAPS_SMALL[elem_type])
            self.threads //= 2
        if caps:
            cap.update(caps)
        if os.environ.get("FDK_PLAN_SHIFT"):  # diagnostic: initial cluster size (log2 of Morton cells), refined down as needed
            cap["shift"] = int(os.environ["FDK_PLAN_SHIFT"])
        assert 2 * cap["inc_max"] <= self.threads
        self.elem_type = elem_type
        dev = conn.device
        n_el, nne = conn.shape
        assert nne == cap["nne"]
        n_nodes = pattern.n_nodes
        self.n_nodes, self.n_elems, self.nne = n_nodes, n_el, nne
        self.pattern = pattern
        conn64 = conn.to(torch.int64)
        M = n_el * nne
        assert M < 2**31, "incidence count must fit int32"

        # ---- node -> incidences, grouped by node, element-ascending ----
        nid = conn64.reshape(-1)
        perm = torch.argsort(nid, stable=True)
        inc_e_by_node = perm // nne
        inc_l_by_node = perm % nne
        inc_count = torch.bincount(nid, minlength=n_nodes)
        inc_ptr_node = torch.zeros(n_nodes + 1, dtype=torch.int64, device=dev)
        inc_ptr_node[1:] = torch.cumsum(inc_count, 0)
        deg = pattern.blk_indptr[1:] - pattern.blk_indptr[:-1]

        # ---- high-valence nodes: one node alone must fit a cluster (its incident elements are the cluster's touched
        # elements).  Nodes beyond that are left out of the clusters; their rows are assembled by the rows kernel
        # (csrc/fdk_rows.cuh, Assembly._heavy_rows), one CTA per node.  Unstructured tet meshes have a few.
        alone_max = min(cap["te_max"], cap["inc_max"])
        heavy = inc_count > alone_max
        if owned is not None:
            heavy &= owned.to(dev)
        self.heavy_nodes = torch.nonzero(heavy).reshape(-1).to(torch.int32)
        if self.heavy_nodes.numel():
            owned = (~heavy) if owned is None else (owned.to(dev) & ~heavy)
        # ---- nodes no element refers to (the parts of an AssemblySum share one node array, fedoo/core/assembly_sum.py:
        # 35-40; isolated nodes of an imported mesh): their rows are empty, nothing to compute.  They stay out of the
        # clusters -- a cluster's slot records assume that every owned node has at least its diagonal block, and nothing
        # would bound the number of such nodes in one cluster.
        if bool((inc_count == 0).any()):
            owned = (inc_count > 0) if owned is None else (owned.to(dev) & (inc_count > 0))

        # ---- Morton order with unique keys ----
        if owned is not None:
            # only owned nodes get clusters: build the curve on THEIR lattice, so that the Morton blocks are
            # aligned with the owned region (a z-slab starts one halo layer into the local mesh)
            own_idx = torch.nonzero(owned.to(dev)).reshape(-1)
            sub_order, mk = morton_order(coords[own_idx])
            order = own_idx[sub_order]
        else:
            order, mk = morton_order(coords)
        n_mesh_nodes = n_nodes
        n_nodes = int(order.numel())  # from here on: number of OWNED nodes (cluster order)
        if n_nodes > 0:
            first = torch.ones(n_nodes, dtype=torch.bool, device=dev)
            first[1:] = mk[1:] != mk[:-1]
            run_start = torch.cummax(torch.where(first, torch.arange(n_nodes, device=dev), 0), 0).values
            within = torch.arange(n_nodes, device=dev) - run_start
            rbits = max(1, int(within.max()).bit_length())
        else:
            within = mk
            rbits = 1
        mk = (mk << rbits) | within
        shift = torch.full((n_nodes,), cap["shift"] + rbits, dtype=torch.int64, device=dev)
        inc_count_o = inc_count[order]
        deg_o = deg[order]

        # ---- refine clusters until every one fits ----
        for it in range(80):
            ck = mk >> shift
            b = torch.ones(n_nodes, dtype=torch.bool, device=dev)
            if n_nodes > 1:
                b[1:] = (ck[1:] != ck[:-1]) | (shift[1:] != shift[:-1])
            cl_of_pos = torch.cumsum(b.to(torch.int64), 0) - 1
            n_cl = int(cl_of_pos[-1]) + 1 if n_nodes > 0 else 0
            stats = self._cluster_stats(cl_of_pos, n_cl, order, inc_count_o, deg_o, inc_ptr_node, inc_e_by_node, conn64)
            n_inc_c, n_te_c, n_tn_c = stats["n_inc"], stats["n_te"], stats["n_tn"]
            bad = (
                (n_inc_c > cap["inc_max"])
                | (n_te_c > cap["te_max"])
                | (n_tn_c > TN_MAX)
                | (n_inc_c * nne + 256 > ENT_MAX)
            )
            n_bad = int(bad.sum())
            if verbose:
                print(f"[plan] iter {it}: {n_cl} clusters, {n_bad} over capacity")
            if n_bad == 0:
                break
            bad_pos = bad[cl_of_pos]
            if bool((bad_pos & (shift == 0)).any()):
                raise _lib.FdkError(
                    "a single node exceeds the cluster capacity of the kernel "
                    f"(incidences > {cap['inc_max']} or touched elements > {cap['te_max']})"
                )
            shift = torch.where(bad_pos, shift - 1, shift)
        else:
            raise _lib.FdkError("cluster refinement did not converge")

        self.n_clusters = n_cl
        counts_c = torch.bincount(cl_of_pos, minlength=n_cl)
        cl_node_ptr = torch.zeros(n_cl + 1, dtype=torch.int64, device=dev)
        cl_node_ptr[1:] = torch.cumsum(counts_c, 0)

        # ---- incidences in cluster order, node-major (grouped by owned node, element-ascending) ----
        cl_finc_ptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=dev)
        cl_finc_ptr[1:] = torch.cumsum(inc_count_o, 0)
        src = _expand_ranges(inc_ptr_node[order], inc_count_o)
        inc_e = inc_e_by_node[src]
        inc_l = inc_l_by_node[src]
        inc_q = torch.repeat_interleave(torch.arange(n_nodes, device=dev), inc_count_o)
        inc_cl = cl_of_pos[inc_q]
        n_inc_tot = int(inc_e.numel())
        cl_inc_ptr = cl_finc_ptr[cl_node_ptr]  # (n_cl+1) incidence range of each cluster (either order)

        # ---- touched elements ----
        te_keys = torch.unique(inc_cl * n_el + inc_e)  # sorted
        te_cl = te_keys // max(n_el, 1)
        te_elem = te_keys - te_cl * n_el
        cl_te_ptr = torch.zeros(n_cl + 1, dtype=torch.int64, device=dev)
        cl_te_ptr[1:] = torch.cumsum(torch.bincount(te_cl, minlength=n_cl), 0)
        te_idx = torch.searchsorted(te_keys, inc_cl * n_el + inc_e)
        le = te_idx - cl_te_ptr[inc_cl]
        assert le.numel() == 0 or int(le.max()) < 4096
        inc_desc = le | (inc_l << 12)
        cl_of_node = torch.full((n_mesh_nodes,), -1, dtype=torch.int64, device=dev)
        cl_of_node[order] = cl_of_pos
        te_own = cl_of_node[conn64[te_elem, 0]] == te_cl if len(te_elem) else torch.zeros(0, dtype=torch.bool, device=dev)

        # ---- kernel thread order: element-major (cluster, touched element, local node) ----
        # the lanes of one element then read the same dN/dx rows in phase 2 (shared-memory broadcast)
        em_perm = torch.argsort(te_idx * nne + inc_l)  # te_idx already sorts by cluster first
        em_pos = torch.empty_like(em_perm)
        em_pos[em_perm] = torch.arange(n_inc_tot, device=dev)
        nm_rank = torch.arange(n_inc_tot, device=dev) - cl_inc_ptr[inc_cl]  # node-major rank in the cluster
        # per touched element: its first thread (cluster-local) and the mask of its owned local nodes
        n_te_tot = int(te_keys.numel())
        te_inc = torch.full((n_te_tot,), 1 << 40, dtype=torch.int64, device=dev)
        te_inc.scatter_reduce_(0, te_idx, em_pos - cl_inc_ptr[inc_cl], reduce="amin")
        te_mask = torch.zeros(n_te_tot, dtype=torch.int64, device=dev)
        te_mask.scatter_add_(0, te_idx, torch.ones_like(inc_l) << inc_l)

        # ---- touched nodes + local connectivity ----
        tn_all = te_cl[:, None] * n_mesh_nodes + conn64[te_elem]  # (n_te_total, nne)
        tn_keys = torch.unique(tn_all.reshape(-1))
        tn_cl = tn_keys // max(n_mesh_nodes, 1)
        tn_node = tn_keys - tn_cl * n_mesh_nodes
        cl_tn_ptr = torch.zeros(n_cl + 1, dtype=torch.int64, device=dev)
        cl_tn_ptr[1:] = torch.cumsum(torch.bincount(tn_cl, minlength=n_cl), 0)
        lconn = torch.searchsorted(tn_keys, tn_all.reshape(-1)).reshape(-1, nne) - cl_tn_ptr[te_cl][:, None]

        # ---- slots (block rows in cluster order) ----
        cl_slot_ptr = torch.zeros(n_nodes + 1, dtype=torch.int64, device=dev)
        cl_slot_ptr[1:] = torch.cumsum(deg_o, 0)
        cl_bptr = pattern.blk_indptr[:-1][order]
        total_slots = int(cl_slot_ptr[-1])
        assert owned is not None or total_slots == pattern.blk_nnz

        # ---- staging entries: every (incidence, local column node j) block, sorted by CSR slot ----
        # entry position inside the cluster = rank among the cluster's entries in (slot, thread, j)
        # order + one gap entry per preceding block row (spreads the rows over the smem banks)
        I = order[inc_q]  # row node of each incidence
        J = conn64[inc_e]  # (M, nne)
        pos = torch.searchsorted(pattern.keys, (I[:, None] * n_mesh_nodes + J).reshape(-1)).reshape(-1, nne)
        pcol = pos - pattern.blk_indptr[I][:, None]
        slot = (cl_slot_ptr[inc_q][:, None] + pcol).reshape(-1)  # cluster-order slot id of every entry
        del pos, pcol, J, I
        n_ent_tot = n_inc_tot * nne
        tie = (em_pos[:, None] * nne + torch.arange(nne, device=dev)[None, :]).reshape(-1)
        ent_order = torch.argsort(slot * max(n_ent_tot, 1) + tie)
        rank = torch.empty_like(ent_order)
        rank[ent_order] = torch.arange(n_ent_tot, device=dev)
        del ent_order, tie
        row_local = inc_q - cl_node_ptr[inc_cl]  # local index of the row node in its cluster
        dst = rank.reshape(-1, nne) - (cl_inc_ptr[inc_cl] * nne)[:, None] + row_local[:, None]
        del rank
        assert dst.numel() == 0 or (int(dst.max()) <= ENT_MAX and int(dst.min()) >= 0)
        gcum = torch.zeros(total_slots + 1, dtype=torch.int64, device=dev)
        gcum[1:] = torch.cumsum(torch.bincount(slot, minlength=total_slots), 0)
        del slot
        slot0_c = cl_slot_ptr[cl_node_ptr]  # (n_cl+1) first slot of each cluster (+ end)
        ent0_c = gcum[slot0_c]  # == cl_inc_ptr * nne
        counts_c = cl_node_ptr[1:] - cl_node_ptr[:-1]
        # slot records (csrc/fdk_assemble.cuh phase 3): first staging entry | column-node local index << 16
        # | owner local index << 24; one end sentinel (owner 0xFF) per cluster
        slot_rec = torch.zeros(total_slots + n_cl, dtype=torch.int64, device=dev)
        slot_cnt = gcum[1:] - gcum[:-1]
        if total_slots:
            q_of_slot = torch.repeat_interleave(torch.arange(n_nodes, device=dev), deg_o)
            cl_of_slot = cl_of_pos[q_of_slot]
            sl = torch.arange(total_slots, device=dev)
            n_loc = q_of_slot - cl_node_ptr[cl_of_slot]
            e0 = gcum[:-1] - ent0_c[cl_of_slot] + n_loc
            # cluster-local touched-node index of every slot's column node (K.u residual in phase 3)
            J_of_slot = pattern.blk_indices.to(torch.int64)[_expand_ranges(cl_bptr, deg_o)]
            slot_tn = torch.searchsorted(tn_keys, cl_of_slot * n_mesh_nodes + J_of_slot) - cl_tn_ptr[cl_of_slot]
            del J_of_slot
            assert int(e0.max()) <= ENT_MAX and int(slot_tn.max()) <= 0xFF and int(n_loc.max()) < 0xFF
            slot_rec[sl + cl_of_slot] = e0 | (slot_tn << 16) | (n_loc << 24)
            # heavy slots (cluster-local index), cluster by cluster
            heavy = torch.nonzero(slot_cnt > HEAVY_T).reshape(-1)
            heavy_slot = heavy - slot0_c[cl_of_slot[heavy]]
            cl_heavy_ptr = torch.zeros(n_cl + 1, dtype=torch.int64, device=dev)
            cl_heavy_ptr[1:] = torch.cumsum(torch.bincount(cl_of_slot[heavy], minlength=n_cl), 0)
            del q_of_slot, cl_of_slot, sl, n_loc, e0, slot_tn
        else:
            heavy_slot = torch.zeros(0, dtype=torch.int64, device=dev)
            cl_heavy_ptr = torch.zeros(n_cl + 1, dtype=torch.int64, device=dev)
        if n_cl:
            cidx = torch.arange(n_cl, device=dev)
            end = (ent0_c[1:] - ent0_c[:-1]) + counts_c  # end sentinel (incl. all gaps)
            assert int(end.max()) <= ENT_MAX
            slot_rec[slot0_c[1:] + cidx] = end | (0xFF << 24)

        # ---- capacities ----
        def cmax(x):
            return int(x.max()) if x.numel() else 0

        n_slots_c = slot0_c[1:] - slot0_c[:-1]
        n_inc_c = cl_inc_ptr[1:] - cl_inc_ptr[:-1]
        self.caps = dict(
            cap_te=cmax(cl_te_ptr[1:] - cl_te_ptr[:-1]),
            cap_tn=cmax(cl_tn_ptr[1:] - cl_tn_ptr[:-1]),
            cap_inc=cmax(n_inc_c),
            cap_owned=cmax(counts_c),
            cap_slots=cmax(n_slots_c),
            cap_ent=cmax(n_inc_c * nne + counts_c),
            cap_heavy=cmax(cl_heavy_ptr[1:] - cl_heavy_ptr[:-1]),
        )
        self.stats = dict(
            n_clusters=n_cl,
            touched_elems_total=int(len(te_elem)),
            redundancy=float(len(te_elem)) / max(n_el, 1),
            mean_owned=float(n_nodes) / max(n_cl, 1),
            n_owned=n_nodes,
        )
        self.n_owned = n_nodes

        # ---- gather lists: the inverse of dst, per cluster at an even offset ent0 ----
        n_ent_c = n_inc_c * nne + counts_c
        n_ent_pad = (n_ent_c + 1) & ~1
        ent0_pad = torch.zeros(n_cl + 1, dtype=torch.int64, device=dev)
        ent0_pad[1:] = torch.cumsum(n_ent_pad, 0)
        assert int(ent0_pad[-1]) < 2**31
        ent_src = torch.zeros(int(ent0_pad[-1]), dtype=torch.int64, device=dev)
        if n_inc_tot:
            thread_loc = em_pos - cl_inc_ptr[inc_cl]  # element-major thread index of every incidence
            val = thread_loc[:, None] * nne + torch.arange(nne, device=dev)[None, :]
            ent_src[(ent0_pad[inc_cl][:, None] + dst).reshape(-1)] = val.reshape(-1)
            del val, thread_loc

        # ---- packed per-cluster header (one 64-byte load per cluster in the kernel) ----
        hdr = torch.zeros((n_cl, 16), dtype=torch.int64, device=dev)
        if n_cl:
            hdr[:, 0] = cl_node_ptr[:-1]
            hdr[:, 1] = counts_c
            hdr[:, 2] = cl_te_ptr[:-1]
            hdr[:, 3] = cl_te_ptr[1:] - cl_te_ptr[:-1]
            hdr[:, 4] = cl_tn_ptr[:-1]
            hdr[:, 5] = cl_tn_ptr[1:] - cl_tn_ptr[:-1]
            hdr[:, 6] = cl_inc_ptr[:-1]
            hdr[:, 7] = n_inc_c
            hdr[:, 8] = cl_heavy_ptr[:-1]
            hdr[:, 9] = cl_heavy_ptr[1:] - cl_heavy_ptr[:-1]
            hdr[:, 10] = slot0_c[:-1] & 0xFFFFFFFF
            hdr[:, 11] = slot0_c[:-1] >> 32
            hdr[:, 12] = n_slots_c
            hdr[:, 13] = ent0_pad[:-1]
        hdr = torch.where(hdr >= 2**31, hdr - 2**32, hdr).to(torch.int32).contiguous()  # low words as int32 bits

        # ---- device arrays in their kernel dtypes (per-incidence arrays in kernel thread order) ----
        i32, u16, u8 = torch.int32, torch.uint16, torch.uint8
        self.t = dict(
            cl_hdr=hdr,
            cl_node_ptr=cl_node_ptr.to(i32),
            cl_node=order.to(i32),
            cl_bptr=cl_bptr.contiguous(),
            cl_slot_ptr=cl_slot_ptr,
            cl_finc_ptr=cl_finc_ptr.to(i32),
            cl_slot_loc=(cl_slot_ptr[:-1] - slot0_c[cl_of_pos]).to(i32) if n_nodes else cl_slot_ptr[:0].to(i32),
            cl_finc_loc=(cl_finc_ptr[:-1] - cl_inc_ptr[cl_of_pos]).to(i32) if n_nodes else cl_slot_ptr[:0].to(i32),
            cl_inc_ptr=cl_inc_ptr.to(i32),
            inc_desc=inc_desc[em_perm].to(u16),
            ent_src=ent_src.to(u16),
            inc_fdst=nm_rank[em_perm].to(u16),
            cl_te_ptr=cl_te_ptr.to(i32),
            cl_te_elem=te_elem.to(i32),
            cl_te_own=te_own.to(u8),
            # one row of padding: the kernel copies the aligned 4-byte words that cover a cluster's byte range
            cl_lconn=torch.cat([lconn.to(u8), torch.zeros((1, nne), dtype=u8, device=dev)]).contiguous(),
            te_desc=((te_inc & 0xFFFF) | ((te_mask & 0xFFFF) << 16)).to(torch.uint32),
            cl_tn_ptr=cl_tn_ptr.to(i32),
            cl_tn_node=tn_node.to(i32),
            slot_rec=slot_rec.to(torch.uint32),
            cl_heavy_ptr=cl_heavy_ptr.to(i32),
            heavy_slot=heavy_slot.to(torch.uint32),
        )
        self._structs = {}
        # hex8: conflict-free staging layout of the balanced kernel, computed by the library (csrc/fdk_color.cuh);
        # device plans only (the CPU emulator of the tests follows the generic kernel's layout)
        if elem_type == "hex8" and dev.type == "cuda" and n_cl > 0:
            import ctypes as C

            blk_slot = torch.zeros(max(n_inc_tot * nne, 1), dtype=torch.uint8, device=dev)
            ent_pos = torch.zeros(max(int(ent0_pad[-1]), 1), dtype=torch.uint16, device=dev)
            _lib.check(
                _lib.load().fdk_plan_color_blocks(C.byref(self.struct(3)), _lib.ptr(blk_slot), _lib.ptr(ent_pos),
                                                  _lib.current_stream()),
                "fdk_plan_color_blocks",
            )  # fmt: skip
            self.t["blk_slot"], self.t["ent_pos"] = blk_slot, ent_pos
            self._structs = {}

    @staticmethod
    def _cluster_stats(cl_of_pos, n_cl, order, inc_count_o, deg_o, inc_ptr_node, inc_e_by_node, conn64):
        dev = cl_of_pos.device
        n_nodes = int(inc_ptr_node.numel()) - 1  # mesh nodes
        n_el = conn64.shape[0]
        n_inc = torch.zeros(n_cl, dtype=torch.int64, device=dev).index_add_(0, cl_of_pos, inc_count_o)
        src = _expand_ranges(inc_ptr_node[order], inc_count_o)
        inc_e = inc_e_by_node[src]
        inc_cl = torch.repeat_interleave(cl_of_pos, inc_count_o)
        te_keys = torch.unique(inc_cl * n_el + inc_e)
        te_cl = te_keys // max(n_el, 1)
        n_te = torch.bincount(te_cl, minlength=n_cl)
        te_elem = te_keys - te_cl * n_el
        tn_keys = torch.unique((te_cl[:, None] * n_nodes + conn64[te_elem]).reshape(-1))
        n_tn = torch.bincount(tn_keys // max(n_nodes, 1), minlength=n_cl)
        return dict(n_inc=n_inc, n_te=n_te, n_tn=n_tn)

    def struct(self, nvar):
        """C struct for an operator with ``nvar`` variables per node."""
        if nvar not in self._structs:
            s = _lib.PlanStruct()
            s.elem_type = _lib.ELEM_IDS[self.elem_type]
            s.n_nodes = self.n_nodes
            s.n_elems = self.n_elems
            s.n_clusters = self.n_clusters
            s.nvar = nvar
            s.blk_nnz = self.pattern.blk_nnz
            s.threads = self.threads
            for k, v in self.caps.items():
                setattr(s, k, v)
            for k, v in self.t.items():
                setattr(s, k, v.data_ptr())
            self._structs[nvar] = s
        return self._structs[nvar]

    def metadata_bytes(self):
        return sum(v.numel() * v.element_size() for v in self.t.values())
