"""NumPy emulation of the cluster kernel's phases 2-3 driven by a Plan (CPU, tests only).

It consumes exactly the arrays the CUDA kernel consumes (incidence descriptors, local
connectivity, gather lists, slot addressing) but takes the per-Gauss-point geometry from the
oracle, so a wrong plan shows up as a wrong K on the CPU, without a GPU."""

import numpy as np
import torch

from fedoo_b200.plan import Pattern, Plan
from oracle import fedoo_oracle as fo


def numpy_pattern(elements, n_nodes):
    e = np.asarray(elements, dtype=np.int64)
    nne = e.shape[1]
    row = np.repeat(e, nne, axis=1).reshape(-1)
    col = np.tile(e, (1, nne)).reshape(-1)
    keys = np.unique(row * n_nodes + col)
    indices = (keys % n_nodes).astype(np.int32)
    indptr = np.searchsorted(keys, np.arange(n_nodes + 1, dtype=np.int64) * n_nodes)
    return Pattern(n_nodes, torch.from_numpy(indptr.astype(np.int64)), torch.from_numpy(indices), torch.from_numpy(keys))


def make_plan(elem_type, nodes, elements, **kw):
    pat = numpy_pattern(elements, len(nodes))
    plan = Plan(elem_type, torch.from_numpy(np.asarray(nodes, dtype=float)), torch.from_numpy(np.asarray(elements, dtype=np.int32)), pat, **kw)
    return plan, pat


def slot0_chk(t, q0):
    return t["cl_slot_ptr"][q0]


def emulate_iso(plan, pat, nodes, elements, lam, mu, allow_unwritten=False):
    """Follows csrc/fdk_assemble.cuh: threads = incidences in element-major order, blocks scattered to
    the slot-sorted staging array (inc_dst), heavy slots pre-reduced, slot gather by contiguous runs."""
    from fedoo_b200.plan import HEAVY_T

    def host(v):
        if v.dtype == torch.uint16:
            return v.view(torch.int16).numpy().astype(np.int64) & 0xFFFF
        if v.dtype == torch.uint32:
            return v.view(torch.int32).numpy().astype(np.int64) & 0xFFFFFFFF
        return v.numpy()

    t = {k: host(v) for k, v in plan.t.items()}
    elements = np.asarray(elements)
    G, wdet = fo.geometry(nodes, elements, plan.elem_type)
    dim = G.shape[2]
    nne = elements.shape[1]
    nv = dim
    blk_nnz = pat.blk_nnz
    K = np.full(nv * nv * blk_nnz, np.nan)
    written = np.zeros(nv * nv * blk_nnz, dtype=np.int32)
    for c in range(plan.n_clusters):
        q0, q1 = t["cl_node_ptr"][c], t["cl_node_ptr"][c + 1]
        n_owned = q1 - q0
        te0 = t["cl_te_ptr"][c]
        tn0 = t["cl_tn_ptr"][c]
        inc0, inc1 = t["cl_inc_ptr"][c], t["cl_inc_ptr"][c + 1]
        n_inc = inc1 - inc0
        assert 2 * n_inc <= plan.threads
        hdr = t["cl_hdr"][c].astype(np.int64)
        assert list(hdr[:10]) == [q0, n_owned, te0, t["cl_te_ptr"][c + 1] - te0, tn0, t["cl_tn_ptr"][c + 1] - tn0,
                                  inc0, n_inc, t["cl_heavy_ptr"][c], t["cl_heavy_ptr"][c + 1] - t["cl_heavy_ptr"][c]]
        assert ((hdr[11] & 0xFFFFFFFF) << 32 | (hdr[10] & 0xFFFFFFFF)) == t["cl_slot_ptr"][q0]
        assert hdr[12] == t["cl_slot_ptr"][q1] - t["cl_slot_ptr"][q0]
        assert np.array_equal(t["cl_slot_loc"][q0:q1], t["cl_slot_ptr"][q0:q1] - slot0_chk(t, q0))
        assert np.array_equal(t["cl_finc_loc"][q0:q1], t["cl_finc_ptr"][q0:q1] - t["cl_finc_ptr"][q0])
        assert n_inc == t["cl_finc_ptr"][q1] - t["cl_finc_ptr"][q0]
        slot0 = t["cl_slot_ptr"][q0]
        n_slots = t["cl_slot_ptr"][q1] - slot0
        rec = t["slot_rec"][slot0 + c : slot0 + c + n_slots + 1]
        off = rec & 0xFFFF
        assert off[-1] == n_inc * nne + n_owned <= plan.caps["cap_ent"] and rec[-1] >> 24 == 0xFF
        ent0 = int(hdr[13])
        assert ent0 % 2 == 0
        n_ent = n_inc * nne + n_owned
        ent_src = t["ent_src"][ent0 : ent0 + n_ent]
        blocks = np.full((n_inc, nne, dim, dim), np.nan)  # incidence-major staging (kernel thread order)
        fdst_seen = np.zeros(n_inc, dtype=int)
        owned_nodes = t["cl_node"][q0:q1]
        prev_le = -1
        for m in range(inc0, inc1):  # one kernel thread each
            desc = int(t["inc_desc"][m])
            le, i = desc & 0xFFF, desc >> 12
            assert le >= prev_le  # element-major thread order
            prev_le = le
            e = t["cl_te_elem"][te0 + le]
            node = elements[e, i]
            n = int(np.nonzero(owned_nodes == node)[0][0])  # the row node is owned by this cluster
            lc = t["cl_lconn"][te0 + le]
            assert np.array_equal(t["cl_tn_node"][tn0 + lc.astype(np.int64)], elements[e])
            fd = int(t["inc_fdst"][m])
            assert t["cl_finc_ptr"][q0 + n] - t["cl_finc_ptr"][q0] <= fd < t["cl_finc_ptr"][q0 + n + 1] - t["cl_finc_ptr"][q0]
            fdst_seen[fd] += 1
            blocks[m - inc0] = np.einsum("g,gc,gaj->jca", wdet[e], G[e, :, :, i], G[e])
        assert (fdst_seen == 1).all()
        # per touched element: first thread and owned-node mask (tensor-core producer, hex8)
        if nne <= 8:
            descs = t["inc_desc"][inc0:inc1]
            for le in range(t["cl_te_ptr"][c + 1] - te0):
                mine = np.nonzero((descs & 0xFFF) == le)[0]
                assert len(mine) > 0 and mine[0] == (t["te_desc"][te0 + le] & 0xFFFF) and (np.diff(mine) == 1).all()
                mask = 0
                for m in mine:
                    mask |= 1 << int(descs[m] >> 12)
                assert mask == t["te_desc"][te0 + le] >> 16
                assert list(descs[mine] >> 12) == sorted(descs[mine] >> 12)
        used = np.zeros(n_inc * nne, dtype=int)
        # heavy slots
        h0, h1 = t["cl_heavy_ptr"][c], t["cl_heavy_ptr"][c + 1]
        heavy = set(int(x) for x in t["heavy_slot"][h0:h1])
        for n in range(n_owned):
            sb0 = t["cl_slot_ptr"][q0 + n] - slot0
            deg = t["cl_slot_ptr"][q0 + n + 1] - t["cl_slot_ptr"][q0 + n]
            bp = t["cl_bptr"][q0 + n]
            I = owned_nodes[n]
            for pcol in range(deg):
                s = sb0 + pcol
                e0 = int(off[s])
                assert rec[s] >> 24 == n
                cnt = int(off[s + 1]) - e0 - (1 if (rec[s] >> 24) != (rec[s + 1] >> 24) else 0)
                assert (cnt > HEAVY_T) == (s in heavy)
                acc = np.zeros((dim, dim))
                for ent in ent_src[e0 : e0 + cnt]:
                    assert used[ent] == 0
                    used[ent] = 1
                    acc += blocks[ent // nne, ent % nne]
                assert not np.isnan(acc).any()
                Jn = pat.blk_indices[bp + pcol].item()
                assert t["cl_tn_node"][tn0 + int((rec[s] >> 16) & 0xFF)] == Jn
                assert cnt == np.sum((elements == I).any(axis=1) & (elements == Jn).any(axis=1))
                Kb = lam * acc + mu * acc.T + mu * np.trace(acc) * np.eye(dim)
                for cc in range(nv):
                    for aa in range(nv):
                        dst = cc * nv * blk_nnz + nv * bp + aa * deg + pcol
                        K[dst] = Kb[cc, aa]
                        written[dst] += 1
        assert used.sum() == n_inc * nne  # every block is gathered exactly once
    assert (written <= 1).all() and (allow_unwritten or (written == 1).all()), "every CSR value must be written exactly once"
    return K
