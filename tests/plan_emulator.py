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


def emulate_iso(plan, pat, nodes, elements, lam, mu, allow_unwritten=False):
    t = {k: v.numpy() if v.dtype != torch.uint16 else v.view(torch.int16).numpy().astype(np.int64) & 0xFFFF for k, v in plan.t.items()}
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
        te0 = t["cl_te_ptr"][c]
        tn0 = t["cl_tn_ptr"][c]
        inc0, inc1 = t["cl_inc_ptr"][q0], t["cl_inc_ptr"][q1]
        slot0 = t["cl_slot_ptr"][q0]
        S = np.zeros((inc1 - inc0, nne, dim, dim))
        for n in range(q1 - q0):
            node = t["cl_node"][q0 + n]
            for m in range(t["cl_inc_ptr"][q0 + n], t["cl_inc_ptr"][q0 + n + 1]):
                desc = int(t["inc_desc"][m])
                le, i = desc & 0xFFF, desc >> 12
                e = t["cl_te_elem"][te0 + le]
                assert elements[e, i] == node
                lc = t["cl_lconn"][te0 + le]
                assert np.array_equal(t["cl_tn_node"][tn0 + lc.astype(np.int64)], elements[e])
                S[m - inc0] = np.einsum("g,gc,gaj->jca", wdet[e], G[e, :, :, i], G[e])
        gbase = t["cl_g_base"][c]
        goff = t["g_off"][slot0 + c : slot0 + c + (t["cl_slot_ptr"][q1] - slot0) + 1]
        assert goff[-1] == t["cl_g_base"][c + 1] - gbase
        for n in range(q1 - q0):
            sb0 = t["cl_slot_ptr"][q0 + n] - slot0
            deg = t["cl_slot_ptr"][q0 + n + 1] - t["cl_slot_ptr"][q0 + n]
            bp = t["cl_bptr"][q0 + n]
            for pcol in range(deg):
                s = sb0 + pcol
                acc = np.zeros((dim, dim))
                for tt in range(goff[s], goff[s + 1]):
                    ent = int(t["g_ent"][gbase + tt])
                    acc += S[ent >> 4, ent & 15]
                Kb = lam * acc + mu * acc.T + mu * np.trace(acc) * np.eye(dim)
                for cc in range(nv):
                    for aa in range(nv):
                        dst = cc * nv * blk_nnz + nv * bp + aa * deg + pcol
                        K[dst] = Kb[cc, aa]
                        written[dst] += 1
    assert (written <= 1).all() and (allow_unwritten or (written == 1).all()), "every CSR value must be written exactly once"
    return K
