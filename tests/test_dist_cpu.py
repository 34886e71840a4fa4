"""CPU tests of the multi-GPU host logic: node partition, local meshes, owned-row plans and the
global-vector all-gather (gloo, world_size 2)."""

import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from fedoo_b200 import dist as fdist
from fedoo_b200 import meshgen
from fedoo_b200.plan import Plan
from oracle import fedoo_oracle as fo

from plan_emulator import emulate_iso, numpy_pattern


def _mesh():
    nodes, el = meshgen.box_hex8(7, 6, 9)
    return meshgen.jitter_nodes(nodes, 7, 6, 9), el


@pytest.mark.parametrize("kind", ["slabs", "rcb"])
def test_owned_rows_equal_global_rows(kind):
    """Each rank's owned rows, assembled from its local mesh only, are the rows of the global K."""
    nodes, el = _mesh()
    world = 3
    part = fdist.partition_slabs(len(nodes), world, plane=7 * 6) if kind == "slabs" else fdist.partition_rcb(nodes, world)
    assert set(np.unique(part)) == set(range(world))
    lam, mu = 1.3, 0.8
    H = np.zeros((6, 6))
    H[:3, :3] = lam
    H[np.arange(3), np.arange(3)] += 2 * mu
    H[np.arange(3, 6), np.arange(3, 6)] = mu
    Kref = fo.assemble_stiffness(nodes, el, "hex8", H, 3)
    n = len(nodes)
    seen = np.zeros(n, dtype=int)
    for r in range(world):
        loc = fdist.extract_local(nodes, el, part, r)
        assert np.all(np.diff(loc.node_gid) > 0)  # monotone renumbering
        pat = numpy_pattern(loc.elements, len(loc.nodes))
        plan = Plan("hex8", torch.from_numpy(loc.nodes), torch.from_numpy(loc.elements.astype(np.int32)), pat,
                    owned=torch.from_numpy(loc.owned))  # fmt: skip
        assert plan.n_owned == int(loc.owned.sum())
        Kloc = emulate_iso(plan, pat, loc.nodes, loc.elements, lam, mu, allow_unwritten=True)
        bptr = pat.blk_indptr.numpy()
        bidx = pat.blk_indices.numpy()
        nl = len(loc.nodes)
        for li in np.nonzero(loc.owned)[0]:
            gi = loc.node_gid[li]
            seen[gi] += 1
            deg = bptr[li + 1] - bptr[li]
            for v in range(3):
                grow = Kref.getrow(v * n + gi)
                lrow_start = v * 3 * pat.blk_nnz + 3 * bptr[li]
                vals = Kloc[lrow_start : lrow_start + 3 * deg]
                cols = np.concatenate([vp * n + loc.node_gid[bidx[bptr[li] : bptr[li + 1]]] for vp in range(3)])
                assert np.array_equal(cols, grow.indices)  # same columns, same order
                assert np.abs(vals - grow.data).max() <= 1e-12 * np.abs(Kref.data).max()
    assert (seen == 1).all()


def test_box_local_slab_matches_global_extraction():
    n, world = 6, 4
    nodes, el = meshgen.box_hex8(n + 1, n + 1, n + 1)
    part = fdist.partition_slabs(len(nodes), world, plane=(n + 1) ** 2)
    for r in range(world):
        a = fdist.box_local_slab(n, r, world)
        b = fdist.extract_local(nodes, el, part, r)
        assert np.array_equal(a.node_gid, b.node_gid) and np.array_equal(a.elem_gid, b.elem_gid)
        assert np.array_equal(a.elements, b.elements) and np.array_equal(a.owned, b.owned)
        assert np.abs(a.nodes - b.nodes).max() < 1e-15
    # jitter is a function of the global node id: ranks agree on shared nodes
    j0, j1 = fdist.box_local_slab(n, 0, world, jitter=True), fdist.box_local_slab(n, 1, world, jitter=True)
    common, i0, i1 = np.intersect1d(j0.node_gid, j1.node_gid, return_indices=True)
    assert len(common) and np.array_equal(j0.nodes[i0], j1.nodes[i1])
    assert np.abs(j0.nodes - fdist.box_local_slab(n, 0, world).nodes).max() > 0


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    import torch.distributed as dist

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    nodes, el = _mesh()
    part = fdist.partition_rcb(nodes, world)
    loc = fdist.extract_local(nodes, el, part, rank)
    G, wdet = fo.geometry(nodes, el, "hex8")
    U = np.random.default_rng(0).standard_normal(3 * len(nodes))
    H = fo.elastic_isotropic_H(10.0, 0.3)
    Dref = fo.residual(G, wdet, el, fo.stress_gp(H, fo.strain_gp(G, el, U, len(nodes), 3)), len(nodes), 3)
    # local D: owned entries correct, halo entries garbage
    nl = len(loc.nodes)
    D_local = np.full(3 * nl, 1e30)
    for v in range(3):
        D_local[v * nl + np.nonzero(loc.owned)[0]] = Dref[v * len(nodes) + loc.owned_gid]
    ex = fdist.VectorExchange(loc, 3)
    Dg = ex.allgather(torch.from_numpy(D_local)).numpy()
    out[rank] = bool(np.array_equal(Dg, Dref))
    dist.destroy_process_group()


def test_vector_allgather_gloo():
    world = 2
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))
