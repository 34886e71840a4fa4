"""CPU tests of the cluster plan (host logic): partition properties, capacities and a NumPy
emulation of the kernel that must reproduce the reference K from the plan arrays alone."""

import os

import numpy as np
import pytest
import torch

from fedoo_b200 import meshgen
from fedoo_b200.plan import TN_MAX, _CAPS
from oracle import fedoo_oracle as fo

from mesh_util import delaunay_tet4
from plan_emulator import emulate_iso, make_plan


def _mesh(kind):
    if kind == "hex8":
        nodes, el = meshgen.box_hex8(9, 7, 6)
        return meshgen.jitter_nodes(nodes, 9, 7, 6), el
    if kind == "tet4":
        nodes, hexes = meshgen.box_hex8(6, 5, 5)
        return meshgen.jitter_nodes(nodes, 6, 5, 5), meshgen.hex8_to_tet4(hexes)
    if kind == "tet10":
        nodes, hexes = meshgen.box_hex8(4, 4, 3)
        nodes = meshgen.jitter_nodes(nodes, 4, 4, 3)
        return meshgen.tet4_to_tet10(nodes, meshgen.hex8_to_tet4(hexes), bulge=0.03)
    nodes, el = meshgen.rect_quad4(12, 9)
    return meshgen.jitter_nodes_2d(nodes, 12, 9), el


@pytest.mark.parametrize("kind", ["hex8", "tet4", "tet10", "quad4"])
def test_plan_reproduces_reference_K(kind):
    nodes, el = _mesh(kind)
    plan, pat = make_plan(kind, nodes, el)
    dim = nodes.shape[1]
    lam, mu = 115384.61538461539, 76923.07692307692
    K = emulate_iso(plan, pat, nodes, el, lam, mu)
    H = fo.elastic_isotropic_H(200e3, 0.3)
    Kref = fo.assemble_stiffness(nodes, el, kind, H, dim)
    assert np.array_equal(Kref.indptr[: len(nodes) + 1] // dim, pat.blk_indptr.numpy())
    assert np.abs(K - Kref.data).max() <= 1e-12 * np.abs(Kref.data).max()


@pytest.mark.parametrize("kind", ["hex8", "tet4", "tet10", "quad4"])
def test_plan_partition_and_caps(kind):
    nodes, el = _mesh(kind)
    plan, pat = make_plan(kind, nodes, el)
    t = plan.t
    order = t["cl_node"].numpy()
    assert np.array_equal(np.sort(order), np.arange(len(nodes)))  # every node owned exactly once
    cap = _CAPS[kind]
    assert plan.caps["cap_inc"] <= cap["inc_max"]
    assert plan.caps["cap_te"] <= cap["te_max"]
    assert plan.caps["cap_tn"] <= TN_MAX
    # every element is owned by exactly one cluster
    own = t["cl_te_own"].numpy().astype(bool)
    te = t["cl_te_elem"].numpy()
    assert np.array_equal(np.sort(te[own]), np.arange(len(el)))
    assert plan.stats["redundancy"] >= 1.0


def test_plan_forced_refinement():
    """Tiny capacities force many refinement rounds; result must stay exact."""
    nodes, el = _mesh("hex8")
    plan, pat = make_plan("hex8", nodes, el, caps=dict(inc_max=40, te_max=20))
    assert plan.caps["cap_inc"] <= 40 and plan.caps["cap_te"] <= 20
    K = emulate_iso(plan, pat, nodes, el, 1.5, 0.7)
    G, wdet = fo.geometry(nodes, el, "hex8")
    H = np.zeros((6, 6))
    H[:3, :3] = 1.5
    H[np.arange(3), np.arange(3)] += 1.4
    H[np.arange(3, 6), np.arange(3, 6)] = 0.7
    Kref = fo.assemble_stiffness(nodes, el, "hex8", H, 3)
    assert np.abs(K - Kref.data).max() <= 1e-12 * np.abs(Kref.data).max()


def test_plan_unreferenced_nodes_and_single_node_overflow():
    nodes, el = _mesh("hex8")
    n_ref = len(nodes)
    nodes = np.vstack([nodes, [[5.0, 5.0, 5.0], [6.0, 6.0, 6.0]]])  # two isolated nodes
    plan, pat = make_plan("hex8", nodes, el)
    assert plan.n_nodes == len(nodes)
    # nodes without elements have empty rows: they are not part of any cluster
    assert plan.n_owned == n_ref and plan.t["cl_node"].cpu().numpy().max() < n_ref
    # a node whose incident elements alone exceed a cluster's capacity is left out of the clusters: its row goes to the
    # rows kernel (csrc/fdk_rows.cuh).  With inc_max = 4 every node touching more than 4 elements is such a node.
    small, _ = make_plan("hex8", nodes, el, caps=dict(inc_max=4, te_max=80))
    n_inc = np.bincount(el.reshape(-1), minlength=len(nodes))
    heavy = np.flatnonzero(n_inc > 4)
    assert heavy.size > 0 and np.array_equal(small.heavy_nodes.cpu().numpy(), heavy)
    assert small.n_owned == n_ref - heavy.size
    owned_by_clusters = small.t["cl_node"].cpu().numpy()
    assert not np.intersect1d(owned_by_clusters, heavy).size
    assert plan.heavy_nodes.numel() == 0


@pytest.mark.parametrize("where", ["inside", "far", "few", "first"])
def test_plan_with_nodes_no_element_refers_to(where):
    """The parts of an AssemblySum share one node array (fedoo/core/assembly_sum.py:35-40): each part's mesh carries the
    other parts' nodes without any element.  Hundreds of them inside the bounding box, far away, a handful, or numbered
    BEFORE the referenced ones: K from the plan arrays must stay the reference K (rows of those nodes empty)."""
    nodes, el = _mesh("quad4")
    rng = np.random.default_rng(0)
    lo, hi = nodes.min(axis=0), nodes.max(axis=0)
    if where == "inside":
        extra = rng.uniform(lo, hi, (600, 2))
    elif where == "far":
        extra = np.full((600, 2), 500.0)
    else:
        extra = rng.uniform(lo, hi, (5, 2))
    if where == "first":
        nodes, el = np.vstack([extra, nodes]), el + len(extra)
    else:
        nodes = np.vstack([nodes, extra])
    plan, pat = make_plan("quad4", nodes, el)
    assert plan.n_owned == len(nodes) - len(extra)
    K = emulate_iso(plan, pat, nodes, el, 1.5, 0.7, allow_unwritten=False)
    H = np.zeros((6, 6))
    H[:3, :3] = 1.5
    H[np.arange(3), np.arange(3)] += 1.4
    H[np.arange(3, 6), np.arange(3, 6)] = 0.7
    Kref = fo.assemble_stiffness(nodes, el, "quad4", H, 2)
    assert np.array_equal(Kref.indptr[: len(nodes) + 1] // 2, pat.blk_indptr.numpy())
    assert np.abs(K - Kref.data).max() <= 1e-12 * np.abs(Kref.data).max()


def test_plan_on_delaunay_tets():
    """Unstructured tetrahedra (valence up to ~40 per vertex): the plan arrays alone reproduce the reference K."""
    pts, el = delaunay_tet4(300, 0)
    assert np.bincount(el.reshape(-1)).max() > 36
    plan, pat = make_plan("tet4", pts, el)
    K = emulate_iso(plan, pat, pts, el, 115384.61538461539, 76923.07692307692)
    Kref = fo.assemble_stiffness(pts, el, "tet4", fo.elastic_isotropic_H(200e3, 0.3), 3)
    assert np.abs(K - Kref.data).max() <= 1e-12 * np.abs(Kref.data).max()


def test_slab_ranks_get_the_same_bricks_as_one_gpu():
    """The Morton curve is built on the lattice of the OWNED nodes with per-axis bucket counts following the bounding
    box: the z-slabs of a multi-GPU partition (structured or jittered) are cut into the same 4x4x2-node bricks as the
    whole box, so the total number of clusters does not grow with the number of ranks."""
    from fedoo_b200 import dist as fdist

    n = 24  # 25^3 nodes
    totals = {}
    for jitter in (False, True):
        for world in (1, 2, 4):
            tot = 0
            for rank in range(world):
                loc = fdist.box_local_slab(n, rank, world, jitter=jitter)
                owned = torch.from_numpy(loc.owned) if world > 1 else None
                plan, _ = make_plan("hex8", loc.nodes, loc.elements, owned=owned, small=False)
                tot += plan.n_clusters
            totals[(jitter, world)] = tot
    one = totals[(False, 1)]
    assert totals[(True, 1)] == one
    for key, tot in totals.items():
        assert tot <= 1.08 * one, (key, tot, one)  # slab faces may add a layer of partial bricks, nothing more
