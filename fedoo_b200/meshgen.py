"""Synthetic mesh generators for the assembly benchmarks and parity tests.

The numbering conventions follow the reference generators so that a mesh made
here is identical (node for node, element for element) to the one the
reference would build:

* ``box_hex8``: nodes x-fastest, then y, then z; hex8 connectivity
  ``[n, n+1, n+nx+1, n+nx]`` + the same on the next z-plane, elements looped
  k (slowest), j, i -- ``fedoo/mesh/simple.py:500-505,752-767``.
* tet4 / tet10 local node order as ``fedoo/lib_elements/tetrahedron.py:112-114,
  139-143`` (tet10 mid-edge nodes on edges (0-1),(1-2),(0-2),(0-3),(1-3),(2-3)).

Everything is vectorised NumPy (the reference's hex8 generator is a Python list
comprehension: 3.9 s at 1 M elements).
"""

from __future__ import annotations

import numpy as np


def box_hex8(nx=11, ny=11, nz=11, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0, z_min=0.0, z_max=1.0):
    """Structured hex8 box; nx, ny, nz are NODE counts (as in fd.mesh.box_mesh)."""
    Y, Z, X = np.meshgrid(np.linspace(y_min, y_max, ny), np.linspace(z_min, z_max, nz), np.linspace(x_min, x_max, nx))
    nodes = np.stack([X.reshape(-1), Y.reshape(-1), Z.reshape(-1)], axis=1)
    i = np.arange(nx - 1, dtype=np.int64)[None, None, :]
    j = np.arange(ny - 1, dtype=np.int64)[None, :, None]
    k = np.arange(nz - 1, dtype=np.int64)[:, None, None]
    n0 = (nx * j + i + k * nx * ny).reshape(-1)
    off = np.array([0, 1, nx + 1, nx, nx * ny, nx * ny + 1, nx * ny + nx + 1, nx * ny + nx], dtype=np.int64)
    elements = n0[:, None] + off[None, :]
    return nodes, elements


def box_node_sets(nx, ny, nz):
    """left/right/bottom/top/front/back node sets of ``box_hex8`` (simple.py:769-796)."""
    idx = np.arange(nx * ny * nz).reshape(nz, ny, nx)
    return {
        "left": idx[:, :, 0].reshape(-1),
        "right": idx[:, :, -1].reshape(-1),
        "front": idx[:, 0, :].reshape(-1),
        "back": idx[:, -1, :].reshape(-1),
        "bottom": idx[0].reshape(-1),
        "top": idx[-1].reshape(-1),
    }


def jitter_nodes(nodes, nx, ny, nz, amplitude=0.2, seed=1):
    """Displace interior nodes of a structured box by U(-a h, a h) per axis."""
    nodes = np.array(nodes, dtype=float, copy=True)
    rng = np.random.default_rng(seed)
    idx = np.arange(nx * ny * nz).reshape(nz, ny, nx)
    interior = idx[1:-1, 1:-1, 1:-1].reshape(-1)
    ext = nodes.max(axis=0) - nodes.min(axis=0)
    h = ext / np.array([nx - 1, ny - 1, nz - 1])
    nodes[interior] += rng.uniform(-amplitude, amplitude, size=(interior.size, 3)) * h
    return nodes


def hex8_to_tet4(elements):
    """Split every hex8 into 6 tet4 sharing the 0-6 diagonal (conforming on structured boxes)."""
    e = np.asarray(elements)
    pat = np.array([[0, 1, 2, 6], [0, 2, 3, 6], [0, 3, 7, 6], [0, 7, 4, 6], [0, 4, 5, 6], [0, 5, 1, 6]])
    return e[:, pat].reshape(-1, 4)


_TET_EDGES = np.array([[0, 1], [1, 2], [0, 2], [0, 3], [1, 3], [2, 3]])


def tet4_to_tet10(nodes, elements, bulge=0.0, seed=3):
    """Add mid-edge nodes (fedoo order).  ``bulge`` > 0 displaces them randomly
    by ``bulge`` x edge length so that elements have curved edges (non-constant Jacobian)."""
    nodes = np.asarray(nodes, dtype=float)
    e = np.asarray(elements, dtype=np.int64)
    n_nodes = len(nodes)
    pairs = np.sort(e[:, _TET_EDGES].reshape(-1, 2), axis=1)
    key = pairs[:, 0] * n_nodes + pairs[:, 1]
    uniq, inv = np.unique(key, return_inverse=True)
    a, b = uniq // n_nodes, uniq % n_nodes
    mid = 0.5 * (nodes[a] + nodes[b])
    if bulge:
        rng = np.random.default_rng(seed)
        length = np.linalg.norm(nodes[a] - nodes[b], axis=1)
        mid += rng.uniform(-bulge, bulge, size=mid.shape) * length[:, None]
    new_nodes = np.vstack([nodes, mid])
    new_elements = np.hstack([e, n_nodes + inv.reshape(-1, 6)])
    return new_nodes, new_elements


def rect_quad4(nx=11, ny=11, x_min=0.0, x_max=1.0, y_min=0.0, y_max=1.0):
    """Structured quad4 rectangle, nodes x-fastest, counter-clockwise connectivity."""
    Y, X = np.meshgrid(np.linspace(y_min, y_max, ny), np.linspace(x_min, x_max, nx), indexing="ij")
    nodes = np.stack([X.reshape(-1), Y.reshape(-1)], axis=1)
    i = np.arange(nx - 1, dtype=np.int64)[None, :]
    j = np.arange(ny - 1, dtype=np.int64)[:, None]
    n0 = (nx * j + i).reshape(-1)
    elements = n0[:, None] + np.array([0, 1, nx + 1, nx], dtype=np.int64)[None, :]
    return nodes, elements


def jitter_nodes_2d(nodes, nx, ny, amplitude=0.2, seed=1):
    nodes = np.array(nodes, dtype=float, copy=True)
    rng = np.random.default_rng(seed)
    idx = np.arange(nx * ny).reshape(ny, nx)
    interior = idx[1:-1, 1:-1].reshape(-1)
    ext = nodes.max(axis=0) - nodes.min(axis=0)
    h = ext / np.array([nx - 1, ny - 1])
    nodes[interior] += rng.uniform(-amplitude, amplitude, size=(interior.size, 2)) * h
    return nodes


def hole_plate_quad4(nr=11, nt=11, length=100.0, height=100.0, radius=20.0):
    """Plate [-length/2, length/2] x [-height/2, height/2] with a central circular hole, as a ring of
    8 (nr-1)(nt-1) quad4 (same counts as fd.mesh.hole_plate_mesh, fedoo/mesh/structured_mesh.py:247: nr
    nodes from the hole to the outer edge, nt nodes on each half of an outer edge).  Nodes (t, r) with t
    around the hole (periodic, 8 (nt-1) columns) and r outwards: a straight blend of the circle point at
    the angle of the outer-edge point and that point."""
    L, H = 0.5 * length, 0.5 * height
    s = np.linspace(0.0, 1.0, nt)[:-1]  # along one half edge
    # outer boundary, counter-clockwise from (L, 0): 8 half edges
    corners = np.array([[L, 0], [L, H], [0, H], [-L, H], [-L, 0], [-L, -H], [0, -H], [L, -H], [L, 0]], dtype=float)
    outer = np.concatenate([corners[k] + s[:, None] * (corners[k + 1] - corners[k]) for k in range(8)])
    theta = np.arctan2(outer[:, 1], outer[:, 0])
    inner = radius * np.stack([np.cos(theta), np.sin(theta)], axis=1)
    r = np.linspace(0.0, 1.0, nr)
    nodes = (inner[:, None, :] + r[None, :, None] * (outer - inner)[:, None, :]).reshape(-1, 2)  # index t*nr + r
    n_t = len(outer)
    t = np.arange(n_t, dtype=np.int64)[:, None]
    k = np.arange(nr - 1, dtype=np.int64)[None, :]
    tn = (t + 1) % n_t
    elements = np.stack([t * nr + k, t * nr + k + 1, tn * nr + k + 1, tn * nr + k], axis=2).reshape(-1, 4)
    return nodes, elements


def extrude_quad4_to_hex8(nodes2d, quads, thickness, n_layers):
    """Extrude a quad4 mesh along z into n_layers element layers of hex8 (node layer k at z = k thickness / n_layers;
    hex8 local order = the quad at the lower layer, then at the upper one, fedoo/mesh/functions.py:272)."""
    nodes2d = np.asarray(nodes2d, dtype=float)
    quads = np.asarray(quads, dtype=np.int64)
    n2 = len(nodes2d)
    z = np.linspace(0.0, thickness, n_layers + 1)
    nodes = np.concatenate([np.tile(nodes2d, (n_layers + 1, 1)), np.repeat(z, n2)[:, None]], axis=1)
    lay = np.arange(n_layers, dtype=np.int64)[:, None, None] * n2
    elements = np.concatenate([quads[None] + lay, quads[None] + lay + n2], axis=2).reshape(-1, 8)
    return nodes, elements
