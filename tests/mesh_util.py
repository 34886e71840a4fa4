"""Unstructured meshes for the tests: Delaunay tetrahedra of random points (irregular valence, some of it beyond a
cluster's capacity -- the rows kernel's share), slivers dropped, orientation fixed."""

import numpy as np


def delaunay_tet4(n_points, seed):
    from scipy.spatial import Delaunay

    pts = np.random.default_rng(seed).uniform(0, 1, (n_points, 3))
    el = Delaunay(pts).simplices.astype(np.int64)
    X = pts[el]
    vol = np.einsum("ij,ij->i", np.cross(X[:, 1] - X[:, 0], X[:, 2] - X[:, 0]), X[:, 3] - X[:, 0]) / 6
    neg = vol < 0
    el[neg] = el[neg][:, [0, 2, 1, 3]]
    return pts, el[np.abs(vol) > 0.02 * np.abs(vol).mean()]
