"""NumPy/SciPy restatement of fedoo's global-operator assembly path (CPU oracle).

TEST INFRASTRUCTURE ONLY -- see ``oracle/__init__.py``.  The product package
``fedoo_b200`` never imports this module.

Every function restates, in plain NumPy, the algorithm of the reference
(3MAH/fedoo v0.7.0) and cites the reference file:line it follows (paths are
relative to the reference checkout).  The algorithmic *structure* of the
reference is kept on purpose (per-derivative operators -> batched A^T diag(c) B
per element -> COO->CSR "convert" gather -> nvar x nvar block tiling), so that
timing this module is a fair "port" CPU baseline of the reference path.

Parity status:
  * elasticity (hex8/tet4/tet10/quad4), per-GP tangent, residual, strain/stress
    and heat equation: PINNED against arrays produced by the reference itself
    (``oracle/gen_golden.py`` -> ``tests/golden/*.npz``; checked by
    ``tests/test_oracle_golden.py``).
  * J2 plasticity (``j2_radial_return``): the STATE UPDATE (sigma, p, eps_p) is PINNED on the reference's own
    per-Gauss-point loop -- constitutivelaw/elasto_plasticity.py:303-376, dead code at this commit, driven standalone
    by ``oracle/gen_golden_j2.py`` with its three breakages shimmed at run time (renamed tensor methods, hardening-slope
    sign, local tolerance) -> ``tests/golden/j2_reference.npz``, agreement 5e-13.  The TANGENT is a definition, not a
    result: the consistent tangent is checked by finite differences of sigma(eps); the continuum tangent
    (``j2_continuum_tangent``) follows :275-300 with the corrected sign.  The living path's arithmetic is simcoon's
    (C++, absent): the reference's one J2 known-answer test (tests/test_octet.py:80-81) is reproduced by the
    REFERENCE'S OWN driver + this law (``oracle/gen_golden_octet_driver.py``) to 1.4e-3 / 2.7e-4 of its published
    numbers, outside the test's tolerance -- that residual is simcoon's cutting-plane iteration at freshly yielded
    points (one Newton correction per increment makes the answers tangent-dependent), and stays UNPINNED.

Conventions (reference): global dof = var * n_nodes + node
(core/problem.py:89-91); Gauss-point index = gp * n_elements + element
(core/mesh.py:1137-1140); Voigt order [xx, yy, zz, xy, xz, yz] with engineering
shear strains (core/modelingspace.py:289-316).
"""

from __future__ import annotations

import numpy as np
from scipy import sparse

# ---------------------------------------------------------------------------
# Element tables  (lib_elements/hexahedron.py:178-248, tetrahedron.py:106-208,
# quadrangle.py:125-167, Gauss rules hexahedron.py:22-27,134-135,
# tetrahedron.py:21-61,72-97, quadrangle.py:24-26)
# ---------------------------------------------------------------------------


class ElementTable:
    """Shape-function values/derivatives at the Gauss points of one element type."""

    def __init__(self, name, xi_gp, w_gp, shape, dshape):
        self.name = name
        self.xi_gp = np.asarray(xi_gp, dtype=float)
        self.w_gp = np.asarray(w_gp, dtype=float)
        self.ngp = len(self.w_gp)
        self.dim = self.xi_gp.shape[1]
        self.N = np.array([shape(x) for x in self.xi_gp])  # (ngp, nne)
        self.dN = np.array([dshape(x) for x in self.xi_gp])  # (ngp, dim, nne)
        self.nne = self.N.shape[1]


def _hex8():
    a = 0.5773502691896258
    xi_gp = [[sx * a, sy * a, sz * a] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)]
    nd = np.array(
        [[-1, -1, -1], [1, -1, -1], [1, 1, -1], [-1, 1, -1], [-1, -1, 1], [1, -1, 1], [1, 1, 1], [-1, 1, 1]],
        dtype=float,
    )

    def shape(x):
        return 0.125 * (1 + nd[:, 0] * x[0]) * (1 + nd[:, 1] * x[1]) * (1 + nd[:, 2] * x[2])

    def dshape(x):
        f = [1 + nd[:, d] * x[d] for d in range(3)]
        return np.array(
            [0.125 * nd[:, 0] * f[1] * f[2], 0.125 * nd[:, 1] * f[0] * f[2], 0.125 * nd[:, 2] * f[0] * f[1]]
        )

    return ElementTable("hex8", xi_gp, np.ones(8), shape, dshape)


def _quad4():
    a = 1 / np.sqrt(3)
    xi_gp = [[-a, -a], [a, -a], [a, a], [-a, a]]
    nd = np.array([[-1, -1], [1, -1], [1, 1], [-1, 1]], dtype=float)

    def shape(x):
        return 0.25 * (1 + nd[:, 0] * x[0]) * (1 + nd[:, 1] * x[1])

    def dshape(x):
        return np.array([0.25 * nd[:, 0] * (1 + nd[:, 1] * x[1]), 0.25 * nd[:, 1] * (1 + nd[:, 0] * x[0])])

    return ElementTable("quad4", xi_gp, np.ones(4), shape, dshape)


def _tet4():
    a = 0.1381966011250105
    b = 0.5854101966249685
    xi_gp = [[a, a, a], [a, a, b], [a, b, a], [b, a, a]]

    def shape(x):
        return np.array([x[1], x[2], 1 - x[0] - x[1] - x[2], x[0]])

    def dshape(x):
        return np.array([[0.0, 0.0, -1.0, 1.0], [1.0, 0.0, -1.0, 0.0], [0.0, 1.0, -1.0, 0.0]])

    return ElementTable("tet4", xi_gp, np.full(4, 1 / 24), shape, dshape)


def _tet10():
    a = 0.25
    b1 = 0.3197936278296299
    b2 = 0.09197107805272303
    c1 = 0.040619116511110234
    c2 = 0.724086765841831
    d = 0.05635083268962915
    e = 0.4436491673103708
    xi_gp = [
        [a, a, a],
        [b1, b1, b1], [b1, b1, c1], [b1, c1, b1], [c1, b1, b1],
        [b2, b2, b2], [b2, b2, c2], [b2, c2, b2], [c2, b2, b2],
        [d, d, e], [d, e, d], [e, d, d], [d, e, e], [e, d, e], [e, e, d],
    ]  # fmt: skip
    f1 = 0.011511367871045397
    f2 = 0.01198951396316977
    w = [8.0 / 405] + [f1] * 4 + [f2] * 4 + [5.0 / 567] * 6

    def shape(x):
        xi, eta, zeta = x
        m = 1 - xi - eta - zeta
        return np.array(
            [
                eta * (2 * eta - 1), zeta * (2 * zeta - 1), m * (1 - 2 * xi - 2 * eta - 2 * zeta),
                xi * (2 * xi - 1), 4 * eta * zeta, 4 * zeta * m, 4 * eta * m, 4 * xi * eta,
                4 * xi * zeta, 4 * xi * m,
            ]
        )  # fmt: skip

    def dshape(x):
        xi, eta, zeta = x
        m = 1 - xi - eta - zeta
        return np.array(
            [
                [0.0, 0.0, 1 - 4 * m, -1 + 4 * xi, 0.0, -4 * zeta, -4 * eta, 4 * eta, 4 * zeta, 4 * (m - xi)],
                [-1 + 4 * eta, 0.0, 1 - 4 * m, 0.0, 4 * zeta, -4 * zeta, 4 * (m - eta), 4 * xi, 0.0, -4 * xi],
                [0.0, -1 + 4 * zeta, 1 - 4 * m, 0.0, 4 * eta, 4 * (m - zeta), -4 * eta, 0.0, 4 * xi, -4 * xi],
            ]
        )

    return ElementTable("tet10", xi_gp, w, shape, dshape)


_TABLES = {}


def element_table(name: str) -> ElementTable:
    if name not in _TABLES:
        _TABLES[name] = {"hex8": _hex8, "quad4": _quad4, "tet4": _tet4, "tet10": _tet10}[name]()
    return _TABLES[name]


# ---------------------------------------------------------------------------
# Geometry at Gauss points
# ---------------------------------------------------------------------------


def geometry(nodes, elements, elm_type):
    """Jacobian, |det J| w and dN/dx at every Gauss point.

    Follows lib_elements/element_base.py:13-124 (J[el,gp] = dN/dxi(gp) . X_el,
    rows = xi direction, cols = x direction; detJ = abs(det J); J^-1 by LAPACK)
    and core/assembly.py:879-881 (dN/dx = J^-1 . dN/dxi).

    Returns G (n_el, ngp, dim, nne) and wdet (n_el, ngp).
    """
    tab = element_table(elm_type)
    X = np.asarray(nodes, dtype=float)[np.asarray(elements)]  # (n_el, nne, dim)
    J = np.einsum("gdk,ekx->egdx", tab.dN, X)
    detJ = np.abs(np.linalg.det(J))
    invJ = np.linalg.inv(J)  # [x, xi]
    G = np.einsum("egxd,gdk->egxk", invJ, tab.dN)
    wdet = detJ * tab.w_gp[None, :]
    return G, wdet


def quadrature_vector(wdet):
    """gp-major quadrature vector (core/mesh.py:1196: (detJ*w).T.reshape(-1))."""
    return np.ascontiguousarray(wdet.T).reshape(-1)


# ---------------------------------------------------------------------------
# Symbolic CSR pattern  (core/_sparsematrix.py:148-174, 225-284, 310-315)
# ---------------------------------------------------------------------------


class Pattern:
    """Block (node-node) pattern + its nvar x nvar tiling and the COO->CSR gather."""

    def __init__(self, elements, n_nodes, nvar, block_mask=None, n_global_dof=0):
        elements = np.asarray(elements)
        n_el, nne = elements.shape
        self.n_el, self.nne, self.n_nodes, self.nvar = n_el, nne, n_nodes, nvar
        # element COO rows/cols, (n_el, nne, nne) ravelled  (_sparsematrix.py:162-173)
        row = np.repeat(elements, nne, axis=1).reshape(-1).astype(np.int32)
        col = np.tile(elements, (1, nne)).reshape(-1).astype(np.int32)
        ref = row.astype(np.int64) * n_nodes + col  # (_sparsematrix.py:257-259)
        order = np.argsort(ref, kind="stable")
        val, ind_unique, count = np.unique(ref, return_index=True, return_counts=True)
        self.gather_indices = order  # convert-matrix CSR indices (:260-274)
        self.gather_indptr = np.concatenate(([0], np.cumsum(count))).astype(np.int64)
        self.blk_indices = col[ind_unique].astype(np.int32)  # (:277)
        nb = np.bincount(row[ind_unique], minlength=n_nodes)  # (:279-284)
        self.blk_indptr = np.concatenate(([0], np.cumsum(nb))).astype(np.int64)
        self.blk_nnz = len(val)
        if block_mask is None:
            block_mask = np.ones((nvar, nvar), dtype=bool)
        self.block_mask = np.asarray(block_mask, dtype=bool)
        # bmat tiling (:310-315; scipy _construct.py:1009-1100).  Row v*n+I is the
        # concatenation over v' (present blocks) of v'*n + blockrow(I).
        deg = np.diff(self.blk_indptr)
        per_row = np.concatenate([deg * int(self.block_mask[v].sum()) for v in range(nvar)])
        n_rows = nvar * n_nodes + n_global_dof
        indptr = np.zeros(n_rows + 1, dtype=np.int64)
        np.cumsum(per_row, out=indptr[1 : nvar * n_nodes + 1])
        indptr[nvar * n_nodes + 1 :] = indptr[nvar * n_nodes]
        self.nnz = int(indptr[-1])
        # index dtype rule (scipy get_index_dtype: int32 iff max(nnz, dim) <= 2^31-1)
        self.index_dtype = np.int32 if max(self.nnz, n_rows) <= np.iinfo(np.int32).max else np.int64
        self.indptr = indptr.astype(self.index_dtype)
        self.shape = (n_rows, n_rows)
        self._indices = None

    @property
    def indices(self):
        if self._indices is None:
            n, nvar = self.n_nodes, self.nvar
            out = np.empty(self.nnz, dtype=self.index_dtype)
            deg = np.diff(self.blk_indptr)
            for v in range(nvar):
                cols = [vp for vp in range(nvar) if self.block_mask[v, vp]]
                if not cols:
                    continue
                base = self.indptr[v * n : (v + 1) * n].astype(np.int64)
                for q, vp in enumerate(cols):
                    # destination of block-row entries: base[I] + q*deg[I] + p
                    dst = np.repeat(base + q * deg, deg) + (
                        np.arange(self.blk_nnz) - np.repeat(self.blk_indptr[:-1], deg)
                    )
                    out[dst] = self.blk_indices + vp * n
            self._indices = out
        return self._indices

    def block_values(self, data):
        """vals = Matrix_convertCOOtoCSR @ data.ravel()  (_sparsematrix.py:302)."""
        d = np.asarray(data).reshape(-1)[self.gather_indices]
        return np.add.reduceat(d, self.gather_indptr[:-1])

    def tile_values(self, blocks):
        """Stack nvar x nvar block value arrays into the global data array (bmat)."""
        n, nvar = self.n_nodes, self.nvar
        out = np.zeros(self.nnz)
        deg = np.diff(self.blk_indptr)
        local = np.arange(self.blk_nnz) - np.repeat(self.blk_indptr[:-1], deg)
        for v in range(nvar):
            cols = [vp for vp in range(nvar) if self.block_mask[v, vp]]
            base = self.indptr[v * n : (v + 1) * n].astype(np.int64)
            for q, vp in enumerate(cols):
                dst = np.repeat(base + q * deg, deg) + local
                out[dst] = blocks[v][vp]
        return out

    def csr(self, data):
        return sparse.csr_matrix((data, self.indices, self.indptr), shape=self.shape)


# ---------------------------------------------------------------------------
# Constitutive matrices
# ---------------------------------------------------------------------------


def elastic_isotropic_H(E, nu, dimension="3D"):
    """6x6 Hooke matrix (constitutivelaw/elastic_isotrop.py:35-68)."""
    H = np.zeros((6, 6))
    if dimension == "2Dstress":
        H[0, 0] = H[1, 1] = E / (1 - nu**2)
        H[0, 1] = H[1, 0] = nu * E / (1 - nu**2)
        H[3, 3] = 0.5 * E / (1 + nu)
    else:
        H[0, 0] = H[1, 1] = H[2, 2] = E * (1.0 / (1 + nu) + nu / ((1.0 + nu) * (1 - 2 * nu)))
        H[0, 1] = H[0, 2] = H[1, 2] = E * (nu / ((1 + nu) * (1 - 2 * nu)))
        H[3, 3] = H[4, 4] = H[5, 5] = 0.5 * E / (1 + nu)
        H[1, 0] = H[0, 1]
        H[2, 0] = H[0, 2]
        H[2, 1] = H[1, 2]
    return H


def plane_stress_H(H):
    """core/mechanical3d.py:47-69."""
    H = np.asarray(H)
    out = np.zeros_like(H)
    for i in (0, 1, 3):
        for j in (0, 1, 3):
            out[i, j] = H[i, j] - H[i, 2] * H[j, 2] / H[2, 2]
    return out


def strain_ops(ndim):
    """For each Voigt strain component, the list of (variable, derivative direction).

    core/modelingspace.py:289-316.
    """
    if ndim == 2:
        return [[(0, 0)], [(1, 1)], [], [(0, 1), (1, 0)], [], []]
    return [[(0, 0)], [(1, 1)], [(2, 2)], [(0, 1), (1, 0)], [(0, 2), (2, 0)], [(1, 2), (2, 1)]]


# ---------------------------------------------------------------------------
# Numeric assembly: K  (core/assembly.py:282-460, core/_sparsematrix.py:55-174,286-315)
# ---------------------------------------------------------------------------


def _gp_major(H_gp, n_el, ngp):
    """(.., N) gp-major array -> (.., n_el, ngp)."""
    return np.moveaxis(H_gp.reshape(H_gp.shape[:-1] + (ngp, n_el)), -1, -2)


def stiffness_blocks(G, wdet, H, ndim, assume_sym=True):
    """Per-element block data[v][v'] (n_el, nne, nne) = sum of A^T diag(coef) B terms.

    One batched matmul per weak-form term, exactly as _BlocSparse.addToBlocATB
    (core/_sparsematrix.py:83-89) driven by the term list of
    StressEquilibrium.get_weak_equation (weakform/stress_equilibrium.py:92-145):
    sigma_i = sum_j eps_j H_ij (zero H_ij dropped when H is scalar valued,
    core/diffop.py:177-181), weak form sum_i eps_i(v) sigma_i.  With assume_sym
    only blocks v <= v' are computed and the others are transposes
    (core/assembly.py:313-318, _sparsematrix.py:287-296).
    """
    n_el, ngp = wdet.shape
    nvar = ndim
    ops = strain_ops(ndim)
    H = np.asarray(H)
    per_gp = H.ndim == 3
    if per_gp:
        Hg = _gp_major(H, n_el, ngp)  # (6, 6, n_el, ngp)
    data = [[None] * nvar for _ in range(nvar)]
    for i in range(6):
        for j in range(6):
            if not per_gp and H[i, j] == 0:
                continue
            coef = wdet * (Hg[i, j] if per_gp else H[i, j])  # (n_el, ngp)
            for vvir, dvir in ops[i]:
                for v, d in ops[j]:
                    if assume_sym and vvir > v:
                        continue
                    A = G[:, :, dvir, :]  # (n_el, ngp, nne)
                    B = G[:, :, d, :]
                    new = np.matmul((coef[:, :, None] * A).transpose(0, 2, 1), B)
                    if data[vvir][v] is None:
                        data[vvir][v] = new
                    else:
                        data[vvir][v] += new
    if assume_sym:
        for a in range(nvar):
            for b in range(a):
                data[a][b] = data[b][a].transpose(0, 2, 1)
    return data


def assemble_stiffness(nodes, elements, elm_type, H, ndim, pattern=None, geom=None):
    """Global K (scipy CSR, reference pattern) for StressEquilibrium, small strain."""
    nodes = np.asarray(nodes, dtype=float)
    if pattern is None:
        pattern = Pattern(elements, len(nodes), ndim)
    G, wdet = geom if geom is not None else geometry(nodes, elements, elm_type)
    data = stiffness_blocks(G, wdet, H, ndim)
    blocks = [[pattern.block_values(data[a][b]) for b in range(ndim)] for a in range(ndim)]
    return pattern.csr(pattern.tile_values(blocks))


# ---------------------------------------------------------------------------
# State update: grad u, strain, stress  (core/assembly.py:1045-1112,1285-1336;
# weakform/stress_equilibrium.py:485-524,589-601; elastic_anisotropic.py:36-56)
# ---------------------------------------------------------------------------


def grad_disp_gp(G, elements, U, n_nodes, ndim):
    """grad[a][b] = d u_a / d x_b at every GP, each (N,) gp-major."""
    Ue = np.stack([U[a * n_nodes + np.asarray(elements)] for a in range(ndim)])  # (ndim, n_el, nne)
    g = np.einsum("egbk,aek->abge", G, Ue)  # (ndim, ndim, ngp, n_el)
    return g.reshape(ndim, ndim, -1)


def strain_gp(G, elements, U, n_nodes, ndim):
    """(6, N) Voigt strain, engineering shears, gp-major columns."""
    g = grad_disp_gp(G, elements, U, n_nodes, ndim)
    N = g.shape[-1]
    eps = np.zeros((6, N))
    eps[0], eps[1] = g[0, 0], g[1, 1]
    eps[3] = g[0, 1] + g[1, 0]
    if ndim == 3:
        eps[2] = g[2, 2]
        eps[4] = g[0, 2] + g[2, 0]
        eps[5] = g[1, 2] + g[2, 1]
    return eps


def deformation_gradient(G, elements, U, n_nodes, ndim, fbar=False):
    """F = 1 + grad u at the Gauss points, (3, 3, N) gp-major; with ``fbar`` scaled by (J_mean / J)^(1/3), J_mean the
    mean of det F over the element's Gauss points (weakform/stress_equilibrium.py:542-586, _comp_F / _comp_Fbar)."""
    g = grad_disp_gp(G, elements, U, n_nodes, ndim)  # (ndim, ndim, N)
    N = g.shape[-1]
    F = np.zeros((3, 3, N))
    F[np.arange(3), np.arange(3)] = 1.0
    F[:ndim, :ndim] += g
    if fbar:
        n_el = len(elements)
        J = np.linalg.det(F.transpose(2, 0, 1))
        Jc = np.mean(J.reshape(-1, n_el), axis=0)
        F = F * ((Jc / J.reshape(-1, n_el)).ravel() ** (1 / 3))
    return F


def stress_gp(H, eps):
    """sigma_i = sum_j eps_j H_ij, H 6x6 or (6,6,N)."""
    H = np.asarray(H)
    if H.ndim == 3:
        return np.einsum("ijn,jn->in", H, eps)
    return H @ eps


# ---------------------------------------------------------------------------
# Residual  D = - int B^T sigma   (core/assembly.py:400-411)
# ---------------------------------------------------------------------------


def residual(G, wdet, elements, sigma, n_nodes, ndim, n_global_dof=0):
    n_el, ngp = wdet.shape
    ops = strain_ops(ndim)
    sig = sigma.reshape(6, ngp, n_el).transpose(0, 2, 1)  # (6, n_el, ngp)
    D = np.zeros(ndim * n_nodes + n_global_dof)
    elements = np.asarray(elements)
    for i in range(6):
        coef = sig[i] * wdet  # (n_el, ngp)
        for vvir, dvir in ops[i]:
            contrib = np.einsum("egk,eg->ek", G[:, :, dvir, :], coef)
            np.subtract.at(D, vvir * n_nodes + elements.reshape(-1), contrib.reshape(-1))
    return D


# ---------------------------------------------------------------------------
# Heat equation  (weakform/heat_equation.py:78-119,168-187,194-227;
# lumping: core/_sparsematrix.py:91-98)
# ---------------------------------------------------------------------------


def assemble_heat(nodes, elements, elm_type, conductivity, rho_c, dtime, pattern=None, geom=None):
    """K = int grad v . k grad T + lumped (rho c / dt) int v T ; one variable (Temp)."""
    nodes = np.asarray(nodes, dtype=float)
    tab = element_table(elm_type)
    if pattern is None:
        pattern = Pattern(elements, len(nodes), 1)
    G, wdet = geom if geom is not None else geometry(nodes, elements, elm_type)
    k = np.asarray(conductivity, dtype=float)
    if k.ndim == 0:
        k = np.eye(3) * float(k)
    data = None
    for i in range(tab.dim):
        for j in range(tab.dim):
            if k[i, j] == 0:
                continue
            new = np.matmul((wdet[:, :, None] * k[i, j] * G[:, :, i, :]).transpose(0, 2, 1), G[:, :, j, :])
            data = new if data is None else data + new
    if dtime != 0:
        Ngp = np.broadcast_to(tab.N[None], (len(elements),) + tab.N.shape)  # (n_el, ngp, nne)
        mass = np.matmul((wdet[:, :, None] * (rho_c / dtime) * Ngp).transpose(0, 2, 1), Ngp)
        idx = np.arange(tab.nne)
        data[:, idx, idx] += mass.sum(axis=2)
    return pattern.csr(pattern.tile_values([[pattern.block_values(data)]]))


def temp_gp(elements, elm_type, T):
    """Node -> GP interpolation (core/mesh.py:1162-1172), gp-major (N,)."""
    tab = element_table(elm_type)
    Te = np.asarray(T)[np.asarray(elements)]  # (n_el, nne)
    return np.einsum("gk,ek->ge", tab.N, Te).reshape(-1)


def temp_gradient_gp(G, elements, T):
    """list of dim arrays (N,)  (weakform/heat_equation.py:64-70)."""
    Te = np.asarray(T)[np.asarray(elements)]
    g = np.einsum("egdk,ek->dge", G, Te)
    return g.reshape(g.shape[0], -1)


def residual_heat(G, wdet, elements, elm_type, conductivity, rho_c, dtime, T, T_start, n_nodes):
    """D_I = - sum_g w [ grad N_I . k grad T + (rho c/dt) N_I (T_g - Tstart_g) ]."""
    tab = element_table(elm_type)
    n_el, ngp = wdet.shape
    k = np.asarray(conductivity, dtype=float)
    if k.ndim == 0:
        k = np.eye(3) * float(k)
    k = k[: tab.dim, : tab.dim]
    gT = temp_gradient_gp(G, elements, T).reshape(tab.dim, ngp, n_el).transpose(0, 2, 1)  # (dim, n_el, ngp)
    q = np.einsum("ij,jeg->ieg", k, gT)
    contrib = np.einsum("egik,ieg,eg->ek", G, q, wdet)
    if dtime != 0:
        dT = (temp_gp(elements, elm_type, T) - temp_gp(elements, elm_type, T_start)).reshape(ngp, n_el).T
        contrib = contrib + np.einsum("gk,eg,eg->ek", tab.N, dT, wdet) * (rho_c / dtime)
    D = np.zeros(n_nodes)
    np.subtract.at(D, np.asarray(elements).reshape(-1), contrib.reshape(-1))
    return D


# ---------------------------------------------------------------------------
# J2 plasticity with isotropic power-law hardening -- state update pinned on the reference's loop (header)
# (constitutivelaw/elasto_plasticity.py:66-80 elastic H, :127-133 hardening,
#  :154-164 yield function / flow direction, :303-376 trial state + return,
#  :275-300 tangent; sv protocol constitutivelaw/simcoon_umat.py:463-580, EPICP
#  props [E, nu, alpha, sigmaY, k, m] and statev [T, p, EP(6)] :103-113)
# ---------------------------------------------------------------------------

_VOIGT_W = np.array([1.0, 1.0, 1.0, 2.0, 2.0, 2.0])  # s:s weights in Voigt stress space


def j2_radial_return(eps, statev_start, props, tol=1e-12, max_iter=50):
    """Backward-Euler radial return for J2 + R(p) = k p^m, small strain.

    eps          (6, N) total strain (engineering shears)
    statev_start (8, N) [T, p, EP_xx, EP_yy, EP_zz, EP_xy, EP_xz, EP_yz] (EP shears engineering)
    props        [E, nu, alpha, sigmaY, k, m]
    Returns stress (6, N), statev (8, N), tangent (6, 6, N) (consistent tangent).

    The legacy file iterates a cutting-plane loop per Gauss point; for J2 with
    isotropic hardening the flow direction is constant along the return so the
    fixed point equals the scalar radial-return root of
        f(dp) = q_trial - 3 mu dp - sigmaY - k (p0 + dp)^m = 0,
    solved here by Newton with slope -(3 mu + R'(p)) (SURVEY 8c notes the legacy
    slope sign is wrong for tiny p; the converged sigma, p, EP are what is compared).
    """
    E, nu, _alpha, sigY, k, m = [float(x) for x in props]
    mu = 0.5 * E / (1 + nu)
    H = elastic_isotropic_H(E, nu)
    N = eps.shape[1]
    p0 = statev_start[1].copy()
    ep0 = statev_start[2:8]
    sig_tr = H @ (eps - ep0)
    pm = sig_tr[:3].sum(axis=0) / 3.0
    s = sig_tr.copy()
    s[:3] -= pm
    q = np.sqrt(1.5 * (_VOIGT_W[:, None] * s * s).sum(axis=0))

    def R(p):
        return k * np.power(np.maximum(p, 0.0), m)

    def dR(p):
        with np.errstate(divide="ignore", invalid="ignore"):
            v = k * m * np.power(np.maximum(p, 0.0), m - 1.0)
        return np.nan_to_num(v, nan=0.0, posinf=1e300)

    f_tr = q - sigY - R(p0)
    plastic = f_tr > 0
    dp = np.zeros(N)
    idx = np.where(plastic)[0]
    if idx.size:
        qq, pp = q[idx], p0[idx]
        x = np.maximum(f_tr[idx] / (3 * mu), 1e-300)  # upper bound of the root (R increasing)
        lo = np.zeros_like(x)
        hi = x.copy()
        for _ in range(max_iter):
            fx = qq - 3 * mu * x - sigY - R(pp + x)
            hi = np.where(fx < 0, np.minimum(hi, x), hi)
            lo = np.where(fx > 0, np.maximum(lo, x), lo)
            slope = 3 * mu + dR(pp + x)
            xn = x + fx / slope
            bad = ~((xn > lo) & (xn < hi)) | ~np.isfinite(xn)
            xn = np.where(bad, 0.5 * (lo + hi), xn)
            done = np.abs(xn - x) <= tol * np.maximum(np.abs(xn), 1e-300)
            x = xn
            if done.all():
                break
        dp[idx] = x
    qs = np.where(q > 0, q, 1.0)
    nflow = 1.5 * s / qs  # d q / d sigma in Voigt stress space (normal, "n")
    sig = sig_tr - 2 * mu * dp * nflow
    statev = statev_start.copy()
    statev[1] = p0 + dp
    dep = dp * nflow
    dep[3:] *= 2.0  # engineering plastic shear strains
    statev[2:8] = ep0 + dep
    # consistent tangent
    tang = np.repeat(H[:, :, None], N, axis=2)
    if idx.size:
        pn = statev[1, idx]
        Rp = dR(pn)
        beta = 1.0 - 3 * mu * dp[idx] / q[idx]  # q_new / q_trial
        Idev = np.diag([1, 1, 1, 0.5, 0.5, 0.5]) - np.outer([1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0]) / 3.0
        nh = (s[:, idx] / q[idx]) * np.sqrt(1.5)  # unit normal (stress Voigt), nh:nh = 1
        nn = np.einsum("in,jn->ijn", nh, nh)
        K_b = E / (3 * (1 - 2 * nu))
        vol = np.outer([1, 1, 1, 0, 0, 0], [1, 1, 1, 0, 0, 0])[:, :, None] * K_b
        gamma = 1.0 / (1.0 + Rp / (3 * mu)) - (1.0 - beta)
        tang[:, :, idx] = vol + 2 * mu * beta * Idev[:, :, None] - 2 * mu * gamma * nn
    return sig, statev, tang


def j2_continuum_tangent(sig, statev, statev_start, props):
    """Continuum elasto-plastic tangent at the END state, (6, 6, N):  L - (L:n)(n:L) / (n:L:n + R'(p)) on the Gauss
    points whose p grew during the increment, L elsewhere.  This is the operator the reference's legacy law builds
    (elasto_plasticity.py:275-300, with the hardening-slope sign corrected) and the one a cutting-plane umat
    returns; for J2, L:n = 3 mu s/q (stress Voigt) and n:L:n = 3 mu."""
    E, nu, _alpha, _sigY, k, m = [float(x) for x in props]
    mu = 0.5 * E / (1 + nu)
    H = elastic_isotropic_H(E, nu)
    N = sig.shape[1]
    tang = np.repeat(H[:, :, None], N, axis=2)
    idx = np.where(statev[1] > statev_start[1])[0]
    if idx.size:
        s = sig[:, idx].copy()
        s[:3] -= s[:3].sum(axis=0) / 3.0
        q = np.sqrt(1.5 * (_VOIGT_W[:, None] * s * s).sum(axis=0))
        sh = s / q
        Rp = k * m * np.power(statev[1, idx], m - 1.0)
        tang[:, :, idx] -= (3 * mu) ** 2 / (3 * mu + Rp) * np.einsum("in,jn->ijn", sh, sh)
    return tang
