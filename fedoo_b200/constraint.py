"""Periodic boundary conditions with mean-strain global dofs.

Mirrors the slice of ``fd.constraint.PeriodicBC`` (fedoo/constraint/periodic_bc.py:8-111 constructor,
:2634-2808 initialize, :910-1800 the MPC list of a periodic mesh) that ``fd.homogen`` uses: a box-shaped cell whose
opposite faces carry matching nodes, ``periodicity_type="small_strain"``:

    u_i(x+) - u_i(x-) = sum_j sc_ij (x+ - x-)_j E_ij,   sc = 1 on the diagonal and 0.5 off it (E_xy = 2 eps_xy)

with E = the "MeanStrain" global dofs [E_xx, E_yy, E_zz, E_xy, E_xz, E_yz] (2-D: [E_xx, E_yy, E_xy]) appended after the
nodal dofs.  Every node of a max-face is tied to its image on the min-faces (faces -> opposite face, edges and corners
-> the all-min edge / corner), which spans the same constraint space as the reference's face / edge / corner lists.

The reference turns the MPCs into a sparse change-of-basis matrix and forms MatCB^T A MatCB on the host
(fedoo/core/problem.py:277-298).  Here they become a ``MpcMap`` on the device and the reduced operator is applied
matrix-free inside the CG loop (csrc/fdk_solve.cuh: expand / fold kernels around the tiled SpMV).
"""

from __future__ import annotations

import ctypes as C

import numpy as np
import torch

from . import _lib

_E_INDEX_3D = [[0, 3, 4], [3, 1, 5], [4, 5, 2]]
_E_INDEX_2D = [[0, 2], [2, 1]]


class MpcMap:
    """x[slave] = x[master] + coef @ x[n_nodal:]: device arrays + the ctypes mirror of ``struct fdk_mpc``."""

    def __init__(self, n_nodal, n_glob, slave, master, coef):
        slave = np.asarray(slave, dtype=np.int64)
        master = np.asarray(master, dtype=np.int64)
        coef = np.ascontiguousarray(coef, dtype=float).reshape(len(slave), n_glob)
        if np.intersect1d(slave, master).size:
            raise ValueError("a master dof is itself eliminated: chain constraints are not supported")
        if np.unique(slave).size != slave.size:
            raise ValueError("a dof is eliminated twice")
        self.n_nodal, self.n_glob = int(n_nodal), int(n_glob)
        self.slave_h, self.master_h, self.coef_h = slave, master, coef
        self._struct = None

    def _to_device(self):
        """Device copies + the ctypes struct, built on first use (the host arrays alone serve the CPU tests)."""
        from .core import device

        dev = device()
        order = np.argsort(self.master_h, kind="stable")
        mst_dof, counts = np.unique(self.master_h, return_counts=True)
        mst_ptr = np.concatenate([[0], np.cumsum(counts)])
        i32 = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev)  # noqa: E731
        self.slave, self.master = i32(self.slave_h), i32(self.master_h)
        self.coef = torch.from_numpy(self.coef_h).to(dev)
        self.mst_dof, self.mst_ptr, self.mst_slv = i32(mst_dof), i32(mst_ptr), i32(order)
        self.scratch = torch.zeros(_lib.MPC_SCRATCH_DOUBLES, dtype=torch.float64, device=dev)
        self._struct = _lib.MpcStruct(
            self.n_nodal, self.n_glob, len(self.slave_h), _lib.ptr(self.slave), _lib.ptr(self.master), _lib.ptr(self.coef),
            len(mst_dof), _lib.ptr(self.mst_dof), _lib.ptr(self.mst_ptr), _lib.ptr(self.mst_slv), _lib.ptr(self.scratch),
        )  # fmt: skip

    @property
    def n_total(self):
        return self.n_nodal + self.n_glob

    def struct(self):
        if self._struct is None:
            self._to_device()
        return C.byref(self._struct)

    def expand(self, x):
        """x <- T x in place (slave entries from their masters and the global dofs)."""
        assert x.is_cuda and x.dtype == torch.float64 and x.numel() == self.n_total
        _lib.check(_lib.load().fdk_mpc_expand(self.struct(), _lib.ptr(x), _lib.current_stream()), "fdk_mpc_expand")
        return x

    def fold(self, q):
        """q <- T^T q in place (slave rows added to their masters and, weighted, to the global rows; slaves cleared).
        The global rows are OVERWRITTEN by the folded sum: add any load on them afterwards."""
        assert q.is_cuda and q.dtype == torch.float64 and q.numel() == self.n_total
        _lib.check(_lib.load().fdk_mpc_fold(self.struct(), _lib.ptr(q), _lib.current_stream()), "fdk_mpc_fold")
        return q

    def to_scipy(self):
        """The change-of-basis matrix T (n_total x n_total, slave columns empty) for host-side cross-checks."""
        from scipy import sparse

        n, g = self.n_total, self.n_glob
        keep = np.ones(n, dtype=bool)
        keep[self.slave_h] = False
        rows = [np.nonzero(keep)[0], self.slave_h]
        cols = [np.nonzero(keep)[0], self.master_h]
        vals = [np.ones(int(keep.sum())), np.ones(len(self.slave_h))]
        for k in range(g):
            rows.append(self.slave_h)
            cols.append(np.full(len(self.slave_h), self.n_nodal + k))
            vals.append(self.coef_h[:, k])
        return sparse.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


class PeriodicBC:
    """``fd.constraint.PeriodicBC(periodicity_type="small_strain", dim=None, meshperio=True, tol=1e-8)``."""

    def __init__(self, periodicity_type="small_strain", off_axis_rotation=None, dim=None, meshperio=True, tol=1e-8,
                 name="Periodicity"):  # fmt: skip
        if not isinstance(periodicity_type, str):
            raise TypeError("periodicity_type should be a string")
        if periodicity_type == "small_strain":
            self.shear_coef = 0.5
        elif periodicity_type == "finite_strain":
            raise NotImplementedError("finite_strain periodicity (9 displacement-gradient dofs) is not on the path")
        else:
            raise ValueError("periodicity_type should be either 'small_strain' or 'finite_strain'")
        if off_axis_rotation is not None:
            raise NotImplementedError("off_axis_rotation is not on the path")
        if not meshperio:
            raise NotImplementedError("non-periodic meshes (closest-point constraints) are not on the path")
        self.periodicity_type, self.dim, self.meshperio, self.tol = periodicity_type, dim, meshperio, tol
        self.bc_type, self.name = "PeriodicBC", name
        self.mpc = None

    def node_pairs(self, mesh):
        """(slave nodes, master nodes, offset x_slave - x_master as multiples of the cell size)."""
        from scipy.spatial import cKDTree

        X = mesh.nodes[:, : self.dim]
        lo, hi = X.min(axis=0), X.max(axis=0)
        self.d_rve = hi - lo
        on_hi = np.abs(X - hi) < self.tol
        on_lo = np.abs(X - lo) < self.tol
        is_slave = on_hi.any(axis=1)
        slaves = np.nonzero(is_slave)[0]
        cand = np.nonzero(on_lo.any(axis=1) & ~is_slave)[0]
        if len(slaves) == 0 or len(cand) == 0:
            raise ValueError("no boundary nodes found: is the mesh a box-shaped cell?")
        delta = on_hi[slaves] * self.d_rve
        dist, idx = cKDTree(X[cand]).query(X[slaves] - delta)
        if dist.max() > max(self.tol, 1e-6 * self.d_rve.max()):
            raise ValueError("the mesh is not periodic: a node of a max-face has no image on the opposite face")
        return slaves, cand[idx], delta

    def initialize(self, problem):
        """Creates the global dofs at ``pb.bc.add(...)`` time like the reference (periodic_bc.py:2671-2700,
        fedoo/core/boundary_conditions.py:116-118) and builds the constraint map."""
        if self.dim is None:
            self.dim = problem.space.ndim
        if self.dim == 3:
            names, emap = ["E_xx", "E_yy", "E_zz", "E_xy", "E_xz", "E_yz"], _E_INDEX_3D
        elif self.dim == 2:
            names, emap = ["E_xx", "E_yy", "E_xy"], _E_INDEX_2D
        else:
            raise NotImplementedError("1-D periodicity")
        problem.add_global_dof(names, 1, "MeanStrain")
        mesh = problem.mesh
        slaves, masters, delta = self.node_pairs(mesh)
        n = mesh.n_nodes
        disp = ["DispX", "DispY", "DispZ"][: self.dim]
        s_dof, m_dof, coef = [], [], []
        for i, var in enumerate(disp):
            r = problem.space.variable_rank(var)
            c = np.zeros((len(slaves), len(names)))
            for j in range(self.dim):
                c[:, emap[i][j]] += (1.0 if i == j else self.shear_coef) * delta[:, j]
            s_dof.append(r * n + slaves)
            m_dof.append(r * n + masters)
            coef.append(c)
        self.mpc = MpcMap(problem.space.nvar * n, len(names), np.concatenate(s_dof), np.concatenate(m_dof), np.concatenate(coef))
