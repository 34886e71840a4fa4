"""Minimal mirrors of ``fd.problem.Linear`` / ``fd.problem.NonLinear``: the CALLERS of the
assembly path (fedoo/problem/linear.py:7-100, fedoo/problem/non_linear.py:13-224,
fedoo/core/problem.py:18-120).  They hold the dof vector, drive the assembly lifecycle and
expose Dirichlet conditions.  The sparse solve is out of scope of the accelerated path
(BASELINE.json north_star); by default ``solve()`` eliminates Dirichlet dofs and calls SciPy on the
host like the reference's direct solver, so that the reference's known-answer tests can be
replayed end to end.  ``set_solver("cg")`` (fedoo/core/base.py:444-519: the reference's iterative
option, scipy cg with a Jacobi preconditioner :521-537) runs the elimination and the Krylov loop on
the DEVICE instead (csrc/fdk_solve.cuh): K never leaves HBM.
"""

from __future__ import annotations

import numpy as np

from .assembly import Assembly
from .core import _Named


class _BC:
    def __init__(self, pb):
        self._pb = pb
        self.list = []

    def add(self, bc_type, node_set, variable, value=0, **kargs):
        if bc_type != "Dirichlet":
            raise NotImplementedError("only Dirichlet conditions are mirrored")
        if isinstance(node_set, str):
            node_set = self._pb.mesh.node_sets[node_set]
        self.list.append((np.asarray(node_set, dtype=np.int64), variable, value))


class _ProblemBase(_Named):
    _dict = {}

    def __init__(self, assembly, name="MainProblem"):
        if isinstance(assembly, str):
            assembly = Assembly.get_all()[assembly]
        self.assembly = assembly
        self.mesh = assembly.mesh
        self.space = assembly.space
        self.n_global_dof = 0
        self.bc = _BC(self)
        self._dirichlet = None
        self.nlgeom = False
        self.time = 0
        self.dtime = 0
        self._solver = ("direct", True, {})
        self.solver_info = None
        self._register(name)

    def set_solver(self, solver="direct", precond=True, **kargs):
        """fedoo/core/base.py:444-519.  "direct" / "direct_scipy": host spsolve; "cg": Jacobi-preconditioned CG on
        the device (kargs: rtol / tol, maxiter)."""
        solver = solver.lower()
        if solver in ("direct", "direct_scipy"):
            solver = "direct"
        elif solver != "cg":
            raise NameError("Choosen solver not available")
        self._solver = (solver, bool(precond), dict(kargs))

    def _solve(self, A, D, X0=None):
        if self._solver[0] == "cg":
            return self._solve_device(A, D, X0)
        return self._solve_host(A, D, X0)

    def _solve_device(self, A, D, X0=None):
        """A dX = D with the imposed dofs eliminated, on the device: rhs = (D - A Xbc) on the free dofs, then
        Jacobi-PCG restricted to them (fedoo/core/problem.py:277-298, fedoo/core/base.py:521-537)."""
        import torch

        from .core import as_device_f64

        if self._dirichlet is None:
            self.apply_boundary_conditions()
        dofs, vals = self._dirichlet
        n = A.shape[0]
        dev = A.data.device
        Xbc = torch.zeros(n, dtype=torch.float64, device=dev)
        free = torch.ones(n, dtype=torch.uint8, device=dev)
        if len(dofs):
            d_dofs = torch.from_numpy(dofs).to(dev)
            imposed = vals if X0 is None else vals - np.asarray(X0)[dofs]
            Xbc[d_dofs] = torch.from_numpy(np.ascontiguousarray(imposed, dtype=float)).to(dev)
            free[d_dofs] = 0
        rhs = torch.zeros(n, dtype=torch.float64, device=dev) if np.isscalar(D) else as_device_f64(D, dev).clone()
        rhs -= A.matvec(Xbc)
        kargs = self._solver[2]
        rtol = kargs.get("rtol", kargs.get("tol", 1e-8))
        x, it, rel = A.pcg(rhs, free_mask=free, rtol=rtol, maxiter=kargs.get("maxiter"))
        self.solver_info = {"iterations": it, "relative_residual": rel}
        if rel > rtol:
            print(f"Warning: cg solver convergence to tolerance not achieved ({rel:.2e} after {it} iterations)")
        return (x + Xbc).cpu().numpy()

    @property
    def n_dof(self):
        return self.assembly.nvar * self.mesh.n_nodes + self.n_global_dof

    def _var_rank(self, name):
        return self.space.variable_rank(name)

    def _slice(self, X, name):
        if name in ("all", None):
            return X
        n = self.mesh.n_nodes
        if name in ("Disp",):
            return X[: self.space.ndim * n].reshape(self.space.ndim, n)
        r = self._var_rank(name)
        return X[r * n : (r + 1) * n]

    def get_results(self, *args, **kargs):
        """fedoo/core/problem.py:207-265: get_results(assemb, output_list, output_type) or
        get_results(output_list, output_type)."""
        from .results import get_results

        args = list(args)
        if args and (isinstance(args[0], Assembly) or (isinstance(args[0], str) and args[0] in Assembly.get_all())):
            assemb = args.pop(0)
            if isinstance(assemb, str):
                assemb = Assembly.get_all()[assemb]
        else:
            assemb = kargs.pop("assemb", self.assembly)
        output_list = args.pop(0) if args else kargs.pop("output_list")
        output_type = args.pop(0) if args else kargs.pop("output_type", None)
        return get_results(self, assemb, output_list, output_type)

    def apply_boundary_conditions(self):
        """Dirichlet part of fedoo/core/problem.py:335-432."""
        n = self.mesh.n_nodes
        dofs, vals = [], []
        for nodes, var, value in self.bc.list:
            r = self._var_rank(var)
            dofs.append(r * n + nodes)
            vals.append(np.broadcast_to(np.asarray(value, dtype=float), nodes.shape))
        if dofs:
            dofs, vals = np.concatenate(dofs), np.concatenate(vals)
            dofs, first = np.unique(dofs, return_index=True)
            self._dirichlet = (dofs, vals[first])
        else:
            self._dirichlet = (np.zeros(0, dtype=np.int64), np.zeros(0))

    def _solve_host(self, A, D, X0=None):
        """Solve A dX = D with Dirichlet elimination on the host (out of the accelerated scope)."""
        from scipy.sparse.linalg import spsolve

        A = A.tocsr()
        n = A.shape[0]
        if self._dirichlet is None:
            self.apply_boundary_conditions()
        dofs, vals = self._dirichlet
        X = np.zeros(n)
        X[dofs] = vals if X0 is None else vals - X0[dofs]
        free = np.setdiff1d(np.arange(n), dofs)
        rhs = (D if not np.isscalar(D) else np.zeros(n))[free] - (A @ X)[free]
        X[free] = spsolve(A[free][:, free].tocsc(), rhs)
        return X


class Linear(_ProblemBase):
    def __init__(self, assembly, name="MainProblem"):
        super().__init__(assembly, name)
        self._X = 0
        self.assembly.initialize(self)

    def set_X(self, X):
        self._X = X

    def get_X(self):
        return self._X

    def get_dof_solution(self, name="all"):
        if np.isscalar(self._X):
            return self._X
        return self._slice(self._X, name)

    def get_disp(self, name="Disp"):
        return self.get_dof_solution(name)

    def update(self, dtime=1, compute="all"):
        self.assembly.update(self, compute)

    def solve(self, **kargs):
        A = self.assembly.get_global_matrix()
        D = self.assembly.get_global_vector()
        X0 = None if np.isscalar(self._X) else np.asarray(self._X)
        dX = self._solve(A, D, X0)
        self._X = dX if X0 is None else X0 + dX
        if kargs.pop("updateWF", True):
            self.update(compute="none")


class NonLinear(_ProblemBase):
    """Holds U (converged) and dU (current increment): dof solution = U + dU
    (fedoo/problem/non_linear.py:13-131)."""

    def __init__(self, assembly, name="MainProblem"):
        super().__init__(assembly, name)
        self._U = 0
        self._dU = 0
        self.nr_parameters = {"err0": None, "criterion": "Displacement", "tol": 1e-3, "max_subiter": 5, "norm_type": 2}

    def get_dof_solution(self, name="all"):
        if np.isscalar(self._U) and np.isscalar(self._dU):
            return self._U + self._dU
        return self._slice(np.asarray(self._U + self._dU), name)

    def get_temp(self):
        return self.get_dof_solution("Temp")

    def get_disp(self, name="Disp"):
        return self.get_dof_solution(name)

    def initialize(self):
        self.assembly.initialize(self)

    def set_start(self):
        self.assembly.set_start(self)

    def to_start(self):
        self._dU = 0
        self.assembly.to_start(self)

    def update(self, compute="all"):
        self.assembly.update(self, compute)

    def nlsolve(self, dt=0.1, tmax=1.0, max_subiter=5, tol=1e-6):
        """Fixed-step Newton-Raphson loop (fedoo/problem/non_linear.py:434-634 without automatic
        time-step control); Dirichlet values are applied at the first iteration of every
        increment.  Returns the number of increments."""
        self.initialize()
        n_inc = int(round(tmax / dt))
        self.dtime = dt
        for inc in range(n_inc):
            self.time = (inc + 1) * dt
            self.set_start()
            A, D = self.assembly.get_global_matrix(), self.assembly.get_global_vector()
            X0 = np.zeros(self.n_dof) if np.isscalar(self._U) else np.asarray(self._U)
            self._dU = self._solve(A, D, X0)  # elastic prediction with the imposed values
            for _ in range(max_subiter):
                self.update(compute="all")
                D = self.assembly.get_global_vector()
                dofs = self._dirichlet[0]
                res = np.array(D, copy=True)
                res[dofs] = 0.0
                if np.linalg.norm(res) <= tol * max(np.linalg.norm(D), 1e-300):
                    break
                corr = self._solve(self.assembly.get_global_matrix(), D, X0 + self._dU)
                self._dU = self._dU + corr
            self._U = X0 + self._dU
            self._dU = 0
        return n_inc
