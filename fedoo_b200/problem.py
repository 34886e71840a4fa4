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
    """Boundary-condition list (fedoo/core/boundary_conditions.py:90-260): ``add("Dirichlet" | "Neumann", node_set,
    variable(s), value, name=...)``, ``add("Neumann", "E_xx", value)`` on a global dof, ``add(PeriodicBC(...))``,
    ``remove(name)``."""

    def __init__(self, pb):
        self._pb = pb
        self.list = []  # (nodes, variable, value, type, name)
        self.constraints = []

    def add(self, *args, **kargs):
        pb = self._pb
        if not isinstance(args[0], str):  # a constraint object: initialised when added (boundary_conditions.py:116-118)
            bc = args[0]
            bc.initialize(pb)
            self.constraints.append(bc)
            pb._dirichlet = None
            return bc
        bc_type = args[0]
        if bc_type not in ("Dirichlet", "Neumann"):
            raise NotImplementedError("only Dirichlet / Neumann conditions and PeriodicBC are mirrored")
        rest = list(args[1:])
        if rest and isinstance(rest[0], str) and rest[0] in pb._global_dof:  # global dof: no node set
            node_set, variable = np.zeros(1, dtype=np.int64), rest.pop(0)
        else:
            node_set = rest.pop(0) if rest else kargs.pop("node_set")
            variable = rest.pop(0) if rest else kargs.pop("variable")
            if isinstance(node_set, str):
                node_set = pb.mesh.node_sets[node_set]
            node_set = np.atleast_1d(np.asarray(node_set, dtype=np.int64))
        value = rest.pop(0) if rest else kargs.pop("value", 0)
        name = kargs.get("name", "")
        variables = [variable] if isinstance(variable, str) else list(variable)
        for var in variables:
            comps = pb._expand_vector(var)
            per_comp = len(comps) > 1 and np.ndim(value) >= 1 and len(value) == len(comps)  # e.g. "MeanStrain", [E...]
            for i, v in enumerate(comps):
                self.list.append((node_set, v, value[i] if per_comp else value, bc_type, name))
        pb._dirichlet = None

    def remove(self, name):
        self.list = [bc for bc in self.list if bc[4] != name]
        self._pb._dirichlet = None


class _ProblemBase(_Named):
    _dict = {}
    _active = None

    def __init__(self, assembly, name="MainProblem", mesh=None, space=None):
        if isinstance(assembly, str):
            assembly = Assembly.get_all()[assembly]
        self.assembly = assembly
        self.mesh = assembly.mesh if assembly is not None else mesh
        if isinstance(self.mesh, str):
            from .core import Mesh

            self.mesh = Mesh[self.mesh]
        if space is None:
            from .core import ModelingSpace

            space = assembly.space if assembly is not None else ModelingSpace.get_active()
        self.space = space
        self.n_global_dof = 0
        self._global_dof = {}  # variable name -> index among the global dofs
        self._global_vector = {}  # vector name -> list of variable names
        self.bc = _BC(self)
        self._dirichlet = None
        self._neumann = 0
        self.nlgeom = False
        self.time = 0
        self.dtime = 0
        self._solver = ("direct", True, {})
        self.solver_info = None
        self._register(name)
        _ProblemBase._active = self

    def make_active(self):
        _ProblemBase._active = self

    @staticmethod
    def get_active():
        return _ProblemBase._active

    def add_global_dof(self, variable_names, n_dof=1, vector_name=None):
        """fedoo/core/base.py:336-366: dofs that belong to no node, stored after the nodal ones."""
        if isinstance(variable_names, str):
            variable_names = [variable_names]
        if n_dof != 1:
            raise NotImplementedError("one dof per global variable")
        for v in variable_names:
            if v not in self._global_dof:
                self._global_dof[v] = self.n_global_dof
                self.n_global_dof += 1
        if vector_name is not None:
            self._global_vector[vector_name] = list(variable_names)
        self._dirichlet = None
        return np.array([self._global_dof[v] for v in variable_names])

    def _expand_vector(self, var):
        if var in self._global_vector:
            return self._global_vector[var]
        if var == "Disp" and "Disp" not in self.space._variable:
            return ["DispX", "DispY", "DispZ"][: self.space.ndim]
        return [var]

    def _dof_of(self, var, nodes):
        if var in self._global_dof:
            return np.full(len(nodes), self.space.nvar * self.mesh.n_nodes + self._global_dof[var], dtype=np.int64)
        return self._var_rank(var) * self.mesh.n_nodes + nodes

    def _get_vect_component(self, X, name):
        """fedoo/core/problem.py:100-130."""
        n = self.mesh.n_nodes
        if name in self._global_dof:
            i = self.space.nvar * n + self._global_dof[name]
            return X[i : i + 1]
        if name in self._global_vector:
            i = self.space.nvar * n + self._global_dof[self._global_vector[name][0]]
            return X[i : i + len(self._global_vector[name])]
        return self._slice(X, name)

    def get_ext_forces(self, name="all"):
        """External (reaction) forces A X - D (fedoo/core/problem.py:470-497 without the MPC redistribution): the
        product runs on the device matrix, only the vector comes back."""
        import torch

        A = self.get_A()
        X = self.get_X()
        if np.isscalar(X):
            raise ValueError("no solution yet")
        n = A.shape[0]
        Xd = torch.from_numpy(np.ascontiguousarray(np.asarray(X)[:n], dtype=float)).to(A.data.device)
        F = np.zeros(self.n_dof)
        F[:n] = A.matvec(Xd).cpu().numpy()
        D = self.get_D() if hasattr(self, "get_D") else 0
        if not (np.isscalar(D) and D == 0):
            F[: len(D)] -= np.asarray(D)
        return self._get_vect_component(F, name)

    @property
    def _mpc(self):
        maps = [c.mpc for c in self.bc.constraints if getattr(c, "mpc", None) is not None]
        if len(maps) > 1:
            raise NotImplementedError("one constraint map per problem")
        return maps[0] if maps else None

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
        if self._dirichlet is None:
            self.apply_boundary_conditions()
        if not np.isscalar(self._neumann):  # external forces B (fedoo/core/problem.py:288-296: A X = B + D)
            D = self._neumann + (0 if np.isscalar(D) else self._pad(np.asarray(D)))
        if self._solver[0] == "cg":
            return self._solve_device(A, D, X0)
        return self._solve_host(A, D, X0)

    def _pad(self, v):
        return v if v.shape[0] == self.n_dof else np.concatenate([v, np.zeros(self.n_dof - v.shape[0])])

    def _solve_device(self, A, D, X0=None):
        """A dX = D with the imposed dofs eliminated, on the device: rhs = (D - A Xbc) on the free dofs, then
        Jacobi-PCG restricted to them (fedoo/core/problem.py:277-298, fedoo/core/base.py:521-537).  With a
        PeriodicBC the constraint map T is applied matrix-free: T^T A T y = T^T (D - A Xbc), X = T y + Xbc."""
        import torch

        from .core import as_device_f64

        dofs, vals = self._dirichlet
        mpc = self._mpc
        n = self.n_dof if mpc is not None else A.shape[0]
        n_mat = A.shape[0]  # nodal size, or already resized by the global dofs (trailing empty rows)
        dev = A.data.device
        Xbc = torch.zeros(n, dtype=torch.float64, device=dev)
        free = torch.ones(n, dtype=torch.uint8, device=dev)
        if len(dofs):
            d_dofs = torch.from_numpy(dofs).to(dev)
            imposed = vals if X0 is None else vals - np.asarray(X0)[dofs]
            Xbc[d_dofs] = torch.from_numpy(np.ascontiguousarray(imposed, dtype=float)).to(dev)
            free[d_dofs] = 0
        rhs = torch.zeros(n, dtype=torch.float64, device=dev)
        if not np.isscalar(D):
            d = as_device_f64(D, dev)
            rhs[: d.numel()] += d
        kargs = self._solver[2]
        rtol = kargs.get("rtol", kargs.get("tol", 1e-8))
        if mpc is None:
            rhs -= A.matvec(Xbc)
            x, it, rel = A.pcg(rhs, free_mask=free, rtol=rtol, maxiter=kargs.get("maxiter"))
        else:
            if np.intersect1d(dofs, mpc.slave_h).size:
                raise NotImplementedError("a Dirichlet condition on an eliminated (slave) dof")
            mpc.expand(Xbc)
            rhs[:n_mat] -= A.matvec(Xbc[:n_mat] if n_mat < n else Xbc)[:n_mat]
            load = rhs[mpc.n_nodal :].clone()  # loads on the global dofs (fold overwrites those rows)
            mpc.fold(rhs)
            rhs[mpc.n_nodal :] += load
            free[torch.from_numpy(mpc.slave_h).to(dev)] = 0
            x, it, rel = A.pcg(rhs, free_mask=free, rtol=rtol, maxiter=kargs.get("maxiter"), mpc=mpc)
            mpc.expand(x)
        self.solver_info = {"iterations": it, "relative_residual": rel}
        if rel > rtol:
            print(f"Warning: cg solver convergence to tolerance not achieved ({rel:.2e} after {it} iterations)")
        return (x + Xbc).cpu().numpy()

    def solve_load_cases(self, A, global_loads):
        """Several load cases on one matrix at once, on the device (fd.homogen): ``global_loads`` is (n_global_dof, R),
        the loads on the global dofs of each case (R = 3 or 6); the nodal loads and all imposed values are zero.
        Returns the (n_dof, R) solutions as a CUDA tensor."""
        import torch

        if self._dirichlet is None:
            self.apply_boundary_conditions()
        dofs, vals = self._dirichlet
        if np.any(vals != 0):
            raise NotImplementedError("non-zero imposed values with several load cases")
        mpc = self._mpc
        dev = A.data.device
        global_loads = np.ascontiguousarray(global_loads, dtype=float)
        n, R = self.n_dof, global_loads.shape[1]
        assert global_loads.shape[0] == self.n_global_dof and A.n_nodal == n - self.n_global_dof
        rhs = torch.zeros((n, R), dtype=torch.float64, device=dev)  # T^T of a load on the global rows is itself
        if self.n_global_dof:
            rhs[n - self.n_global_dof :] = torch.from_numpy(global_loads).to(dev)
        free = torch.ones(n, dtype=torch.uint8, device=dev)
        if len(dofs):
            free[torch.from_numpy(dofs).to(dev)] = 0
        if mpc is not None:
            if np.intersect1d(dofs, mpc.slave_h).size:
                raise NotImplementedError("a Dirichlet condition on an eliminated (slave) dof")
            free[torch.from_numpy(mpc.slave_h).to(dev)] = 0
        kargs = self._solver[2]
        rtol = kargs.get("rtol", kargs.get("tol", 1e-8))
        X, it, rel = A.pcg_multi(rhs, free_mask=free, rtol=rtol, maxiter=kargs.get("maxiter"), mpc=mpc)
        self.solver_info = {"iterations": it, "relative_residual": max(rel), "relative_residuals": rel}
        if max(rel) > rtol:
            print(f"Warning: cg solver convergence to tolerance not achieved ({max(rel):.2e} after {it} iterations)")
        return X

    @property
    def n_dof(self):
        return self.space.nvar * self.mesh.n_nodes + self.n_global_dof

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

    def apply_boundary_conditions(self, t_fact=1.0):
        """Dirichlet and Neumann part of fedoo/core/problem.py:335-432; incremental problems pass the time factor:
        values ramp linearly from 0 (fedoo/core/boundary_conditions.py: start_value + t_fact (value - start_value))."""
        dofs, vals = [], []
        B = None
        for nodes, var, value, bc_type, _name in self.bc.list:
            d = self._dof_of(var, nodes)
            v = np.broadcast_to(np.asarray(value, dtype=float) * t_fact, nodes.shape)
            if bc_type == "Dirichlet":
                dofs.append(d)
                vals.append(v)
            else:
                if B is None:
                    B = np.zeros(self.n_dof)
                np.add.at(B, d, v)
        self._neumann = 0 if B is None else B
        if dofs:
            dofs, vals = np.concatenate(dofs), np.concatenate(vals)
            dofs, first = np.unique(dofs, return_index=True)
            self._dirichlet = (dofs, vals[first])
        else:
            self._dirichlet = (np.zeros(0, dtype=np.int64), np.zeros(0))

    def _solve_host(self, A, D, X0=None):
        """Solve A dX = D with Dirichlet (and periodic MPC) elimination on the host like the reference
        (fedoo/core/problem.py:277-298; out of the accelerated scope, kept for cross-checks)."""
        from scipy import sparse
        from scipy.sparse.linalg import spsolve

        A = A.tocsr()
        mpc = self._mpc
        n = self.n_dof if mpc is not None else A.shape[0]
        if A.shape[0] < n:
            A = sparse.csr_matrix((A.data, A.indices, np.concatenate([A.indptr, np.full(n - A.shape[0], A.indptr[-1])])), shape=(n, n))
        dofs, vals = self._dirichlet
        X = np.zeros(n)
        X[dofs] = vals if X0 is None else vals - X0[dofs]
        keep = np.ones(n, dtype=bool)
        keep[dofs] = False
        rhs = np.zeros(n) if np.isscalar(D) else self._pad(np.asarray(D, dtype=float)).copy() if mpc is not None else np.asarray(D, dtype=float)
        if mpc is None:
            free = np.nonzero(keep)[0]
            rhs = rhs[free] - (A @ X)[free]
            X[free] = spsolve(A[free][:, free].tocsc(), rhs)
            return X
        T = mpc.to_scipy()
        X = T @ X
        keep[mpc.slave_h] = False
        free = np.nonzero(keep)[0]
        Tf = T[:, free]
        y = spsolve((Tf.T @ A @ Tf).tocsc(), Tf.T @ (rhs - A @ X))
        return Tf @ y + X


class Problem(_ProblemBase):
    """``fd.Problem(A, B, D, mesh, name)``: the generic linear problem A X = B + D (fedoo/core/problem.py:18-98) that
    ``fd.homogen`` builds around an already assembled stiffness matrix."""

    def __init__(self, A=0, B=0, D=0, mesh=None, name="MainProblem", space=None):
        super().__init__(None, name, mesh=mesh, space=space)
        self._A, self._B, self._D = A, B, D
        self._X = 0

    def set_A(self, A):
        self._A = A

    def get_A(self):
        return self._A

    def set_B(self, B):
        self._B = B

    def set_D(self, D):
        self._D = D

    def get_D(self):
        return self._D

    def get_X(self):
        return self._X

    def get_dof_solution(self, name="all"):
        return self._X if np.isscalar(self._X) else self._get_vect_component(self._X, name)

    def solve(self, **kargs):
        rhs = self._B + self._D
        self._X = self._solve(self._A, rhs)


class Linear(_ProblemBase):
    def __init__(self, assembly, name="MainProblem"):
        super().__init__(assembly, name)
        self._X = 0
        self.assembly.initialize(self)

    def set_X(self, X):
        self._X = X

    def set_A(self, A):
        """fedoo/problem/linear.py: the matrix is the assembly's; kept for API parity with fd.homogen."""
        self._A = A

    def get_A(self):
        A = getattr(self, "_A", None)
        return A if A is not None else self.assembly.get_global_matrix()

    def get_D(self):
        return self.assembly.get_global_vector()

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
    """Holds U (converged) and dU (current increment): dof solution = U + dU.  The increment loop follows the
    reference step by step (fedoo/problem/non_linear.py:134-165 elastic prediction, :259-338 Newton-Raphson error,
    :385-432 solve_time_increment, :434-634 nlsolve) so that a script written for it iterates identically: Dirichlet
    values ramp with the time factor, the prediction uses the matrix of set_start (elastic tangent), every
    sub-iteration updates the state (compute="vector"), tests the error, reassembles the tangent and corrects."""

    def __init__(self, assembly, name="MainProblem"):
        super().__init__(assembly, name)
        self._U = 0
        self._dU = 0
        self._X = 0
        self._err0 = None
        self.t0, self.tmax = 0.0, 1.0
        self.err_num = 1e-8
        self.print_info = 0
        self.nr_parameters = {"err0": None, "criterion": "Displacement", "tol": 5e-3, "max_subiter": 10,
                              "dt_increase_niter": None, "norm_type": 2}  # fmt: skip

    def set_nr_criterion(self, criterion="Displacement", **kargs):
        """fedoo/problem/non_linear.py:339-383."""
        if criterion not in ["Displacement", "Force", "Work"]:
            raise NameError('criterion must be set to "Displacement", "Force" or "Work"')
        self.nr_parameters["criterion"] = criterion
        for key in kargs:
            if key not in ["err0", "tol", "max_subiter", "dt_increase_niter", "norm_type"]:
                raise NameError("Newton Raphson parameters should be in ['err0', 'tol', 'max_subiter', "
                                "'dt_increase_niter', 'norm_type']")  # fmt: skip
            self.nr_parameters[key] = kargs[key]

    def get_dof_solution(self, name="all"):
        if np.isscalar(self._U) and np.isscalar(self._dU):
            return self._U + self._dU
        if name == "all" and np.isscalar(self._dU) and self._dU == 0:
            return self._U  # as it is (a pinned / device tensor keeps its placement)
        return self._get_vect_component(np.asarray(self._U + self._dU), name)

    def get_X(self):
        return self._X

    def get_A(self):
        return self.assembly.current.get_global_matrix()

    def get_D(self):
        return self.assembly.current.get_global_vector()

    def get_temp(self):
        return self.get_dof_solution("Temp")

    def get_disp(self, name="Disp"):
        return self.get_dof_solution(name)

    def initialize(self):
        self.assembly.initialize(self)

    def set_start(self):
        if not (np.isscalar(self._dU) and self._dU == 0):
            self._U = self._U + self._dU
            self._dU = 0
        self._err0 = self.nr_parameters["err0"]
        self.assembly.set_start(self)

    def to_start(self):
        self._dU = 0
        self._err0 = self.nr_parameters["err0"]
        self.assembly.to_start(self)

    def update(self, compute="all", updateWeakForm=True):
        if updateWeakForm:
            self.assembly.update(self, compute)
        else:
            self.assembly.current.assemble_global_mat(compute)

    @property
    def t_fact(self):
        return (self.time + self.dtime - self.t0) / (self.tmax - self.t0)

    def _total(self):
        U = self._U + self._dU
        return np.zeros(self.n_dof) if np.isscalar(U) else np.asarray(U)

    def _free_dofs(self):
        keep = np.ones(self.n_dof, dtype=bool)
        keep[self._dirichlet[0]] = False
        if self._mpc is not None:
            keep[self._mpc.slave_h] = False
        return np.nonzero(keep)[0]

    def elastic_prediction(self):
        self.apply_boundary_conditions(self.t_fact)
        self._X = self._solve(self.get_A(), self.get_D(), self._total())  # imposes the increments of the Dirichlet values
        self._dU = self._dU + self._X

    def NewtonRaphsonIncrement(self):
        self._X = self._solve(self.get_A(), self.get_D(), self._total())  # imposed dofs already at their values
        self._dU = self._dU + self._X

    def NewtonRaphsonError(self):
        """fedoo/problem/non_linear.py:259-338 (B = 0 beyond what _solve adds: the Neumann vector)."""
        norm_type = self.nr_parameters["norm_type"]
        free = self._free_dofs()
        if len(free) == 0:
            return 0
        crit = self.nr_parameters["criterion"]
        X = np.asarray(self._X)
        D = self.get_D()
        R = (0 if np.isscalar(D) else self._pad(np.asarray(D))) + self._neumann  # B + D
        R = np.zeros(self.n_dof) if np.isscalar(R) else R
        if self._err0 is None:
            if crit == "Displacement":
                err0 = np.linalg.norm(np.asarray(self._dU), norm_type)
                if err0 == 0:
                    return 1
                return np.linalg.norm(X[free], norm_type) / err0
            if crit == "Force":
                err0 = np.linalg.norm(self.get_ext_forces(), norm_type)
                if err0 == 0:
                    return 1
                return np.linalg.norm(R[free], norm_type) / err0
            self._err0 = 1
            self._err0 = self.NewtonRaphsonError()
            return 1
        if crit == "Displacement":
            return np.linalg.norm(X[free], norm_type) / self._err0
        if crit == "Force":
            return np.linalg.norm(R[free], norm_type) / self._err0
        return np.linalg.norm(X[free] * R[free], norm_type) / self._err0

    def get_ext_forces(self, name="all"):
        A, X = self.get_A(), self._total()
        import torch

        n = A.shape[0]
        F = np.zeros(self.n_dof)
        F[:n] = A.matvec(torch.from_numpy(np.ascontiguousarray(X[:n])).to(A.data.device)).cpu().numpy()
        D = self.get_D()
        if not np.isscalar(D):
            F[: len(D)] -= np.asarray(D)
        return self._get_vect_component(F, name)

    def solve_time_increment(self, max_subiter=None, tol_nr=None):
        """fedoo/problem/non_linear.py:385-432."""
        max_subiter = self.nr_parameters["max_subiter"] if max_subiter is None else max_subiter
        tol_nr = self.nr_parameters["tol"] if tol_nr is None else tol_nr
        self.elastic_prediction()
        subiter, err = 0, 1.0
        for subiter in range(max_subiter):
            self.update(compute="vector")
            err = self.NewtonRaphsonError()
            if self.print_info > 1:
                print("     Subiter {} - Time: {:.5f} - Err: {:.5f}".format(subiter, self.time + self.dtime, err))
            if err < tol_nr:
                return 1, subiter, err
            self.update(compute="matrix", updateWeakForm=False)
            self.NewtonRaphsonIncrement()
        return 0, subiter, err

    def nlsolve(self, dt=0.1, update_dt=True, tmax=None, t0=None, dt_min=1e-6, max_subiter=None, dt_increase_niter=None,
                tol_nr=None, print_info=None, **_ignored):  # fmt: skip
        """fedoo/problem/non_linear.py:434-634 without outputs / callbacks.  Returns the number of increments."""
        if tmax is not None:
            self.tmax = tmax
        if t0 is not None:
            self.t0 = t0
        max_subiter = self.nr_parameters["max_subiter"] if max_subiter is None else max_subiter
        if dt_increase_niter is None:
            dt_increase_niter = self.nr_parameters["dt_increase_niter"] or max_subiter // 3
        tol_nr = self.nr_parameters["tol"] if tol_nr is None else tol_nr
        if print_info is not None:
            self.print_info = print_info
        self.time = self.t0
        if np.isscalar(self._U) and self._U == 0:
            self.initialize()
        restart, n_inc = False, 0
        while self.time < self.tmax - self.err_num:
            self.dtime = min(dt, self.tmax - self.time)
            if restart:
                self.to_start()
                restart = False
            else:
                self.set_start()
            convergence, n_iter, err = self.solve_time_increment(max_subiter, tol_nr)
            if convergence:
                self.time = self.time + self.dtime
                n_inc += 1
                if self.print_info > 0:
                    print("Iter {} - Time: {:.5f} - dt {:.5f} - NR iter: {} - Err: {:.5f}".format(n_inc, self.time, dt, n_iter, err))
                if update_dt and n_iter < dt_increase_niter and dt == self.dtime:
                    dt *= 1.25
            elif update_dt:
                dt *= 0.25
                if dt < dt_min:
                    raise NameError("Current time step is inferior to the specified minimal time step (dt_min)")
                restart = True
            else:
                raise NameError("Newton Raphson iteration has not converged (err: {:.5f})- Reduce the time step or use "
                                "update_dt = True".format(err))  # fmt: skip
        self.set_start()
        return n_inc
