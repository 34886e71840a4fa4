"""``fd.homogen``: homogenised stiffness of a periodic cell.

Mirrors fedoo/homogen/tangent_stiffness.py:18-191: a perturbation problem on the already assembled stiffness with
PeriodicBC (mean-strain global dofs), the centre node pinned, and six unit "Neumann" loads on E_xx .. E_yz; the
mean-strain responses give C = inv(DStrain^T) / volume.  The matrix is assembled once by the CUDA assembly path and
stays in HBM; each load case is one constrained Jacobi-PCG on the device (``solver="cg"``, the default here) or the
reference's host elimination + direct solve (``solver="direct"``, for cross-checks).
"""

from __future__ import annotations

import numpy as np

from .assembly import Assembly
from .constraint import PeriodicBC
from .problem import Linear, Problem, _ProblemBase

_STRAIN_3D = ["E_xx", "E_yy", "E_zz", "E_xy", "E_xz", "E_yz"]
_STRAIN_2D = ["E_xx", "E_yy", "E_xy"]


def get_homogenized_stiffness(assemb, meshperio=True, **kargs):
    """fedoo/homogen/tangent_stiffness.py:18-29."""
    if isinstance(assemb, str):
        assemb = Assembly.get_all()[assemb]
    pb = Linear(assemb, name="_homogen_main")
    pb.set_A(assemb.get_global_matrix())
    return get_tangent_stiffness(pb, meshperio, **kargs)


def get_tangent_stiffness(pb=None, meshperio=True, **kargs):
    """fedoo/homogen/tangent_stiffness.py:32-191 (perturbation method, Neumann loads on the mean-strain dofs).
    kargs: solver ("cg" on the device | "direct" on the host), rtol, maxiter, lockstep (default True: the load cases
    are solved together, K read once per iteration for all of them; False: one after the other)."""
    solver = kargs.pop("solver", "cg")
    if pb is None:
        pb = _ProblemBase.get_active()
    elif isinstance(pb, str):
        pb = _ProblemBase.get_all()[pb]
    mesh = pb.mesh
    center = [int(np.linalg.norm(mesh.nodes - mesh.bounding_box.center, axis=1).argmin())]
    ndim = pb.space.ndim
    names = _STRAIN_3D if ndim == 3 else _STRAIN_2D

    registry = _ProblemBase.get_all()
    if "_perturbation" in registry and registry["_perturbation"].mesh is not mesh:
        del registry["_perturbation"]
    if "_perturbation" not in registry:
        pert = Problem(0, 0, 0, mesh, name="_perturbation", space=pb.space)
        pb.make_active()
        pert.bc.add(PeriodicBC("small_strain", meshperio=meshperio))
        pert.bc.add("Dirichlet", center, list(pb.space.list_variables()), 0, name="center")
    else:
        pert = registry["_perturbation"]
    kargs.setdefault("rtol", 1e-10)
    lockstep_opt = kargs.pop("lockstep", True)
    pert.set_solver(solver, **{k: v for k, v in kargs.items() if v is not None and k not in ("solver_type", "pc_type")})
    kargs["lockstep"] = lockstep_opt
    pert.set_A(pb.get_A())

    d_strain, info = [], []
    lockstep = kargs.pop("lockstep", True)
    if solver == "cg" and lockstep:
        # the load cases share K: one lockstep solve reads it once per iteration for all of them
        pert.bc.remove("_Strain")
        pert.apply_boundary_conditions()
        loads = np.zeros((pert.n_global_dof, len(names)))
        for k, name in enumerate(names):
            loads[pert._global_dof[name], k] = 1.0
        X = pert.solve_load_cases(pert.get_A(), loads)  # (n_dof, R) on the device
        pert._X = X[:, -1].cpu().numpy()
        E = X[X.shape[0] - pert.n_global_dof :].cpu().numpy()  # mean-strain response of every case
        d_strain = [np.array([E[pert._global_dof[name], i] for name in names]) for i in range(len(names))]
        info = [pert.solver_info] * len(names)
    else:
        for i in range(len(names)):
            pert.bc.remove("_Strain")
            for k, name in enumerate(names):
                pert.bc.add("Neumann", name, 1.0 if k == i else 0.0, start_value=0, name="_Strain")
            pert.apply_boundary_conditions()
            pert.solve()
            X = pert.get_X()
            d_strain.append(np.array([pert._get_vect_component(X, name)[0] for name in names]))
            info.append(pert.solver_info)
    pert.bc.remove("_Strain")
    pert.load_case_info = info
    return np.linalg.inv(np.array(d_strain).T) / mesh.bounding_box.volume
