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
    kargs: solver ("cg" on the device | "direct" on the host), rtol, maxiter."""
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
    pert.set_solver(solver, **{k: v for k, v in kargs.items() if v is not None and k not in ("solver_type", "pc_type")})
    pert.set_A(pb.get_A())

    d_strain, info = [], []
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
