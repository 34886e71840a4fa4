"""fedoo_b200 -- B200-native (sm_100a) backend for fedoo's global-operator assembly path.

The public names mirror the slice of the ``fedoo`` API that the path exposes
(``fd.Assembly.create / assemble_global_mat / get_global_matrix``, ``fd.weakform``,
``fd.constitutivelaw``, ``fd.problem.Linear / NonLinear``, ``fd.mesh.box_mesh``), so a
script written for the reference runs against this package with ``import fedoo_b200 as fd`` -- or, with the reference
itself installed, ``fedoo_b200.install(fedoo)`` puts the same kernels under the real ``fedoo.Assembly``
(``fedoo_b200/adapter.py``).
All numeric work happens in hand-written CUDA kernels behind the C ABI of ``include/fdk.h``
(``fedoo_b200/_fdk.so``); there is no CPU fallback.
"""

from . import adapter, constitutivelaw, constraint, homogen, mesh, meshgen, problem, solver, weakform
from ._lib import FdkError
from .adapter import install, uninstall
from .assembly import Assembly
from .constitutivelaw import ConstitutiveLaw
from .core import DeviceCSR, GaussPointTensor, Mesh, ModelingSpace
from .problem import Problem
from .weakform import WeakFormBase

WeakForm = WeakFormBase

__version__ = "0.1.0"

__all__ = [
    "install", "uninstall", "adapter", "solver", "Assembly", "ConstitutiveLaw", "DeviceCSR", "FdkError", "GaussPointTensor", "Mesh", "ModelingSpace", "Problem",
    "WeakForm", "WeakFormBase", "constitutivelaw", "constraint", "homogen", "mesh", "meshgen", "problem", "weakform",
]  # fmt: skip
