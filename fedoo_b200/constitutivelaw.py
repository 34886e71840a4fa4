"""Constitutive laws of the assembly path, same constructors and ``assembly.sv`` protocol as
the reference (fedoo/core/base.py:210-260: initialize / update / set_start / to_start / reset
communicate only through ``assembly.sv``).

* ``ElasticIsotrop``      fedoo/constitutivelaw/elastic_isotrop.py:8-77
* ``ElasticAnisotropic``  fedoo/constitutivelaw/elastic_anisotropic.py:10-77
* ``ThermalProperties``   fedoo/constitutivelaw/thermal_prop.py:6-19
* ``ElastoPlasticity`` / ``Simcoon("EPICP")``  fedoo/constitutivelaw/elasto_plasticity.py:230-376,
  fedoo/constitutivelaw/simcoon_umat.py:463-593 (sv protocol: Stress (6,N), Statev (8,N),
  TangentMatrix (6,6,N), elastic reset of the tangent at set_start).
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .core import GaussPointTensor, _Named, device


class ConstitutiveLaw(_Named):
    _dict = {}

    def __init__(self, name=""):
        self._register(name)

    def initialize(self, assembly, pb):
        pass

    def update(self, assembly, pb):
        pass

    def set_start(self, assembly, pb):
        pass

    def to_start(self, assembly, pb):
        pass

    def reset(self):
        pass


def plane_stress_H(H):
    """fedoo/core/mechanical3d.py:47-69."""
    H = np.asarray(H, dtype=float)
    out = np.zeros((6, 6))
    for i in (0, 1, 3):
        for j in (0, 1, 3):
            out[i, j] = H[i, j] - H[i, 2] * H[j, 2] / H[2, 2]
    return out


class ElasticAnisotropic(ConstitutiveLaw):
    """sigma = H eps with H 6x6 (uniform) or (6,6,N) per Gauss point (Fortran order)."""

    def __init__(self, H, name=""):
        ConstitutiveLaw.__init__(self, name)
        self._H = H
        self._H_dev = None

    def get_tangent_matrix(self, assembly=None, dimension=None):
        if dimension is None:
            dimension = assembly.space.get_dimension()
        H = self._H
        if dimension == "2Dstress":
            if np.ndim(H) != 2:
                raise NotImplementedError("plane-stress reduction of a per-Gauss-point tangent")
            return plane_stress_H(H)
        return H

    def get_elastic_matrix(self, dimension="3D"):
        return self.get_tangent_matrix(None, dimension)

    def initialize(self, assembly, pb):
        assembly.sv["TangentMatrix"] = self.get_tangent_matrix(assembly)

    def update(self, assembly, pb):
        # linear law: the tangent is kept; the stress follows the strain of the current dof vector
        if "TangentMatrix" not in assembly.sv:
            assembly.sv["TangentMatrix"] = self.get_tangent_matrix(assembly)
        assembly._elastic_stress_update(self)

    def get_stress_from_strain(self, assembly, strain_tensor, dimension=None):
        """sigma_i = sum_j H_ij eps_j on whatever support the strain lives (nodes, elements, Gauss points):
        fedoo/constitutivelaw/elastic_anisotropic.py:58-77."""
        from .results import NodeTensor

        H = self.get_tangent_matrix(assembly, dimension)
        if np.ndim(H) != 2:
            raise NotImplementedError("per-Gauss-point tangent: use sv['Stress']")
        eps = [np.asarray(strain_tensor[j]) for j in range(6)]
        return NodeTensor([sum(eps[j] * H[i][j] for j in range(6)) for i in range(6)])

    def tangent_device(self, assembly):
        """Per-GP tangent as a CUDA tensor in (6,6,N) Fortran layout, or None if uniform."""
        H = assembly.sv["TangentMatrix"]
        if isinstance(H, torch.Tensor):
            return H
        if np.ndim(H) == 3:
            if self._H_dev is None or self._H_dev[0] is not H:
                flat = np.asfortranarray(H, dtype=float).ravel(order="K")
                self._H_dev = (H, torch.from_numpy(flat).to(device()))
            return self._H_dev[1]
        return None


class ElasticIsotrop(ElasticAnisotropic):
    def __init__(self, E, nu, name=""):
        ConstitutiveLaw.__init__(self, name)
        self.E, self.nu = E, nu
        self._H_dev = None

    @property
    def G(self):
        return self.E / (1 + self.nu) / 2

    def get_tangent_matrix(self, assembly=None, dimension=None):
        if dimension is None:
            dimension = assembly.space.get_dimension()
        E, nu = self.E, self.nu
        if not (np.isscalar(E) and np.isscalar(nu)):
            raise NotImplementedError("per-Gauss-point E / nu: pass a (6,6,N) tangent to ElasticAnisotropic")
        H = np.zeros((6, 6))
        if dimension == "2Dstress":
            H[0, 0] = H[1, 1] = E / (1 - nu**2)
            H[0, 1] = H[1, 0] = nu * E / (1 - nu**2)
            H[3, 3] = 0.5 * E / (1 + nu)
        else:
            H[0, 0] = H[1, 1] = H[2, 2] = E * (1.0 / (1 + nu) + nu / ((1.0 + nu) * (1 - 2 * nu)))
            H[0, 1] = H[0, 2] = H[1, 2] = E * (nu / ((1 + nu) * (1 - 2 * nu)))
            H[3, 3] = H[4, 4] = H[5, 5] = 0.5 * E / (1 + nu)
            H[1, 0], H[2, 0], H[2, 1] = H[0, 1], H[0, 2], H[1, 2]
        return H

    def lame(self, dimension):
        """(lambda, mu) of the closed-form kernel; plane stress uses lambda* = nu E / (1 - nu^2)."""
        E, nu = float(self.E), float(self.nu)
        mu = 0.5 * E / (1 + nu)
        lam = nu * E / (1 - nu**2) if dimension == "2Dstress" else E * (nu / ((1 + nu) * (1 - 2 * nu)))
        return lam, mu


class ThermalProperties(ConstitutiveLaw):
    def __init__(self, thermal_conductivity, specific_heat, density, name=""):
        ConstitutiveLaw.__init__(self, name)
        if np.isscalar(thermal_conductivity):
            k = thermal_conductivity
            self.thermal_conductivity = [[k, 0, 0], [0, k, 0], [0, 0, k]]
        else:
            self.thermal_conductivity = thermal_conductivity
        self.specific_heat = specific_heat
        self.density = density


class ElastoPlasticity(ConstitutiveLaw):
    """J2 plasticity with isotropic power-law hardening sigma_Y + k p^m, small strain.

    Constructor follows fedoo/constitutivelaw/elasto_plasticity.py:30-80 (E, nu, yield stress)
    with ``set_hardening_function('power', H=k, beta=m)`` (:127-133); state protocol follows the
    living Simcoon path (simcoon_umat.py:463-593).  The update runs ``fdk_j2_update`` on the
    device: backward-Euler radial return + consistent tangent per Gauss point."""

    def __init__(self, E, nu, yield_stress, name=""):
        ConstitutiveLaw.__init__(self, name)
        self.E, self.nu, self.yield_stress = E, nu, yield_stress
        self.k, self.m = 0.0, 1.0
        # "consistent": algorithmic tangent of the radial return (quadratic Newton convergence);
        # "continuum": L - (L:n)(n:L)/(n:L:n + R') at the end state, what simcoon's EPICP returns
        self.tangent = "consistent"
        # keep the per-Gauss-point tangent in its 10-double structured form (see StructuredTangent); 0: the (6,6,N) array
        self.structured_tangent = True

    def set_hardening_function(self, function_type="power", **kargs):
        if function_type.lower() != "power":
            raise NotImplementedError("only the power-law hardening R = H p^beta is available")
        self.k, self.m = float(kargs["H"]), float(kargs["beta"])

    @property
    def props(self):
        return np.array([self.E, self.nu, 0.0, self.yield_stress, self.k, self.m], dtype=float)

    def get_elastic_matrix(self, dimension="3D"):
        return ElasticIsotrop(self.E, self.nu).get_tangent_matrix(None, dimension)

    def initialize(self, assembly, pb):
        if assembly.space.ndim != 3:
            raise NotImplementedError("ElastoPlasticity is available for the 3D modeling space")
        N = assembly.n_gauss_points
        dev = device()
        assembly.sv["Statev"] = torch.zeros((N, 8), dtype=torch.float64, device=dev)
        assembly.sv["Stress"] = GaussPointTensor(torch.zeros((N, 6), dtype=torch.float64, device=dev))
        assembly.sv["TangentMatrix"] = self.get_elastic_matrix()
        assembly.sv_component.update({"T": ("Statev", 0), "P": ("Statev", 1), "EP": ("Statev", slice(2, 8))})

    def update(self, assembly, pb):
        strain = assembly.sv["Strain"]
        if not isinstance(strain, GaussPointTensor):
            return  # no displacement yet
        lib = _lib.load()
        N = assembly.n_gauss_points
        dev = device()
        sv0 = assembly.sv_start.get("Statev")
        if sv0 is None:
            sv0 = torch.zeros((N, 8), dtype=torch.float64, device=dev)
        stress = torch.empty((N, 6), dtype=torch.float64, device=dev)
        statev = torch.empty((N, 8), dtype=torch.float64, device=dev)
        props = self.props
        _lib.check(lib.fdk_set_option(b"j2_continuum_tangent", int(self.tangent == "continuum")), "fdk_set_option")
        lazy_U = getattr(strain, "U", None) if (getattr(strain, "_dev", 0) is None and not getattr(strain, "fbar", False)) else None
        if self.structured_tangent and lazy_U is not None:
            # strain not materialised yet: geometry, grad u, strain and the radial return in one pass (fdk_j2_update_from_dofs);
            # sv["Strain"] stays lazy
            r1 = torch.empty((N, 10), dtype=torch.float64, device=dev)
            coords, conn = assembly._coords(), assembly.mesh.device_arrays()[1]
            _lib.check(
                lib.fdk_j2_update_from_dofs(
                    _lib.ELEM_IDS[assembly.elm_type], assembly.mesh.n_nodes, assembly.mesh.n_elements, _lib.ptr(conn),
                    _lib.ptr(coords), _lib.ptr(lazy_U), _lib.ptr(props), _lib.ptr(sv0), _lib.ptr(stress), _lib.ptr(statev),
                    None, _lib.ptr(r1), _lib.current_stream(),
                ),
                "fdk_j2_update_from_dofs",
            )  # fmt: skip
            assembly.sv["TangentMatrix"] = StructuredTangent(r1)
        elif self.structured_tangent:
            # 10 doubles per Gauss point instead of 36 (csrc/fdk_gp.cuh); the (6,6,N) array is built only if somebody reads it
            r1 = torch.empty((N, 10), dtype=torch.float64, device=dev)
            _lib.check(
                lib.fdk_j2_update_r1(
                    N, _lib.ptr(props), _lib.ptr(strain.device_tensor), _lib.ptr(sv0), _lib.ptr(stress), _lib.ptr(statev),
                    _lib.ptr(r1), _lib.current_stream(),
                ),
                "fdk_j2_update_r1",
            )  # fmt: skip
            assembly.sv["TangentMatrix"] = StructuredTangent(r1)
        else:
            tangent = torch.empty(N * 36, dtype=torch.float64, device=dev)
            _lib.check(
                lib.fdk_j2_update(
                    N, _lib.ptr(props), _lib.ptr(strain.device_tensor), _lib.ptr(sv0), _lib.ptr(stress), _lib.ptr(statev),
                    _lib.ptr(tangent), _lib.current_stream(),
                ),
                "fdk_j2_update",
            )  # fmt: skip
            assembly.sv["TangentMatrix"] = tangent
        assembly.sv["Stress"] = GaussPointTensor(stress)
        assembly.sv["Statev"] = statev

    def set_start(self, assembly, pb):
        # elastic prediction for the next increment (simcoon_umat.py:591-593)
        assembly.sv["TangentMatrix"] = self.get_elastic_matrix()

    def tangent_device(self, assembly):
        """(6,6,N) Fortran-ordered device tensor of the per-Gauss-point tangent, or None when it is uniform."""
        H = assembly.sv["TangentMatrix"]
        if isinstance(H, StructuredTangent):
            return H.full()
        return H if isinstance(H, torch.Tensor) else None

    def tangent_r1_device(self, assembly):
        """The structured form [lam', mu', kappa, n^(6), 0] per Gauss point, or None."""
        H = assembly.sv["TangentMatrix"]
        return H.r1 if isinstance(H, StructuredTangent) else None


class StructuredTangent:
    """sv["TangentMatrix"] of the J2 law in its structured form (N, 10): C = lam' 1(x)1 + 2 mu' I_sym - kappa n^(x)n^.
    Indexing ``H[i][j]`` (what the reference's weak form does, stress_equilibrium.py:112-117), ``np.asarray`` and
    ``full()`` expand it to the (6,6,N) array of the reference's protocol (simcoon_umat.py:556-580)."""

    def __init__(self, r1):
        self.r1 = r1
        self._full = None
        self._host = None

    def full(self):
        if self._full is None:
            N = self.r1.shape[0]
            out = torch.empty(N * 36, dtype=torch.float64, device=self.r1.device)
            _lib.check(_lib.load().fdk_j2_tangent_expand(N, _lib.ptr(self.r1), _lib.ptr(out), _lib.current_stream()),
                       "fdk_j2_tangent_expand")  # fmt: skip
            self._full = out
        return self._full

    def __array__(self, dtype=None, copy=None):
        if self._host is None:
            N = self.r1.shape[0]
            self._host = self.full().cpu().numpy().reshape(N, 6, 6).transpose(2, 1, 0)  # (i, j, n), Fortran (6,6,N)
        return self._host

    def __getitem__(self, i):
        return np.asarray(self)[i]

    @property
    def shape(self):
        return (6, 6, self.r1.shape[0])


def Simcoon(umat_name, props, name=""):
    """``fd.constitutivelaw.Simcoon("EPICP", [E, nu, alpha, sigmaY, k, m])``
    (fedoo/constitutivelaw/simcoon_umat.py:103-113); only the J2 'EPICP' law is on this path."""
    if umat_name.upper() != "EPICP":
        raise NotImplementedError(f"umat '{umat_name}' is not on the accelerated path (only EPICP)")
    E, nu, _alpha, sigY, k, m = [float(x) for x in props]
    law = ElastoPlasticity(E, nu, sigY, name=name)
    law.set_hardening_function("power", H=k, beta=m)
    law.tangent = "continuum"  # simcoon's cutting-plane umat returns the continuum tangent at the end state
    return law
