"""Weak forms of the assembly path (same constructors / lifecycle hooks as the reference,
fedoo/core/weakform.py:66-157).  The reference turns a weak form into a symbolic DiffOp term
list that ``Assembly`` interprets term by term (fedoo/core/diffop.py,
fedoo/core/assembly.py:303-450); here each weak form names the fused CUDA operator that
integrates all of its terms at once.
"""

from __future__ import annotations

import numpy as np

from .constitutivelaw import ConstitutiveLaw
from .core import ModelingSpace, _Named


class WeakFormBase(_Named):
    _dict = {}

    def __init__(self, name="", space=None):
        self.space = space if space is not None else ModelingSpace.get_active()
        self.assembly_options = {}
        self.constitutivelaw = None
        self._register(name)

    def initialize(self, assembly, pb):
        pass

    def update(self, assembly, pb):
        pass

    def update_2(self, assembly, pb):
        pass

    def set_start(self, assembly, pb):
        pass

    def to_start(self, assembly, pb):
        pass

    def reset(self):
        pass


def _resolve_law(law):
    return ConstitutiveLaw.get_all()[law] if isinstance(law, str) else law


class StressEquilibrium(WeakFormBase):
    """int eps(v) : sigma  (fedoo/weakform/stress_equilibrium.py:26-217), small strain only.

    K = int B^T H B with H = sv['TangentMatrix']; D = -int B^T sv['Stress'].
    nlgeom / F-bar / 2Daxi need simcoon in the reference and are outside this path."""

    operator = "elastic"

    def __init__(self, constitutivelaw, name="", nlgeom=False, space=None):
        law = _resolve_law(constitutivelaw)
        if name == "":
            name = law.name
        WeakFormBase.__init__(self, name, space)
        if nlgeom:
            raise NotImplementedError("nlgeom requires simcoon in the reference and is not on the accelerated path")
        self.space.new_variable("DispX")
        self.space.new_variable("DispY")
        if self.space.ndim == 3:
            self.space.new_variable("DispZ")
        self.constitutivelaw = law
        self.nlgeom = False
        self._fbar = False  # small-strain F-bar stabilisation (fedoo/weakform/stress_equilibrium.py:84,372-384)
        self.assembly_options["assume_sym"] = True

    @property
    def fbar(self):
        return self._fbar

    @fbar.setter
    def fbar(self, value):
        if not isinstance(value, bool):
            raise TypeError("bool expeted for fbar")
        self._fbar = value

    def initialize(self, assembly, pb):
        assembly._nlgeom = False
        assembly.sv.setdefault("Stress", 0)
        assembly.sv.setdefault("Strain", 0)
        assembly.sv["DispGradient"] = 0

    def update(self, assembly, pb):
        U = pb.get_dof_solution()
        if np.isscalar(U) and U == 0:
            assembly.sv["DispGradient"] = 0
            assembly.sv["Stress"] = 0
            assembly.sv["Strain"] = 0
        else:
            assembly._strain_update(U)


class SteadyHeatEquation(WeakFormBase):
    """int grad v . k grad T (fedoo/weakform/heat_equation.py:12-119)."""

    operator = "heat"
    transient = False

    def __init__(self, thermal_constitutivelaw, name=None, nlgeom=False, space=None):
        law = _resolve_law(thermal_constitutivelaw)
        WeakFormBase.__init__(self, law.name if name is None else name, space)
        self.space.new_variable("Temp")
        self.constitutivelaw = law

    def initialize(self, assembly, pb):
        assembly._thermal_state_update(pb, initialize=True)

    def update(self, assembly, pb):
        assembly._thermal_state_update(pb)

    def set_start(self, assembly, pb):
        assembly._thermal_set_start(pb)


class HeatEquation(SteadyHeatEquation):
    """Conduction + lumped rho c dT/dt (fedoo/weakform/heat_equation.py:122-227: WeakFormSum of
    SteadyHeatEquation and TemperatureTimeDerivative with mat_lumping=[False, True], which
    ``Assembly.create`` collapses to a single assembly, fedoo/core/assembly.py:1614-1625)."""

    transient = True
