"""The reference's one J2 known-answer test (tests/test_octet.py) driven by the REFERENCE'S OWN problem code with the
oracle's J2 law standing in for simcoon.  TEST INFRASTRUCTURE ONLY.

simcoon (C++) is absent, so ``fd.constitutivelaw.Simcoon("EPICP", ...)`` cannot run.  Here a stand-in module
``simcoon.simmit`` is registered in ``sys.modules`` before fedoo is imported; its ``umat`` is the NumPy restatement
``oracle.fedoo_oracle.j2_radial_return`` (pinned on the reference's elasto_plasticity.py loop by
``oracle/gen_golden_j2.py``) with the continuum tangent L - (L:n)(n:L)/(n:L:n + R') that a cutting-plane umat returns.
Everything else -- mesh import, PeriodicBC, NonLinear / Newton-Raphson with the "Work" criterion, assembly, direct
solver, get_results -- is the unmodified reference.  The result pins the whole DRIVER side of the replay:

    Stress[4][222] = 72.27748615821348   (reference + real simcoon: 72.3765265291865, test tolerance 1e-3)
    Strain[2][876] = 0.030477251173930353 (reference + real simcoon: 0.03046909551762696, tolerance 1e-6)

i.e. the 1.4e-3 / 2.7e-4 distance to the reference's numbers is a property of the LAW (simcoon's own cutting-plane
iteration: each increment of this test does exactly one Newton correction, so the answers depend on the tangent of the
points that have just yielded), not of the assembly, the constraint map or the Newton loop.  The CUDA replay
(tests/test_gpu_parity.py::test_octet_replay...) must reproduce THESE numbers.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_octet_driver.py   -> tests/golden/octet_driver_j2.npz
"""

from __future__ import annotations

import importlib.util
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.environ.get("FEDOO_REFERENCE", "/root/reference")
sys.path.insert(0, ROOT)
from oracle import fedoo_oracle as fo  # noqa: E402


def umat(name, etot, Detot, F0, F1, sigma, DR, props, statev, t, dt, Wm, temp, ndi=3):
    """Signature of simcoon.simmit.umat as fedoo calls it (constitutivelaw/simcoon_umat.py:509-526,561-578)."""
    assert name == "EPICP" and ndi == 3
    props = np.asarray(props, dtype=float).ravel()
    eps = np.asarray(etot) + np.asarray(Detot)
    sv0 = np.asarray(statev, dtype=float)
    sig, sv, _ = fo.j2_radial_return(eps, sv0, props)
    tang = fo.j2_continuum_tangent(sig, sv, sv0, props)
    return sig, sv, np.asarray(Wm), np.asfortranarray(tang)


def main():
    simcoon = types.ModuleType("simcoon")
    simmit = types.ModuleType("simcoon.simmit")
    simmit.umat = umat
    simcoon.simmit = simmit
    simcoon.__version__ = "stand-in"
    sys.modules["simcoon"] = simcoon
    sys.modules["simcoon.simmit"] = simmit
    sys.path.insert(0, REF)
    import fedoo as fd

    spec = importlib.util.spec_from_file_location("ref_test_octet", os.path.join(REF, "tests", "test_octet.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    try:
        mod.test_octet()
        passed = True
    except AssertionError:
        passed = False
    res = fd.Problem.get_all()["MainProblem"].get_results("Assembly", ["Strain", "Stress", "Disp"], "GaussPoint")
    stress = np.asarray(res.gausspoint_data["Stress"])
    strain = np.asarray(res.gausspoint_data["Strain"])
    print("reference asserts pass:", passed)
    print("Stress[4][222] =", repr(float(stress[4][222])), " Strain[2][876] =", repr(float(strain[2][876])))
    out = os.path.join(ROOT, "tests", "golden", "octet_driver_j2.npz")
    np.savez_compressed(out, stress=stress, strain=strain, known=np.array([stress[4][222], strain[2][876]]))
    print("wrote", out, os.path.getsize(out), "bytes")


if __name__ == "__main__":
    main()
