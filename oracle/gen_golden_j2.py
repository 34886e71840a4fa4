"""Reference-derived pin for the J2 state update.  TEST INFRASTRUCTURE ONLY.

The living J2 path of the reference computes in simcoon (C++, not under /root/reference, not installed).  The only J2
arithmetic that exists in the reference tree is ``fedoo/constitutivelaw/elasto_plasticity.py`` -- dead at this commit
(never reachable through ``Assembly.update``), but its yield function (:154-155), flow direction (:157-164), trial
state and per-Gauss-point cutting-plane loop (``ComputeStress``, :303-376) are plain Python that still runs once three
breakages are shimmed AT RUN TIME, IN THIS GENERATOR ONLY (nothing of the reference is edited or copied):

  * ``StressTensorList.vonMises`` / ``.toStrain`` were renamed ``von_mises`` / ``to_strain`` (util/voigt_tensors.py:270,288)
    -> aliases set on the class below;
  * the loop's Newton slope is ``B - dR/dp`` = 3 mu - R' (:356), the wrong sign (it reverses for
    p < (beta H / 3 mu)^(1/(1-beta))) -> the hardening function is given through the reference's own
    ``SetHardeningFunction("user", ...)`` hook with the derivative NEGATED, so that the unchanged expression
    ``B - dphi_dp`` evaluates 3 mu + R'; the yield function itself (which uses R, not R') is untouched;
  * the local tolerance (default 1e-6 on f, in stress units) is tightened with the class's own
    ``SetNewtonRaphsonTolerance``.

What is stored: strain states (6, N), start state (p, eps_p), and the reference loop's sigma, p, eps_p for the Gauss
points where that loop converged (it can step to a negative p from a barely-yielded virgin point and then returns NaN:
those points are flagged ``valid = False`` and excluded).  The tangent is NOT stored: the legacy formula carries the same
sign error and simcoon's tangent is a different definition; tests check the tangent by finite differences of sigma(eps).

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_j2.py     (needs /root/reference; writes tests/golden/j2_reference.npz)
"""

from __future__ import annotations

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.environ.get("FEDOO_REFERENCE", "/root/reference"))

import fedoo as fd  # noqa: E402
from fedoo.constitutivelaw.elasto_plasticity import ElastoPlasticity  # noqa: E402
from fedoo.util.voigt_tensors import StrainTensorList, StressTensorList  # noqa: E402

# run-time shims of the renamed methods (the reference files are not touched)
StressTensorList.vonMises = StressTensorList.von_mises
StressTensorList.toStrain = StressTensorList.to_strain


def make_law(E, nu, sigY, k, m, tol):
    law = ElastoPlasticity(E, nu, sigY)

    def R(p):
        return k * p**m

    def minus_dR(p):  # NEGATED on purpose, see the module docstring
        with np.errstate(divide="ignore", invalid="ignore"):
            return -np.nan_to_num(m * k * p ** (m - 1), posinf=1)

    law.SetHardeningFunction("user", HardeningFunction=R, HardeningFunctionDerivative=minus_dR)
    law.SetNewtonRaphsonTolerance(tol)
    law.reset()
    return law


def run_case(props, eps_list, tol=1e-9):
    """Drive ComputeStress over a sequence of total strains (one call = one converged increment)."""
    E, nu, _alpha, sigY, k, m = props
    law = make_law(E, nu, sigY, k, m, tol)
    out = []
    N = eps_list[0].shape[1]
    p0 = np.zeros(N)
    ep0 = np.zeros((6, N))
    ok_so_far = np.ones(N, bool)
    for eps in eps_list:
        with np.errstate(all="ignore"):
            sig = law.ComputeStress(StrainTensorList(eps.copy()))
        sig = np.asarray(sig.asarray(), dtype=float)
        p = np.asarray(law.GetPlasticity(), dtype=float).copy()
        ep = np.asarray(law.get_strain().asarray(), dtype=float).copy()
        f = StressTensorList(sig).von_mises() - sigY - k * np.maximum(p, 0) ** m
        yielded = p > p0
        valid = ok_so_far & np.isfinite(sig).all(axis=0) & np.isfinite(p) & np.isfinite(ep).all(axis=0) & (p >= p0)
        valid &= np.where(yielded, np.abs(f) <= tol, f <= tol)
        out.append(dict(eps=eps, p0=p0.copy(), ep0=ep0.copy(), sig=sig, p=p, ep=ep, valid=valid))
        # next increment starts from this state (NewTimeIncrement :198-203 commits p and eps_p)
        ok_so_far = valid
        sig_clean = np.where(valid, sig, 0.0)  # noqa: F841
        # invalid points: reset their history so the law object stays finite
        p_c = np.where(valid, p, 0.0)
        ep_c = np.where(valid, ep, 0.0)
        law._ElastoPlasticity__currentP = p_c.copy()
        law._ElastoPlasticity__currentPlasticStrainTensor = StrainTensorList(ep_c.copy())
        law.NewTimeIncrement()
        p0, ep0 = p_c, ep_c
    return out


def main():
    rng = np.random.default_rng(7)
    cases = {}
    for tag, props in (("plate", (200e3, 0.3, 1e-5, 300.0, 1000.0, 0.3)), ("octet", (1e5, 0.3, 1e-5, 300.0, 1000.0, 0.25))):
        N = 1500
        E = props[0]
        # strain directions with random amplitude: roughly 60 % of the points beyond first yield (eps_y = sigY / E)
        d = rng.standard_normal((6, N))
        amp = rng.uniform(0.2, 6.0, N) * (props[3] / E)
        eps1 = d / np.linalg.norm(d, axis=0) * amp
        eps2 = eps1 + (rng.standard_normal((6, N)) * 0.8 * (props[3] / E))  # non-proportional second increment
        eps3 = 0.3 * eps2  # unloading (elastic from the committed state, or reverse yielding)
        res = run_case(props, [eps1, eps2, eps3])
        for i, r in enumerate(res):
            for key, v in r.items():
                cases[f"{tag}_{i}_{key}"] = v
        cases[f"{tag}_props"] = np.array(props)
        for i, r in enumerate(res):
            y = (r["p"] > r["p0"]) & r["valid"]
            print(f"{tag} increment {i}: valid {r['valid'].sum()} / {N}, yielding {y.sum()}")
    out = os.path.join(ROOT, "tests", "golden", "j2_reference.npz")
    np.savez_compressed(out, **cases)
    print("wrote", out, os.path.getsize(out), "bytes; fedoo", fd.__version__)


if __name__ == "__main__":
    main()
