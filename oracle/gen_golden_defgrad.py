#!/usr/bin/env python
"""Golden vectors for the finite-strain kinematics the reference computes in its own Python:
fedoo/weakform/stress_equilibrium.py:542-586, `_comp_F` (F = 1 + grad u) and `_comp_Fbar` (F (J_mean / J)^(1/3)),
produced by the REFERENCE's own functions in the build container:

    PYTHONPATH=/root/reference python oracle/gen_golden_defgrad.py

The functions are module-level helpers taking (assembly, displacement); the reference only calls them under nlgeom,
which needs simcoon for what FOLLOWS them (strain measures, objective rates) -- they themselves run on a plain
small-strain assembly.  Meshes: tests/golden/hex8_jitter.npz and tet10_box.npz, a dof vector large enough for det F to
vary by tens of percent.  Test infrastructure only.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import fedoo as fd  # noqa: E402
from fedoo.weakform import stress_equilibrium as se  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    out = {}
    for name, elm in (("hex8_jitter", "hex8"), ("tet10_box", "tet10")):
        g = np.load(os.path.join(OUT, name + ".npz"))
        fd.Assembly.delete_memory()
        fd.ModelingSpace("3D")
        mesh = fd.Mesh(g["nodes"], g["elements"].astype(np.int64), elm, name="Domain")
        fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="law")
        fd.weakform.StressEquilibrium("law", name="wf")
        a = fd.Assembly.create("wf", "Domain", elm, name="A")
        fd.problem.Linear("A")
        h = (g["nodes"].max(axis=0) - g["nodes"].min(axis=0)).min() / 8
        U = np.random.default_rng(21).standard_normal(3 * mesh.n_nodes) * 0.04 * h
        se._comp_F(a, U)
        F = np.array(a.sv["F"])
        a.sv_start.pop("F", None)
        se._comp_Fbar(a, U)
        Fbar = np.array(a.sv["F"])
        J = np.linalg.det(F.transpose(2, 0, 1))
        Jb = np.linalg.det(Fbar.transpose(2, 0, 1)).reshape(a.n_elm_gp, -1)
        assert F.shape == (3, 3, a.n_gauss_points) and J.min() > 0.2 and J.max() - J.min() > 0.1
        assert np.abs(Jb - Jb.mean(axis=0)).max() < 1e-12  # what F-bar does: det is constant per element
        out[name + "_U"], out[name + "_F"], out[name + "_Fbar"] = U, F, Fbar
        print(name, F.shape, "det F in", J.min(), J.max())
    np.savez_compressed(os.path.join(OUT, "defgrad.npz"), **out)


if __name__ == "__main__":
    main()
