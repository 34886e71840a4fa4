#!/usr/bin/env python
"""Golden vectors for the small-strain F-bar option of StressEquilibrium (SURVEY 8f rank 4,
fedoo/weakform/stress_equilibrium.py:84,213-214,527-540), produced by the REFERENCE itself in the build container:

    PYTHONPATH=/root/reference python oracle/gen_golden_fbar.py

Jittered hex8 box (the mesh of tests/golden/hex8_jitter.npz), nearly incompressible law (nu = 0.45), random dof vector:
sv["DispGradient"], sv["Strain"], sv["Stress"] at the Gauss points and the global vector D after update with
wf.fbar = True.  Test infrastructure only.
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import fedoo as fd  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    g = np.load(os.path.join(OUT, "hex8_jitter.npz"))
    fd.ModelingSpace("3D")
    mesh = fd.Mesh(g["nodes"], g["elements"].astype(np.int64), "hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.45, name="law")
    wf = fd.weakform.StressEquilibrium("law", name="wf")
    wf.fbar = True
    a = fd.Assembly.create("wf", "Domain", "hex8", name="A")
    pb = fd.problem.Linear("A")
    U = np.random.default_rng(11).standard_normal(3 * mesh.n_nodes) * 1e-3
    pb.set_X(U)  # linear problem: the dof solution
    a.update(pb, compute="all")
    grad = np.array([[np.asarray(a.sv["DispGradient"][i][j]) for j in range(3)] for i in range(3)])
    strain = np.asarray(a.sv["Strain"].asarray())
    stress = np.array([np.asarray(s) for s in a.sv["Stress"]])
    D = np.asarray(a.get_global_vector())
    # what F-bar changes: the volumetric part is constant per element
    tr = (grad[0, 0] + grad[1, 1] + grad[2, 2]).reshape(8, -1)
    assert np.abs(tr - tr.mean(axis=0)).max() < 1e-15
    np.savez_compressed(os.path.join(OUT, "hex8_jitter_fbar.npz"), E=200e3, nu=0.45, U=U, grad=grad, strain=strain,
                        stress=stress, D=D)  # fmt: skip
    print("fbar golden:", grad.shape, strain.shape, stress.shape, D.shape, "|D|", np.abs(D).max())


if __name__ == "__main__":
    main()
