"""Replay of the reference's tests/test_octet.py (tet4 octet-truss cell, EPICP J2 law, PeriodicBC, mean shear strain
E_xy = 0.1 in five increments, Work criterion with tol 0.1) on the CUDA path; prints the two values the reference
test pins (tests/test_octet.py:80-81)."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
import fedoo_b200 as fd  # noqa: E402

g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "octet_truss_tet4.npz"))
solver = sys.argv[1] if len(sys.argv) > 1 else "cg"
fd.Assembly.delete_memory()
fd.ModelingSpace("3D")
mesh = fd.Mesh(g["nodes"], g["elements"], "tet4", name="Domain2")
props = np.array([1e5, 0.3, 1e-5, 300, 1000, 0.25])
fd.constitutivelaw.Simcoon("EPICP", props, name="ConstitutiveLaw")
fd.weakform.StressEquilibrium("ConstitutiveLaw", name="WeakForm", nlgeom=False)
fd.Assembly.create("WeakForm", "Domain2", "tet4", name="Assembly")
pb = fd.problem.NonLinear("Assembly")
pb.set_nr_criterion(criterion="Work")
pb.set_solver(solver, rtol=float(os.environ.get("RTOL", "1e-12")))
pb.bc.add(fd.constraint.PeriodicBC("small_strain", dim=3))
pb.bc.add("Dirichlet", int(g["center"]), "Disp", 0)
pb.bc.add("Dirichlet", 0, "MeanStrain", [0, 0, 0, 0.1, 0, 0])
tol_nr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.1
pb.nlsolve(dt=0.2, tmax=1, update_dt=False, tol_nr=tol_nr, print_info=int(os.environ.get("PRINT_INFO", "2")))
res = pb.get_results("Assembly", ["Strain", "Stress"], "GaussPoint")
S, E = res.gausspoint_data["Stress"], res.gausspoint_data["Strain"]
print("Stress[4][222] =", S[4][222], " (reference 72.3765265291865 +- 1e-3)")
print("Strain[2][876] =", E[2][876], " (reference 0.03046909551762696 +- 1e-6)")
