"""Golden for a transient heat solve from the DEFAULT zero start (the path tests/test_thermal3D.py exercises:
NonLinear + HeatEquation, first update with a scalar-0 temperature, heat_equation.py:140-147).  TEST INFRASTRUCTURE ONLY.

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_heat_nlsolve.py   (needs /root/reference)
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.environ.get("FEDOO_REFERENCE", "/root/reference"))
import fedoo as fd  # noqa: E402

g = np.load(os.path.join(ROOT, "tests", "golden", "tet4_box.npz"))
fd.Assembly.delete_memory()
fd.ModelingSpace("3D")
mesh = fd.Mesh(np.array(g["nodes"]), np.array(g["elements"]), "tet4", name="Domain")
fd.constitutivelaw.ThermalProperties(500, 0.5, 7800, name="ThermalLaw")
fd.weakform.HeatEquation("ThermalLaw")
fd.Assembly.create("ThermalLaw", "Domain", name="Assembling")
pb = fd.problem.NonLinear("Assembling")
pb.set_nr_criterion("Displacement", tol=5e-2, max_subiter=5, err0=100)  # tests/test_thermal3D.py:45
right = mesh.find_nodes("X", mesh.bounding_box.xmax)
pb.bc.add("Dirichlet", right, "Temp", 3)
pb.nlsolve(dt=10 / 3, tmax=10, update_dt=True)  # tests/test_thermal3D.py:75
T = pb.get_dof_solution()
print("T range", T.min(), T.max(), "n_nodes", mesh.n_nodes)
np.savez_compressed(os.path.join(ROOT, "tests", "golden", "heat_nlsolve_tet4.npz"), T=T, right=right, tmax=10.0, n_inc=3)
