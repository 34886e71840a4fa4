#!/usr/bin/env python
"""Mesh fixture for the replay of the reference's tests/test_octet.py (the only reference test that pins J2 results:
tet4 octet-truss cell, Simcoon EPICP, PeriodicBC, NonLinear.nlsolve; known answers at tests/test_octet.py:80-81).
The reference cannot run that test here (simcoon is absent); this script only converts its mesh with the reference's
own importer so that the GPU box, which has no /root/reference, can replay the script:

    PYTHONPATH=/root/reference python oracle/gen_golden_octet.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
import fedoo as fd  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

fd.ModelingSpace("3D")
fd.mesh.import_file("/root/reference/tests/octet_truss.msh", name="Domain")
mesh = fd.Mesh["Domain2"]
center = mesh.nearest_node(mesh.bounding_box.center)
assert mesh.elm_type == "tet4"
np.savez_compressed(os.path.join(OUT, "octet_truss_tet4.npz"), nodes=mesh.nodes, elements=mesh.elements.astype(np.int32),
                    center=int(center))  # fmt: skip
print(mesh.n_nodes, mesh.n_elements, center, mesh.bounding_box.size)
