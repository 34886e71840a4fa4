#!/usr/bin/env python
"""Golden vectors for results extraction, produced by the REFERENCE itself (run in the build
container only: PYTHONPATH=/root/reference python oracle/gen_golden_results.py).  Test infrastructure, never
imported by the product path.

  results_plate.npz      replay of tests/test_platewithhol.py: mesh, boundary sets, solution, and
                         pb.get_results("Assembly", ["Stress_vm", "Strain", "Stress"], <Node|Element|GaussPoint>)
                         with the test's two known answers asserted here
  results_cantilever.npz tests/test_cantilever_beam_3D_model.py: node strain / stress of the solved state
  results_hex8.npz, results_tet10.npz   random displacement field: node / element conversion of Strain, Stress, Stress_vm
"""
import os
import sys

import numpy as np

sys.path.insert(0, "/root/reference")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fedoo as fd  # noqa: E402

from fedoo_b200 import meshgen  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def extract(pb, name, fields=("Stress_vm", "Strain", "Stress")):
    out = {}
    for typ, attr in (("Node", "node_data"), ("Element", "element_data"), ("GaussPoint", "gausspoint_data")):
        res = pb.get_results(name, list(fields), typ)
        for f in fields:
            out[f"{f}_{typ}"] = np.asarray(getattr(res, attr)[f])
    return out


def plate():
    fd.Assembly.delete_memory()
    fd.ModelingSpace("2Dstress")
    mesh = fd.mesh.hole_plate_mesh(nr=11, nt=11, length=100, height=100, radius=20, elm_type="quad4", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(2e5, 0.3, name="ElasticLaw")
    fd.weakform.StressEquilibrium("ElasticLaw", name="WeakForm")
    fd.Assembly.create("WeakForm", "Domain", name="Assembly", MeshChange=True)
    pb = fd.problem.Linear("Assembly")
    left = mesh.find_nodes("X", mesh.bounding_box.xmin)
    right = mesh.find_nodes("X", mesh.bounding_box.xmax)
    bottom = mesh.find_nodes("Y", mesh.bounding_box.ymin)
    pb.bc.add("Dirichlet", left, "DispX", 0)
    pb.bc.add("Dirichlet", bottom, "DispY", 0)
    pb.bc.add("Dirichlet", right, "DispX", 0.1)
    pb.apply_boundary_conditions()
    pb.solve()
    res = pb.get_results("Assembly", ["Stress_vm", "Strain"], "Node")
    U = pb.get_disp()
    assert U[0, 40] == 0.1
    assert abs(U[1, 40] + 0.01962855744173) < 1e-10
    assert abs(res.node_data["Stress_vm"][282] - 175.50126302014) < 1e-10
    out = dict(nodes=mesh.nodes, elements=mesh.elements.astype(np.int32), left=left, right=right, bottom=bottom,
               U_sol=pb.get_dof_solution("all"))  # fmt: skip
    out.update(extract(pb, "Assembly"))
    np.savez_compressed(os.path.join(OUT, "results_plate.npz"), **out)
    print("results_plate", {k: v.shape for k, v in out.items()})


def cantilever():
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    mesh = fd.mesh.box_mesh(11, 5, 5, 0, 1000, 0, 100, 0, 100, "hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="ElasticLaw")
    fd.weakform.StressEquilibrium("ElasticLaw", name="weakform")
    fd.Assembly.create("weakform", "Domain", "hex8", name="Assembling")
    pb = fd.problem.Linear("Assembling")
    for v in ("DispX", "DispY", "DispZ"):
        pb.bc.add("Dirichlet", mesh.node_sets["left"], v, 0)
    pb.bc.add("Dirichlet", mesh.node_sets["right"], "DispY", -10)
    pb.apply_boundary_conditions()
    pb.solve()
    asm = fd.Assembly["Assembling"]
    strain = asm.get_strain(pb.get_dof_solution(), "Node", nlgeom=False)
    stress = fd.ConstitutiveLaw["ElasticLaw"].get_stress_from_strain(asm, strain)
    assert abs(stress[5][-1] + 0.9007983467254552) < 1e-10
    out = dict(U_sol=pb.get_dof_solution("all"), strain_node_legacy=np.asarray(strain.asarray()),
               stress_node_legacy=np.asarray(stress.asarray()))  # fmt: skip
    out.update(extract(pb, "Assembling"))
    np.savez_compressed(os.path.join(OUT, "results_cantilever.npz"), **out)
    print("results_cantilever", {k: v.shape for k, v in out.items()})


def random_field(tag, nodes, elements, elm_type, seed):
    fd.Assembly.delete_memory()
    fd.ModelingSpace("3D")
    fd.Mesh(nodes, elements, elm_type, name="Domain")
    fd.constitutivelaw.ElasticIsotrop(1e5, 0.3, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    a = fd.Assembly.create("wf", "Domain", elm_type, name="A")
    pb = fd.problem.Linear("A")
    U = np.random.default_rng(seed).standard_normal(pb.n_dof) * 1e-3
    pb.set_X(U)
    a.update(pb, compute="none")
    out = dict(nodes=np.asarray(nodes, dtype=float), elements=np.asarray(elements, dtype=np.int32), U=U)
    out.update(extract(pb, "A"))
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, {k: v.shape for k, v in out.items()})


if __name__ == "__main__":
    plate()
    cantilever()
    nodes, elements = meshgen.box_hex8(7, 6, 5)
    random_field("results_hex8", meshgen.jitter_nodes(nodes, 7, 6, 5), elements, "hex8", 11)
    nodes, hexes = meshgen.box_hex8(5, 4, 4)
    nodes = meshgen.jitter_nodes(nodes, 5, 4, 4)
    n10, e10 = meshgen.tet4_to_tet10(nodes, meshgen.hex8_to_tet4(hexes), bulge=0.03)
    random_field("results_tet10", n10, e10, "tet10", 12)
