"""Generate golden input/output vectors by running the REFERENCE (3MAH/fedoo) in-process.

Run in the build container only (the reference does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 PYTHONPATH=/root/reference python oracle/gen_golden.py

Outputs small ``.npz`` fixtures under ``tests/golden/``.  Each fixture holds the
inputs (mesh, material, dof vector) and what the reference's own
``Assembly.update / assemble_global_mat`` produced for them (K.indptr / indices /
data, global_vector, sv["Strain"], sv["Stress"], ...).  ``tests/`` compares both
the oracle restatement and the CUDA path against these arrays.

For the two real meshes shipped with the reference, the mesh arrays are not
copied; the gyroid (tet4) mesh is stored as arrays since the thermal test of the
reference is built on it, and only K.v products / hashes are kept for the values.
"""

import hashlib
import json
import os
import sys
import warnings

import numpy as np

warnings.filterwarnings("ignore")
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

import fedoo as fd  # noqa: E402  (the reference, via PYTHONPATH)

from fedoo_b200 import meshgen  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def fresh(space):
    fd.Assembly.delete_memory()
    return fd.ModelingSpace(space)


def sv_array(x):
    return np.array(x.asarray()) if hasattr(x, "asarray") else np.array(x)


def elastic_case(tag, space, nodes, elements, elm_type, E, nu, seed=0, H_gp=None, extra=None, store_K=True):
    fresh(space)
    mesh = fd.Mesh(nodes, elements, elm_type, name="Domain")
    if H_gp is None:
        fd.constitutivelaw.ElasticIsotrop(E, nu, name="law")
    else:
        fd.constitutivelaw.ElasticAnisotropic(H_gp, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    a = fd.Assembly.create("wf", "Domain", elm_type, name="A")
    pb = fd.problem.Linear("A")
    U = np.random.default_rng(seed).standard_normal(pb.n_dof) * 1e-3
    pb.set_X(U)
    a.update(pb, compute="all")
    K, D = a.global_matrix, a.global_vector
    v = np.random.default_rng(100 + seed).standard_normal(K.shape[0])
    out = dict(
        nodes=np.asarray(nodes, dtype=float),
        elements=np.asarray(elements, dtype=np.int32),
        E=E,
        nu=nu,
        U=U,
        D=D,
        strain=sv_array(a.sv["Strain"])[:, :: (1 if store_K else 97)],
        stress=sv_array(a.sv["Stress"])[:, :: (1 if store_K else 97)],
        K_shape=np.array(K.shape),
        K_nnz=K.nnz,
        K_indptr_sha=sha(K.indptr),
        K_indices_sha=sha(K.indices),
        K_fro=np.linalg.norm(K.data),
        K_absmax=np.abs(K.data).max(),
        v=v,
        Kv=K @ v,
    )
    if store_K:
        out.update(K_indptr=K.indptr, K_indices=K.indices, K_data=K.data)
    if H_gp is not None:
        out["H_gp"] = H_gp
    if extra:
        out.update(extra)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, K.shape, K.nnz, out["K_indptr_sha"], out["K_indices_sha"], K.indptr.dtype)
    return a, pb, mesh


def random_spd_tangent(n, seed):
    rng = np.random.default_rng(seed)
    A = rng.standard_normal((n, 6, 6)) * 0.2
    H0 = np.zeros((6, 6))
    lam, mu = 115384.6, 76923.1
    H0[:3, :3] = lam
    H0[np.arange(3), np.arange(3)] += 2 * mu
    H0[np.arange(3, 6), np.arange(3, 6)] = mu
    L = np.linalg.cholesky(H0)
    M = L[None] @ (np.eye(6)[None] + 0.5 * (A + A.transpose(0, 2, 1))) @ L.T[None]
    M = 0.5 * (M + M.transpose(0, 2, 1))
    return np.asfortranarray(M.transpose(1, 2, 0))  # (6, 6, N)


def main():
    # ---- 1. cantilever hex8 (tests/test_cantilever_beam_3D_model.py:15-31) + its solve ----
    nodes, elements = meshgen.box_hex8(11, 5, 5, 0, 1000, 0, 100, 0, 100)
    ref_mesh = fd.mesh.box_mesh(11, 5, 5, 0, 1000, 0, 100, 0, 100, "hex8", name="chk")
    assert np.array_equal(ref_mesh.elements, elements) and np.array_equal(ref_mesh.nodes, nodes)
    sets = meshgen.box_node_sets(11, 5, 5)
    for k in ("left", "right", "top", "bottom", "front", "back"):
        assert np.array_equal(np.asarray(ref_mesh.node_sets[k]), sets[k]), k
    # known-answer solve exactly as the reference test does
    fresh("3D")
    mesh = fd.mesh.box_mesh(11, 5, 5, 0, 1000, 0, 100, 0, 100, "hex8", name="Domain")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="ElasticLaw")
    fd.weakform.StressEquilibrium("ElasticLaw", name="weakform")
    fd.Assembly.create("weakform", "Domain", "hex8", name="Assembling")
    pb = fd.problem.Linear("Assembling")
    pb.bc.add("Dirichlet", mesh.node_sets["left"], "DispX", 0)
    pb.bc.add("Dirichlet", mesh.node_sets["left"], "DispY", 0)
    pb.bc.add("Dirichlet", mesh.node_sets["left"], "DispZ", 0)
    pb.bc.add("Dirichlet", mesh.node_sets["right"], "DispY", -10)
    pb.apply_boundary_conditions()
    pb.solve()
    U_sol = pb.get_dof_solution("all").copy()
    asm = fd.Assembly["Assembling"]
    TensorStrain = asm.get_strain(pb.get_dof_solution(), "Node", nlgeom=False)
    TensorStress = fd.ConstitutiveLaw["ElasticLaw"].get_stress_from_strain(asm, TensorStrain)
    assert abs(TensorStress[5][-1] + 0.9007983467254552) < 1e-10
    extra = dict(
        U_sol=U_sol,
        stress_gp_sol=sv_array(asm.sv["Stress"]),
        known_answer=np.array(TensorStress[5][-1]),
    )
    elastic_case("hex8_cantilever", "3D", nodes, elements, "hex8", 200e3, 0.3, extra=extra)

    # ---- 2. jittered hex8 box, isotropic and per-GP tangent ----
    nodes, elements = meshgen.box_hex8(7, 6, 5)
    nodes = meshgen.jitter_nodes(nodes, 7, 6, 5)
    elastic_case("hex8_jitter", "3D", nodes, elements, "hex8", 200e3, 0.3, seed=1)
    H_gp = random_spd_tangent(len(elements) * 8, seed=5)
    elastic_case("hex8_jitter_Hgp", "3D", nodes, elements, "hex8", 0.0, 0.0, seed=2, H_gp=H_gp)

    # ---- 3. tet4 (jittered box split in 6 tets): elasticity + heat ----
    nodes, hexes = meshgen.box_hex8(6, 5, 5)
    nodes = meshgen.jitter_nodes(nodes, 6, 5, 5)
    tets = meshgen.hex8_to_tet4(hexes)
    elastic_case("tet4_box", "3D", nodes, tets, "tet4", 1e5, 0.3, seed=3)
    heat_case("tet4_box_heat", nodes, tets, "tet4")

    # ---- 4. tet10 with curved edges ----
    nodes, hexes = meshgen.box_hex8(5, 4, 4)
    nodes = meshgen.jitter_nodes(nodes, 5, 4, 4)
    n10, e10 = meshgen.tet4_to_tet10(nodes, meshgen.hex8_to_tet4(hexes), bulge=0.03)
    elastic_case("tet10_box", "3D", n10, e10, "tet10", 1e5, 0.3, seed=4)

    # ---- 5. quad4 plate with hole, plane stress (tests/test_platewithhol.py:9-22) ----
    fresh("2Dstress")
    m = fd.mesh.hole_plate_mesh(nr=11, nt=11, length=100, height=100, radius=20, elm_type="quad4", name="tmp")
    elastic_case("quad4_plate", "2Dstress", m.nodes.copy(), m.elements.copy(), "quad4", 2e5, 0.3, seed=6)
    nodes, quads = meshgen.rect_quad4(9, 7)
    nodes = meshgen.jitter_nodes_2d(nodes, 9, 7)
    elastic_case("quad4_jitter_pstrain", "2Dplane", nodes, quads, "quad4", 2e5, 0.3, seed=7)

    # ---- 6. real meshes of the reference: gyroid (tet4) mesh stored, octet (tet10) fingerprint only ----
    fresh("3D")
    fd.mesh.import_file("/root/reference/tests/gyroid.msh", name="Domain")
    gm = fd.Mesh["Domain"]
    gn, ge = gm.nodes.copy(), gm.elements.copy()
    elastic_case("tet4_gyroid", "3D", gn, ge, "tet4", 200e3, 0.3, seed=8, store_K=False)
    heat_case("tet4_gyroid_heat", gn, ge, "tet4", store_mesh=False, store_K=False)

    fp = {}
    fresh("3D")
    fd.mesh.import_file("/root/reference/util/meshes/octet_truss_quad.msh", name="Domain")
    om = fd.Mesh["Domain"].mesh_dict["tet10"]
    om = fd.Mesh(om.nodes, om.elements, "tet10", name="Octet")
    fd.constitutivelaw.ElasticIsotrop(200e3, 0.3, name="law")
    fd.weakform.StressEquilibrium("law", name="wf")
    a = fd.Assembly.create("wf", "Octet", "tet10", name="A")
    pb = fd.problem.Linear("A")
    U = np.random.default_rng(9).standard_normal(pb.n_dof) * 1e-3
    pb.set_X(U)
    a.update(pb, compute="all")
    K = a.global_matrix
    v = np.random.default_rng(109).standard_normal(K.shape[0])
    fp["tet10_octet"] = dict(
        mesh_file="util/meshes/octet_truss_quad.msh",
        n_nodes=int(om.n_nodes),
        n_elements=int(om.n_elements),
        nodes_sha=sha(om.nodes),
        elements_sha=sha(om.elements.astype(np.int32)),
        E=200e3,
        nu=0.3,
        U_seed=9,
        v_seed=109,
        K_shape=list(K.shape),
        K_nnz=int(K.nnz),
        K_indptr_sha=sha(K.indptr),
        K_indices_sha=sha(K.indices),
        K_fro=float(np.linalg.norm(K.data)),
        K_absmax=float(np.abs(K.data).max()),
        Kv_norm=float(np.linalg.norm(K @ v)),
        Kv_head=[float(x) for x in (K @ v)[:8]],
        D_norm=float(np.linalg.norm(a.global_vector)),
        D_head=[float(x) for x in a.global_vector[:8]],
    )
    print("tet10_octet", fp["tet10_octet"]["K_nnz"], fp["tet10_octet"]["K_indptr_sha"])
    with open(os.path.join(OUT, "fingerprints.json"), "w") as f:
        json.dump(fp, f, indent=1)


def heat_case(tag, nodes, elements, elm_type, store_mesh=True, store_K=True):
    """tests/test_thermal3D.py:28-31,75 material; survey Appendix A driving recipe."""
    fresh("3D")
    mesh = fd.Mesh(nodes, elements, elm_type, name="Domain")
    k, c, rho, dt = 500.0, 0.5, 7800.0, 10 / 3
    fd.constitutivelaw.ThermalProperties(k, c, rho, name="ThermalLaw")
    fd.weakform.HeatEquation("ThermalLaw")
    a = fd.Assembly.create("ThermalLaw", "Domain", name="A")
    pb = fd.problem.NonLinear("A")
    rng = np.random.default_rng(3)
    T0 = rng.uniform(0, 3, mesh.n_nodes)
    dT = rng.uniform(-0.1, 0.1, mesh.n_nodes)
    pb.dtime = dt
    pb._U = T0.copy()
    pb._dU = 0
    pb.initialize()
    a.set_start(pb)
    pb._dU = dT
    a.update(pb, "all")
    K, D = a.global_matrix, a.global_vector
    v = np.random.default_rng(103).standard_normal(K.shape[0])
    out = dict(
        k=k, c=c, rho=rho, dt=dt, T_start=T0, T=T0 + dT, D=D,
        temp_gp=np.array(a.sv["Temp"])[:: (1 if store_K else 97)],
        temp_gradient_gp=np.array(a.sv["TempGradient"])[:, :: (1 if store_K else 97)],
        K_shape=np.array(K.shape), K_nnz=K.nnz, K_indptr_sha=sha(K.indptr), K_indices_sha=sha(K.indices),
        K_fro=np.linalg.norm(K.data), K_absmax=np.abs(K.data).max(), v=v, Kv=K @ v,
    )  # fmt: skip
    if store_mesh:
        out.update(nodes=np.asarray(nodes, dtype=float), elements=np.asarray(elements, dtype=np.int32))
    if store_K:
        out.update(K_indptr=K.indptr, K_indices=K.indices, K_data=K.data)
    np.savez_compressed(os.path.join(OUT, tag + ".npz"), **out)
    print(tag, K.shape, K.nnz, out["K_indptr_sha"])


if __name__ == "__main__":
    main()
