"""Drop-in backend under the REAL fedoo: ``fedoo_b200.install(fedoo)``.

The reference has no plugin interface; its seam is duck-typed Python.  ``install`` replaces, on the
reference's own ``fedoo.core.assembly.Assembly`` class, the two methods behind which the whole hot path sits --
and three more methods whose host cost would otherwise dominate once the assembly is fast:

  * ``Assembly.assemble_global_mat(compute)``   fedoo/core/assembly.py:143-470
        -> one-time symbolic pattern + cluster plan, then the sm_100a kernels of libfdk; results are handed back in the
        reference's own types: ``global_matrix`` a host ``scipy.sparse.csr_matrix`` (int32 indices, explicit zeros kept,
        resized for the problem's global dofs), ``global_vector`` a NumPy array or the scalar 0;
  * ``Assembly.get_gp_results(operator, U)``    fedoo/core/assembly.py:1045-1112  (and with it ``get_grad_disp``
        :1285-1336, ``StressEquilibrium.update`` :191-217, ``SteadyHeatEquation.update`` heat_equation.py:64-70)
        -> Gauss-point values / first derivatives of nodal fields on the device, without the elementary-operator
        matrices the reference builds (37 s at 1 M elements);
  * ``StressEquilibrium.update(assembly, pb)``  fedoo/weakform/stress_equilibrium.py:191-217 (small strain) and
    ``ElasticAnisotropic.update(assembly, pb)`` fedoo/constitutivelaw/elastic_anisotropic.py:36-56 (inherited by
    ``ElasticIsotrop``): grad u, strain and sigma = H eps at the Gauss points come from ONE kernel each and are handed back
    in the reference's own containers (``sv["DispGradient"]`` 3x3 list of arrays, ``StrainTensorList`` /
    ``StressTensorList`` over Fortran-ordered (6, N) arrays); the device copy of the stress is kept for the residual.
    (The reference's own versions cost 0.6 s + 1.5 s per update at 1 M elements -- Python sums over 8 M Gauss points.)
  * ``Problem.solve()``  fedoo/core/problem.py:277-300, only when the problem's solver is ``fedoo_b200.solver.pcg``, the
    matrix is one this adapter assembled and the constraints are Dirichlet conditions and / or PeriodicBC's multi-point
    constraints: the masked / constrained Jacobi-PCG then runs on the matrix still in HBM instead of the host forming
    ``MatCB.T @ A @ MatCB`` (``_problem_solve``).

Everything else -- ``Problem``, boundary conditions, ``DiffOp``, the other constitutive laws, outputs -- stays the reference's own
code and talks to the kernels only through ``assembly.sv`` (``TangentMatrix``, ``Stress``, ``TempGradient``, ``Temp``),
exactly as the reference's ``get_weak_equation`` reads them (weakform/stress_equilibrium.py:92-145,
weakform/heat_equation.py:78-119,168-187).  Supported: small-strain ``StressEquilibrium`` (3D, 2Dplane, 2Dstress; uniform,
per-Gauss-point or isotropic tangent), ``SteadyHeatEquation`` / ``HeatEquation`` (lumped capacity), on hex8 / tet4 /
tet10 / quad4 with the default quadrature.  Anything else raises ``NotImplementedError`` (``strict=True``, the default)
or is handed to the reference's original method (``strict=False``; counted in ``stats["delegated"]``) -- there is no
CPU implementation of the path in this package.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from . import assembly as _asm
from . import constitutivelaw as _law
from . import core as _core
from . import weakform as _wf

_SUPPORTED = {"hex8": (8, 3), "tet4": (4, 3), "tet10": (15, 3), "quad4": (4, 2)}  # default n_elm_gp, ndim
stats = {"assembled": 0, "delegated": 0, "gp_results": 0, "gp_state": 0, "device_solves": 0}
_installed = {}


class _PbShim:
    """What the mirror assembly reads from a problem: global dofs (core/assembly.py:192-197) and the time step."""

    def __init__(self, pb):
        self.n_global_dof = int(getattr(pb, "n_global_dof", 0) or 0) if pb is not None else 0
        self.dtime = getattr(pb, "dtime", 0) if pb is not None else 0


def _mirror_space(dimension):
    keep = _core.ModelingSpace._active
    try:
        sp = _core.ModelingSpace(dimension, name="")
    finally:
        _core.ModelingSpace._active = keep
    return sp


def _shadow_mesh(mesh, elm_type):
    """Device-side twin of a fedoo Mesh, cached on it; rebuilt when the arrays are replaced."""
    sh = getattr(mesh, "_fdk_shadow", None)
    if sh is not None and sh[0] is mesh.nodes and sh[1] is mesh.elements:
        return sh[2]
    twin = _core.Mesh(mesh.nodes, mesh.elements, elm_type)
    mesh._fdk_shadow = (mesh.nodes, mesh.elements, twin)
    return twin


def _geometry_ok(a):
    et = a.elm_type
    if et not in _SUPPORTED or a.mesh.elm_type.lower() != et:
        return False
    n_gp, ndim = _SUPPORTED[et]
    if a.n_elm_gp != n_gp or a.mesh.ndim != ndim or a.space.ndim != ndim:
        return False
    if getattr(a, "_use_local_csys", False):
        return False
    return a.mesh.elements.shape[1] == {"hex8": 8, "tet4": 4, "tet10": 10, "quad4": 4}[et]


def _classify(a):
    """'elastic' | 'heat' | None for a fedoo Assembly."""
    if not _geometry_ok(a) or a.weakform is None:
        return None
    wf = a.weakform
    name = type(wf).__name__
    sp = a.space
    if name == "StressEquilibrium":
        if getattr(a, "_nlgeom", False) or sp.get_dimension() not in ("3D", "2Dplane", "2Dstress"):
            return None
        want = ["DispX", "DispY", "DispZ"][: sp.ndim]
        if sp.nvar != sp.ndim or [sp.variable_rank(v) for v in want] != list(range(sp.ndim)):
            return None
        if getattr(wf, "geometric_stiffness", False):
            return None
        return "elastic"
    parts = getattr(wf, "list_weakform", None) or [wf]
    names = [type(w).__name__ for w in parts]
    if set(names) <= {"SteadyHeatEquation", "TemperatureTimeDerivative"} and len(names) == len(set(names)):
        if sp.nvar != 1 or sp.get_dimension() == "2Daxi":
            return None
        lump = a.mat_lumping if isinstance(a.mat_lumping, (list, tuple)) else [a.mat_lumping] * len(parts)
        for w, lu in zip(parts, lump):
            if type(w).__name__ == "TemperatureTimeDerivative" and not lu:
                return None  # consistent capacity matrix: not on this path
        return "heat"
    return None


def _to_gauss_points(v, a):
    """One material coefficient given at nodes, elements or Gauss points -> Gauss points, with the reference's own rule
    for telling them apart (core/mesh.py:1222-1265: the LAST matching of n_nodes, n_elements, n_gauss_points wins;
    element values are tiled over the gp-major order, nodal values interpolated with the shape functions)."""
    n_gp, mesh = a.n_gauss_points, a.mesh
    if v.shape[-1] == n_gp:
        return v
    if v.shape[-1] == mesh.n_elements:
        return np.tile(v, a.n_elm_gp)
    if v.shape[-1] == mesh.n_nodes:
        _, N, _ = _lib.element_table(a.elm_type)  # (n_elm_gp, nne)
        return np.einsum("gk,ek->ge", N, v[np.asarray(mesh.elements)]).reshape(-1)
    raise ValueError("data doesn't match with the number of nodes, number of elements or number of gauss points.")


def _normalize_tangent(H, a):
    """sv['TangentMatrix'] in any of the reference's formats (6x6 floats, 6x6 object array / list of lists with
    per-Gauss-point, per-element or per-node entries, (6,6,N) ndarray -- all indexed H[i][j],
    stress_equilibrium.py:112-117) -> (6,6) float array or (6,6,N) Fortran-ordered array."""
    n_gp = a.n_gauss_points
    if isinstance(H, np.ndarray) and H.dtype != object:
        if H.ndim == 2:
            return np.ascontiguousarray(H, dtype=np.float64)
        if H.ndim == 3 and H.shape[2] == n_gp:
            return np.asfortranarray(H, dtype=np.float64)
        raise NotImplementedError(f"tangent of shape {H.shape}")
    per_gp = any(np.ndim(H[i][j]) > 0 for i in range(6) for j in range(6))
    if not per_gp:
        return np.array([[float(H[i][j]) for j in range(6)] for i in range(6)])
    out = np.zeros((6, 6, n_gp), order="F")
    for i in range(6):
        for j in range(6):
            v = np.asarray(H[i][j], dtype=np.float64)
            out[i, j, :] = _to_gauss_points(v, a) if v.ndim > 0 else v
    return out


class _Backend:
    """Mirror objects of one fedoo Assembly (mesh twin, weak form, law) through which the kernels are called."""

    def __init__(self, a, kind):
        self.kind = kind
        self.mesh_ref = a.mesh
        self.twin = _shadow_mesh(a.mesh, a.elm_type)
        self.space = _mirror_space(a.space.get_dimension())
        self.key = None
        self.asm = None
        # Gauss-point state computed on the device by the patched update methods: the host objects handed to the
        # reference (identity-checked before a device copy is trusted) and what they were made from
        self.state = None
        self.host_index = None  # host copy of the pattern's index arrays (fetched once)

    def valid_for(self, a):
        return self.mesh_ref is a.mesh and self.twin is _shadow_mesh(a.mesh, a.elm_type)

    def elastic(self, a, H):
        law = a.weakform.constitutivelaw
        iso = None
        if H.ndim == 2 and type(law).__name__ == "ElasticIsotrop" and np.isscalar(law.E) and np.isscalar(law.nu):
            m = _law.ElasticIsotrop(float(law.E), float(law.nu))
            if np.array_equal(m.get_tangent_matrix(None, a.space.get_dimension()), H):
                iso = (float(law.E), float(law.nu))
        key = ("iso", iso) if iso else ("gen", H.ndim)
        if self.asm is None or self.key != key:
            mlaw = _law.ElasticIsotrop(*iso) if iso else _law.ElasticAnisotropic(H)
            mwf = _wf.StressEquilibrium(mlaw, name="", space=self.space)
            self.asm = _asm.Assembly(mwf, self.twin, a.elm_type, "")
            self.key = key
        if not iso:
            self.asm.weakform.constitutivelaw._H = H
        self.asm.sv["TangentMatrix"] = H
        return self.asm

    def heat(self, a, cond, rho_c, transient):
        if self.asm is None:
            mlaw = _law.ThermalProperties(cond, 1.0, 1.0)
            cls = _wf.HeatEquation if transient else _wf.SteadyHeatEquation
            self.asm = _asm.Assembly(cls(mlaw, name="", space=self.space), self.twin, a.elm_type, "")
        mlaw = self.asm.weakform.constitutivelaw
        mlaw.thermal_conductivity, mlaw.specific_heat, mlaw.density = cond, 1.0, rho_c
        return self.asm


def _backend(a, kind):
    be = a.__dict__.get("_fdk_backend")
    if be is None or be.kind != kind or not be.valid_for(a):
        be = _Backend(a, kind)
        a._fdk_backend = be
    return be


_device_K = {}  # id(host csr_matrix handed to the reference) -> (weakref to it, the DeviceCSR it was fetched from)


def _remember_device_matrix(host, dev_csr):
    import weakref

    key = id(host)
    _device_K[key] = (weakref.ref(host, lambda _r, k=key: _device_K.pop(k, None)), dev_csr)


def _device_matrix_of(host):
    rec = _device_K.get(id(host))
    return rec[1] if rec is not None and rec[0]() is host else None


_CHUNK = 64 * 1024 * 1024  # elements of the pinned staging buffers (512 MB of float64)
_staging = {}


def _fetch(dev_tensor):
    """A device tensor into a FRESH NumPy array, through two pinned staging buffers (full PCIe rate; the copy out of a
    buffer, multi-threaded by torch, overlaps the next chunk's transfer).  ``tensor.cpu()`` goes through pageable memory:
    1.4 s for the 2.9 GB of a 1 M-element K against 0.3 s this way."""
    n = dev_tensor.numel()
    out = np.empty(n, dtype={torch.float64: np.float64, torch.int32: np.int32, torch.int64: np.int64}[dev_tensor.dtype])
    if n == 0:
        return out
    dst = torch.from_numpy(out)
    key = dev_tensor.dtype
    bufs = _staging.get(key)
    if bufs is None or bufs[0].numel() < min(n, _CHUNK):
        bufs = _staging[key] = [torch.empty(min(n, _CHUNK), dtype=key, pin_memory=True) for _ in range(2)]
    step = bufs[0].numel()
    src = dev_tensor.reshape(-1)
    stream = torch.cuda.current_stream()
    events = [None, None]
    chunks = [(o, min(o + step, n)) for o in range(0, n, step)]
    for k, (o, e) in enumerate(chunks):  # issue chunk k, then drain chunk k - 1 while k is in flight
        b = k & 1
        bufs[b][: e - o].copy_(src[o:e], non_blocking=True)
        events[b] = torch.cuda.Event()
        events[b].record(stream)
        if k > 0:
            po, pe = chunks[k - 1]
            events[b ^ 1].synchronize()
            dst[po:pe].copy_(bufs[b ^ 1][: pe - po])
    po, pe = chunks[-1]
    events[(len(chunks) - 1) & 1].synchronize()
    dst[po:pe].copy_(bufs[(len(chunks) - 1) & 1][: pe - po])
    return out


def _to_host_csr(dev_csr, be):
    """The assembled matrix as the host ``scipy.sparse.csr_matrix`` the reference's solvers expect.  The index arrays of a
    pattern are fetched from HBM once per backend and every matrix gets its own copy of them (as with the reference, a
    caller may edit one matrix in place without touching the next)."""
    from scipy import sparse

    key = (dev_csr.indices.data_ptr(), dev_csr.indptr.data_ptr(), tuple(dev_csr.shape))
    if be.host_index is None or be.host_index[0] != key:
        be.host_index = (key, _fetch(dev_csr.indptr), _fetch(dev_csr.indices))
    _, indptr, indices = be.host_index
    idx = np.empty_like(indices)
    torch.from_numpy(idx).copy_(torch.from_numpy(indices))  # multi-threaded host copy
    host = sparse.csr_matrix((_fetch(dev_csr.data), idx, indptr.copy()), shape=dev_csr.shape)
    _remember_device_matrix(host, dev_csr)  # Problem.solve with the device solver finds K still in HBM
    return host


def _assemble(a, compute, strict, orig):
    if compute == "none":
        return
    kind = _classify(a)
    if kind is None:
        if strict:
            raise NotImplementedError(
                f"fedoo_b200: assembly '{getattr(a, 'name', '')}' ({type(a.weakform).__name__}, {a.elm_type}) is not on "
                "the accelerated path (install(..., strict=False) hands it to the reference's own method)"
            )
        stats["delegated"] += 1
        return orig(a, compute)
    if compute not in ("all", "matrix", "vector"):
        raise ValueError("compute must be 'all', 'matrix', 'vector' or 'none'")
    if a.meshChange:
        a.mesh.__dict__.pop("_fdk_shadow", None)  # node positions changed in place (core/assembly.py:160-165)
    be = _backend(a, kind)
    pb = _PbShim(a._pb)
    n_gp = a.n_gauss_points
    want_mat, want_vec = compute != "vector", compute != "matrix"
    if kind == "elastic":
        H = _normalize_tangent(a.sv["TangentMatrix"], a)
        m = be.elastic(a, H)
        m._pb = pb
        stress = a.sv.get("Stress", 0)
        if np.isscalar(stress) and stress == 0:
            m.sv["Stress"] = 0
        elif be.state is not None and be.state.get("stress_host") is stress:
            m.sv["Stress"] = _core.GaussPointTensor(be.state["stress_dev"], "stress")  # made by _law_update: still in HBM
        else:  # (6, N) -> (N, 6) on the device: the layout of the reference's F-ordered array
            s = np.ascontiguousarray(np.asarray(stress.asarray(), dtype=np.float64).T)
            m.sv["Stress"] = _core.GaussPointTensor(torch.from_numpy(s).to(_core.device()), "stress")
        m.assemble_global_mat(compute)
        if want_mat:
            a.global_matrix = _to_host_csr(m.global_matrix, be)
        if want_vec:
            a.global_vector = m.global_vector if np.isscalar(m.global_vector) else np.array(m.global_vector)
    else:
        wf = a.weakform
        parts = getattr(wf, "list_weakform", None) or [wf]
        steady = next((w for w in parts if type(w).__name__ == "SteadyHeatEquation"), None)
        timed = next((w for w in parts if type(w).__name__ == "TemperatureTimeDerivative"), None)
        tlaw = (steady or timed).constitutivelaw
        cond = np.zeros((3, 3))
        if steady is not None:
            k = tlaw.thermal_conductivity
            cond = np.array([[float(k[i][j]) for j in range(3)] for i in range(3)])
        rho_c = float(tlaw.density * tlaw.specific_heat) if timed is not None else 0.0
        m = be.heat(a, cond, rho_c, timed is not None)
        m._pb = pb
        rcdt = rho_c / pb.dtime if (timed is not None and pb.dtime != 0) else 0.0
        if want_mat:
            m.assemble_global_mat("matrix")
            a.global_matrix = _to_host_csr(m.global_matrix, be)
        if want_vec:
            a.global_vector = _heat_vector(a, m, steady, timed, cond, rcdt, pb.n_global_dof)
    if a._saved_bloc_structure is None:
        a._saved_bloc_structure = m._saved_bloc_structure  # symbolic reuse marker (core/assembly.py:469-470)
    stats["assembled"] += 1


def _heat_vector(a, m, steady, timed, cond, rcdt, n_glob):
    """-int [grad v . (K TempGradient) + (rho c / dt) v (Temp - Temp_start)] from the fields in assembly.sv."""
    dev = _core.device()
    n_gp = a.n_gauss_points
    flux = src = None
    if steady is not None:
        g = a.sv.get("TempGradient", [0, 0, 0])
        if any(not np.array_equal(x, 0) for x in g):
            G = torch.zeros((3, n_gp), dtype=torch.float64, device=dev)
            for j in range(3):
                if not np.array_equal(g[j], 0):
                    G[j] = torch.from_numpy(np.ascontiguousarray(g[j], dtype=np.float64)).to(dev)
            flux = (torch.from_numpy(cond).to(dev) @ G).contiguous()
    if timed is not None and rcdt != 0.0:
        t0 = getattr(timed, "_TemperatureTimeDerivative__temp_start", 0)
        dT = a.sv.get("Temp", 0) - t0
        if not np.array_equal(dT, 0):
            src = (torch.from_numpy(np.ascontiguousarray(dT, dtype=np.float64)).to(dev) * rcdt).contiguous()
    if flux is None and src is None:
        return 0
    from .results import node_incidences

    node_ptr, node_inc = node_incidences(m.mesh)
    coords, conn = m.mesh.device_arrays()
    n_nodes = m.mesh.n_nodes
    fe = torch.empty(m.mesh.n_elements * conn.shape[1], dtype=torch.float64, device=dev)
    D = torch.zeros(n_nodes + n_glob, dtype=torch.float64, device=dev)
    _lib.check(
        _lib.load().fdk_residual_heat_gp(
            _lib.ELEM_IDS[a.elm_type], n_nodes, m.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords), _lib.ptr(flux),
            _lib.ptr(src), _lib.ptr(node_ptr), _lib.ptr(node_inc), _lib.ptr(fe), _lib.ptr(D), _lib.current_stream(),
        ),
        "fdk_residual_heat_gp",
    )  # fmt: skip
    return D.cpu().numpy()


def _gp_state(a, U_dev, fbar, H=None):
    """grad u (9, N), strain (N, 6) [H None] or stress (N, 6) [H given: (6,6) or (6,6,N)] of a fedoo Assembly on the device."""
    twin = _shadow_mesh(a.mesh, a.elm_type)
    coords, conn = twin.device_arrays()
    N, dev = a.n_gauss_points, _core.device()
    lib = _lib.load()
    grad = strain = stress = C_h = tan = center = None
    if H is None:
        grad = torch.empty((9, N), dtype=torch.float64, device=dev)
        strain = torch.empty((N, 6), dtype=torch.float64, device=dev)
    else:
        stress = torch.empty((N, 6), dtype=torch.float64, device=dev)
        if H.ndim == 2:
            C_h = np.ascontiguousarray(H, dtype=np.float64)
        else:
            tan = torch.from_numpy(np.asfortranarray(H).reshape(-1, order="F")).to(dev)
    args = [_lib.ELEM_IDS[a.elm_type], twin.n_nodes, twin.n_elements, _lib.ptr(conn), _lib.ptr(coords), _lib.ptr(U_dev),
            _lib.ptr(C_h), _lib.ptr(tan)]  # fmt: skip
    if fbar:
        center = torch.empty(twin.n_elements, dtype=torch.float64, device=dev)
        rc = lib.fdk_gp_strain_stress_fbar(*args, _lib.ptr(center), _lib.ptr(grad), _lib.ptr(strain), _lib.ptr(stress),
                                           _lib.current_stream())  # fmt: skip
    else:
        rc = lib.fdk_gp_strain_stress(*args, _lib.ptr(grad), _lib.ptr(strain), _lib.ptr(stress), _lib.current_stream())
    _lib.check(rc, "fdk_gp_strain_stress")
    return grad, strain, stress


def _wf_update(wf, a, pb, orig, StrainTensorList):
    """StressEquilibrium.update, small strain: DispGradient and Strain from one kernel (stress_equilibrium.py:191-217,
    485-540, 589-602)."""
    U = pb.get_dof_solution() if type(a).__name__ == "Assembly" else 0
    fbar = bool(getattr(wf, "fbar", False))
    if (np.isscalar(U) or getattr(a, "_nlgeom", False) or _classify(a) != "elastic"
            or (fbar and a.elm_type == "quad4")):  # fmt: skip
        return orig(wf, a, pb)
    be = _backend(a, "elastic")
    n, ndim = a.mesh.n_nodes, a.space.ndim
    U_dev = _core.as_device_f64(np.asarray(U, dtype=np.float64)[: ndim * n])
    grad, strain, _ = _gp_state(a, U_dev, fbar)
    g = _fetch(grad).reshape(9, -1)
    rows = [[g[3 * i + j] if (i < ndim and j < ndim) else 0 for j in range(3)] for i in range(3)]
    # the reference keeps a list of lists (get_grad_disp), or an array once the F-bar correction went through np.array
    a.sv["DispGradient"] = np.array(rows) if fbar else rows
    # the reference's own container, over an (N, 6) buffer seen as the Fortran-ordered (6, N) array
    eps = a.sv["Strain"] = StrainTensorList(_fetch(strain).reshape(-1, 6).T)
    be.state = {"U": U_dev, "fbar": fbar, "strain_host": eps, "stress_host": None, "stress_dev": None}
    stats["gp_state"] += 1


def _law_update(law, a, pb, orig, StressTensorList):
    """ElasticAnisotropic.update (elastic_anisotropic.py:36-56): sigma = H eps at the Gauss points, from the dof vector
    the strain was made of (one kernel; the reference's version is a Python sum over six arrays per component)."""
    be = a.__dict__.get("_fdk_backend") if type(a).__name__ == "Assembly" else None  # not the sub-assemblies of Heterogeneous
    st = be.state if be is not None else None
    if (st is None or a.sv.get("Strain") is not st["strain_host"] or "DStrain" in a.sv or getattr(a, "_nlgeom", False)
            or _classify(a) != "elastic"):  # fmt: skip
        return orig(law, a, pb)
    if "TangentMatrix" in a.sv:  # linear problem: no need to recompute it (the reference's own rule)
        H = a.sv["TangentMatrix"]
    else:
        H = a.sv["TangentMatrix"] = law.get_tangent_matrix(a)
    _, _, stress = _gp_state(a, st["U"], st["fbar"], _normalize_tangent(H, a))
    sig = a.sv["Stress"] = StressTensorList(_fetch(stress).reshape(-1, 6).T)
    st["stress_host"], st["stress_dev"] = sig, stress
    stats["gp_state"] += 1


def _gp_results(a, operator, U, n_elm_gp, use_local_dof, orig):
    """Assembly.get_gp_results for operators made of nodal variables and their first derivatives."""
    ok = _geometry_ok(a) and not use_local_dof and (n_elm_gp is None or n_elm_gp == a.n_elm_gp)
    ok = ok and not (np.isscalar(U)) and all(np.isscalar(c) for c in operator.coef)
    ok = ok and all(o.ordre in (0, 1) and np.isscalar(o.x) for o in operator.op) and not a._get_associated_variables()
    if not ok:
        return orig(a, operator, U, n_elm_gp, use_local_dof)
    twin = _shadow_mesh(a.mesh, a.elm_type)
    coords, conn = twin.device_arrays()
    n, N = twin.n_nodes, a.n_gauss_points
    dev = _core.device()
    U_dev = _core.as_device_f64(np.asarray(U, dtype=np.float64)[: a.space.nvar * n])
    lib = _lib.load()
    fields = {}
    res = 0
    for o, c, ov in zip(operator.op, operator.coef, operator.op_vir):
        assert ov == 1, "Operator virtual are only required to build FE operators, but not to get element results"
        if o.u not in fields:
            temp = torch.empty(N, dtype=torch.float64, device=dev)
            grad = torch.empty((3, N), dtype=torch.float64, device=dev)
            _lib.check(
                lib.fdk_gp_temperature(
                    _lib.ELEM_IDS[a.elm_type], n, twin.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                    _lib.ptr(U_dev[o.u * n : (o.u + 1) * n]), _lib.ptr(temp), _lib.ptr(grad), _lib.current_stream(),
                ),
                "fdk_gp_temperature",
            )  # fmt: skip
            fields[o.u] = (temp, grad)
        f = fields[o.u][0] if o.ordre == 0 else fields[o.u][1][o.x]
        res = res + c * f
    stats["gp_results"] += 1
    return _fetch(res)


def _mpc_map(pb, n_nodal, n_glob):
    """The reference's multi-point constraints (``pb._MFext``: X_slave = M X + Xbc after ``M + M @ M``,
    fedoo/core/problem.py:375-393) as the device constraint map of csrc/fdk_solve.cuh, when they have the form PeriodicBC
    generates: every eliminated dof follows ONE free nodal dof with factor 1 plus any combination of the global
    (mean-strain) dofs.  Entries that point at imposed or eliminated dofs carry no unknown (their share is in Xbc
    already).  Returns None when a constraint does not fit."""
    from . import constraint as _constraint

    M = pb._MFext.tocoo()
    blocked = np.fromiter(pb._dof_blocked, dtype=np.int64, count=len(pb._dof_blocked))
    eliminated = np.zeros(n_nodal + n_glob, dtype=bool)
    eliminated[np.asarray(pb._dof_slave, dtype=np.int64)] = True  # Dirichlet dofs and MPC slaves
    is_blocked = np.zeros(n_nodal + n_glob, dtype=bool)
    is_blocked[blocked] = True
    slaves = np.flatnonzero(eliminated & ~is_blocked)
    if slaves.size == 0 or slaves.max() >= n_nodal:
        return None
    pos = np.full(n_nodal + n_glob, -1, dtype=np.int64)
    pos[slaves] = np.arange(slaves.size)
    r, c, v = M.row.astype(np.int64), M.col.astype(np.int64), M.data
    if (pos[r] < 0).any():
        return None  # a constraint row that is not an eliminated dof
    glob = c >= n_nodal
    coef = np.zeros((slaves.size, n_glob))
    np.add.at(coef, (pos[r[glob]], c[glob] - n_nodal), v[glob])
    coef[:, is_blocked[n_nodal:]] = 0.0  # imposed global dofs: in Xbc
    nodal = ~glob & ~eliminated[c]
    master = np.full(slaves.size, -1, dtype=np.int64)
    count = np.bincount(pos[r[nodal]], minlength=slaves.size)
    if (count != 1).any() or not np.all(v[nodal] == 1.0):
        return None
    master[pos[r[nodal]]] = c[nodal]
    return _constraint.MpcMap(n_nodal, n_glob, slaves, master, coef)


def _problem_solve(pb, orig, kargs):
    """Problem.solve (fedoo/core/problem.py:277-300) when the problem's solver is ``fedoo_b200.solver.pcg``, the matrix
    is one this adapter assembled (its device copy is still in HBM) and the constraints are Dirichlet conditions (MatCB
    is then a selection of the free dofs) and / or multi-point constraints of the periodic form (``_mpc_map``):
    MatCB^T K MatCB y = MatCB^T (B + D - K Xbc) by the masked / constrained Jacobi-PCG on the device, matrix-free in the
    constraints, without forming the reduced matrix on the host or sending K anywhere.  Anything else goes to the
    reference's own method, which hands the host reduced system to the solver callable."""
    from . import solver as _solver

    spec = getattr(pb, "_ProblemBase__solver", None)
    A = getattr(pb, "_Problem__A", None)
    K = _device_matrix_of(A) if A is not None else None
    if not all(hasattr(pb, name) for name in ("_Problem__B", "_Problem__D", "_Xbc", "_dof_free", "_dof_slave")):
        return orig(pb, **kargs)  # not the Problem this short-cut was written against (fedoo/core/problem.py:56-70)
    func = spec[1] if spec is not None and len(spec) >= 3 else None
    bound = dict(getattr(func, "keywords", None) or {})  # functools.partial(fedoo_b200.solver.pcg, rtol=...) is fine too
    if (getattr(func, "func", func) is not _solver.pcg or K is None or len(pb._dof_free) == 0
            or not hasattr(A, "shape") or len(A.shape) != 2):  # fmt: skip
        return orig(pb, **kargs)
    n = pb.n_dof
    n_glob = int(getattr(pb, "n_global_dof", 0) or 0)
    n_mat = int(K.shape[0])  # n, or the nodal size when the matrix was assembled before the global dofs existed (the
    # reference resizes it at this point, core/problem.py:283-284: empty trailing rows and columns)
    mpc = None
    if getattr(pb, "_MFext", None) is not None:  # multi-point constraints: the periodic form runs on the device too
        ok = K.block is not None and ((n_mat == n and K.n_glob == n_glob) or (n_mat == n - n_glob and K.n_glob == 0))
        mpc = _mpc_map(pb, n - n_glob, n_glob) if ok else None
        if mpc is None:
            return orig(pb, **kargs)
    elif n_mat != n:
        return orig(pb, **kargs)
    opts = {**bound, **{k: v for k, v in spec[2].items() if v is not None}}
    rtol = opts.get("rtol", opts.get("tol", 1e-8))
    dev = K.data.device
    Xbc = _core.as_device_f64(np.asarray(pb._Xbc, dtype=np.float64), dev)  # already expanded through the constraints
    rhs = torch.zeros(n, dtype=torch.float64, device=dev)
    rhs[:n_mat] -= K.matvec(Xbc[:n_mat] if n_mat < n else Xbc)[:n_mat]
    for v in (pb._Problem__B, pb._Problem__D):
        if not (np.isscalar(v) and v == 0):
            vd = _core.as_device_f64(np.asarray(v, dtype=np.float64), dev)
            rhs[: vd.numel()] += vd
    free = torch.zeros(n, dtype=torch.uint8, device=dev)
    free[torch.from_numpy(np.asarray(pb._dof_free, dtype=np.int64)).to(dev)] = 1
    if mpc is not None:  # T^T K T y = T^T (B + D - K Xbc), X = T y + Xbc (fedoo/core/problem.py:286-298 with MatCB = T)
        load = rhs[mpc.n_nodal :].clone()  # loads on the global dofs (fold overwrites those rows)
        mpc.fold(rhs)
        rhs[mpc.n_nodal :] += load
    x, it, rel = K.pcg(rhs, free_mask=free, rtol=rtol, maxiter=opts.get("maxiter"), check_every=opts.get("check_every", 10),
                       mpc=mpc)  # fmt: skip
    if mpc is not None:
        mpc.expand(x)
        free = 1  # every entry of T y is meaningful
    _solver.info["iterations"], _solver.info["relative_residual"], _solver.info["on_device_matrix"] = it, rel, True
    if rel > rtol:
        print(f"Warning: fedoo_b200.solver.pcg: convergence to tolerance not achieved ({rel:.2e} after {it} iterations)")
    pb._Problem__X = _fetch(x * free + Xbc)  # x is 0 on the imposed dofs already; the mask makes it explicit
    stats["device_solves"] += 1


def install(fedoo=None, strict=True, state_updates=True):
    """Put the CUDA path under ``fedoo.Assembly`` (the reference package, unmodified).  Returns the module.
    ``state_updates=False`` leaves ``StressEquilibrium.update`` / ``ElasticAnisotropic.update`` to the reference's own
    host code (the kernels then only serve ``assemble_global_mat`` and ``get_gp_results``)."""
    if fedoo is None:
        import fedoo
    _lib.load()  # fail now, loudly, when the extension is missing
    A = fedoo.core.assembly.Assembly
    if id(A) in _installed:
        _installed[id(A)]["strict"][0] = strict
        _installed[id(A)]["state"][0] = state_updates
        return fedoo
    W = fedoo.weakform.stress_equilibrium.StressEquilibrium
    L = fedoo.constitutivelaw.elastic_anisotropic.ElasticAnisotropic
    P = fedoo.core.problem.Problem
    TL = fedoo.util.voigt_tensors  # the reference's own containers of Gauss-point tensors
    orig_asm, orig_gp, orig_wf, orig_law, orig_solve = A.assemble_global_mat, A.get_gp_results, W.update, L.update, P.solve
    flag, state = [strict], [state_updates]

    def assemble_global_mat(self, compute="all"):
        return _assemble(self, compute, flag[0], orig_asm)

    def get_gp_results(self, operator, U, n_elm_gp=None, use_local_dof=False):
        return _gp_results(self, operator, U, n_elm_gp, use_local_dof, orig_gp)

    def wf_update(self, assembly, pb):
        return _wf_update(self, assembly, pb, orig_wf, TL.StrainTensorList) if state[0] else orig_wf(self, assembly, pb)

    def law_update(self, assembly, pb):
        return _law_update(self, assembly, pb, orig_law, TL.StressTensorList) if state[0] else orig_law(self, assembly, pb)

    def problem_solve(self, **kargs):
        return _problem_solve(self, orig_solve, kargs)

    orig_del = A.__dict__.get("delete_memory")  # a staticmethod object (core/assembly.py:755-774)

    def delete_memory():
        (orig_del.__func__ if isinstance(orig_del, staticmethod) else orig_del)()
        _asm.Assembly.delete_memory()  # patterns, plans and device meshes cached behind the reference's assemblies
        _device_K.clear()

    problem_solve.__doc__ = orig_solve.__doc__
    assemble_global_mat.__doc__ = orig_asm.__doc__
    get_gp_results.__doc__ = orig_gp.__doc__
    wf_update.__doc__ = orig_wf.__doc__
    law_update.__doc__ = orig_law.__doc__
    A.assemble_global_mat = assemble_global_mat
    A.get_gp_results = get_gp_results
    W.update = wf_update
    L.update = law_update
    P.solve = problem_solve
    if orig_del is not None:
        A.delete_memory = staticmethod(delete_memory)
    _installed[id(A)] = {"cls": A, "orig": (orig_asm, orig_gp), "strict": flag, "state": state,
                         "wf": (W, orig_wf), "law": (L, orig_law), "solve": (P, orig_solve), "del": orig_del}  # fmt: skip
    return fedoo


def uninstall(fedoo=None):
    if fedoo is None:
        import fedoo
    A = fedoo.core.assembly.Assembly
    rec = _installed.pop(id(A), None)
    if rec is not None:
        A.assemble_global_mat, A.get_gp_results = rec["orig"]
        rec["wf"][0].update = rec["wf"][1]
        rec["law"][0].update = rec["law"][1]
        rec["solve"][0].solve = rec["solve"][1]
        if rec["del"] is not None:
            A.delete_memory = rec["del"]
