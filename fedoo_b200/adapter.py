"""Drop-in backend under the REAL fedoo: ``fedoo_b200.install(fedoo)``.

The reference has no plugin interface; its seam is duck-typed Python.  ``install`` replaces, on the
reference's own ``fedoo.core.assembly.Assembly`` class, the two methods behind which the whole hot path sits:

  * ``Assembly.assemble_global_mat(compute)``   fedoo/core/assembly.py:143-470
        -> one-time symbolic pattern + cluster plan, then the sm_100a kernels of libfdk; results are handed back in the
        reference's own types: ``global_matrix`` a host ``scipy.sparse.csr_matrix`` (int32 indices, explicit zeros kept,
        resized for the problem's global dofs), ``global_vector`` a NumPy array or the scalar 0;
  * ``Assembly.get_gp_results(operator, U)``    fedoo/core/assembly.py:1045-1112  (and with it ``get_grad_disp``
        :1285-1336, ``StressEquilibrium.update`` :191-217, ``SteadyHeatEquation.update`` heat_equation.py:64-70)
        -> Gauss-point values / first derivatives of nodal fields on the device, without the elementary-operator
        matrices the reference builds (37 s at 1 M elements).

Everything else -- ``Problem``, boundary conditions, ``DiffOp``, constitutive laws, outputs -- stays the reference's own
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
stats = {"assembled": 0, "delegated": 0, "gp_results": 0}
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


def _to_host_csr(dev_csr):
    return dev_csr.tocsr()


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
        else:  # (6, N) -> (N, 6) on the device: the layout of the reference's F-ordered array
            s = np.ascontiguousarray(np.asarray(stress.asarray(), dtype=np.float64).T)
            m.sv["Stress"] = _core.GaussPointTensor(torch.from_numpy(s).to(_core.device()), "stress")
        m.assemble_global_mat(compute)
        if want_mat:
            a.global_matrix = _to_host_csr(m.global_matrix)
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
            a.global_matrix = _to_host_csr(m.global_matrix)
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
    return res.cpu().numpy()


def install(fedoo=None, strict=True):
    """Put the CUDA path under ``fedoo.Assembly`` (the reference package, unmodified).  Returns the module."""
    if fedoo is None:
        import fedoo
    _lib.load()  # fail now, loudly, when the extension is missing
    A = fedoo.core.assembly.Assembly
    if id(A) in _installed:
        _installed[id(A)]["strict"][0] = strict
        return fedoo
    orig_asm, orig_gp = A.assemble_global_mat, A.get_gp_results
    flag = [strict]

    def assemble_global_mat(self, compute="all"):
        return _assemble(self, compute, flag[0], orig_asm)

    def get_gp_results(self, operator, U, n_elm_gp=None, use_local_dof=False):
        return _gp_results(self, operator, U, n_elm_gp, use_local_dof, orig_gp)

    assemble_global_mat.__doc__ = orig_asm.__doc__
    get_gp_results.__doc__ = orig_gp.__doc__
    A.assemble_global_mat = assemble_global_mat
    A.get_gp_results = get_gp_results
    _installed[id(A)] = {"cls": A, "orig": (orig_asm, orig_gp), "strict": flag}
    return fedoo


def uninstall(fedoo=None):
    if fedoo is None:
        import fedoo
    A = fedoo.core.assembly.Assembly
    rec = _installed.pop(id(A), None)
    if rec is not None:
        A.assemble_global_mat, A.get_gp_results = rec["orig"]
