"""Results extraction on the device: Gauss-point fields to nodes / elements, von Mises,
and the ``Problem.get_results`` surface the reference's tests assert on.

Mirrors fedoo/core/mesh.py:1149-1160,1267-1308 (``convert_data``: GP -> node through pinv(N_gp) averaged over
the elements around the node; GP -> element: mean over the Gauss points), fedoo/core/output.py:120-330
(``_get_results``: field labels, 'Stress_vm' taken at the Gauss points then converted) and
fedoo/core/dataset.py (``DataSet.node_data / element_data / gausspoint_data``).  The arithmetic runs in
csrc/fdk_results.cuh; host NumPy arrays are produced only at the very end (the reference returns NumPy).
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib
from .core import GaussPointTensor, as_device_f64, device


def node_incidences(mesh):
    """(node_ptr int64 [n_nodes+1], node_inc int32 = element * nne + local node), device, cached on the mesh."""
    cached = getattr(mesh, "_node_inc", None)
    if cached is not None:
        return cached
    conn = mesh.device_arrays()[1].to(torch.int64)
    n_el, nne = conn.shape
    flat = conn.reshape(-1)
    perm = torch.argsort(flat, stable=True)  # grouped by node, element-ascending: a fixed summation order
    counts = torch.bincount(flat, minlength=mesh.n_nodes)
    ptr = torch.zeros(mesh.n_nodes + 1, dtype=torch.int64, device=conn.device)
    ptr[1:] = torch.cumsum(counts, 0)
    assert n_el * nne < 2**31
    mesh._node_inc = (ptr, perm.to(torch.int32).contiguous())
    return mesh._node_inc


_PINV = {}


def pinv_table(elm_type):
    """pinv of the shape functions at the Gauss points (fedoo/core/mesh.py:1149-1151), [nne][ngp]."""
    if elm_type not in _PINV:
        _, N, _ = _lib.element_table(elm_type)  # (ngp, nne)
        _PINV[elm_type] = np.ascontiguousarray(np.linalg.pinv(N))
    return _PINV[elm_type]


def _as_gp_field(data, n_gp):
    """-> (device tensor, ncomp, comp_stride, gp_stride).  Accepts a GaussPointTensor ((N, 6) device), a device or
    host array of shape (N,), (ncomp, N) or (N, ncomp)."""
    if isinstance(data, GaussPointTensor):
        t = data.device_tensor
        return t, 6, 1, 6
    if isinstance(data, (list, tuple)):
        data = np.asarray([np.asarray(d) for d in data])
    t = as_device_f64(data)
    if t.ndim == 1:
        assert t.numel() == n_gp
        return t, 1, 0, 1
    if t.shape[-1] == n_gp:  # (ncomp, N) row-major
        return t, t.shape[0], n_gp, 1
    assert t.shape[0] == n_gp
    return t, t.shape[1], 1, t.shape[1]


def convert_gp(assembly, data, convert_to, von_mises=False):
    """Gauss-point field -> 'Node' | 'Element' | 'GaussPoint' as a device tensor of shape (ncomp_out, n) (or (n,) for
    a scalar field).  ``von_mises``: ``data`` is a 6-component Voigt stress; its von Mises norm is what is converted."""
    mesh = assembly.mesh
    lib = _lib.load()
    n_el, n_nodes = mesh.n_elements, mesh.n_nodes
    nne = mesh.n_elm_nodes
    n_gp = assembly.n_gauss_points
    ngp = n_gp // n_el
    t, ncomp, cs, gs = _as_gp_field(data, n_gp)
    scalar = (ncomp == 1 and t.ndim == 1) or von_mises
    dev = device()
    stream = _lib.current_stream()
    outs = []
    for c0 in range(0, 1 if von_mises else ncomp, 6):  # the kernels convert up to 6 components per pass
        nc = 6 if von_mises else min(6, ncomp - c0)
        base = t.reshape(-1)[c0 * cs :] if c0 else t
        nco = 1 if von_mises else nc
        if convert_to == "Node":
            ptr, inc = node_incidences(mesh)
            out = torch.empty((nco, n_nodes), dtype=torch.float64, device=dev)
            _lib.check(
                lib.fdk_gp_to_node(nne, ngp, n_nodes, n_el, _lib.ptr(ptr), _lib.ptr(inc), _lib.ptr(pinv_table(assembly.elm_type)),
                                   _lib.ptr(base), nc, cs, gs, int(von_mises), _lib.ptr(out), stream),
                "fdk_gp_to_node",
            )  # fmt: skip
        elif convert_to == "Element":
            out = torch.empty((nco, n_el), dtype=torch.float64, device=dev)
            _lib.check(
                lib.fdk_gp_to_element(ngp, n_el, _lib.ptr(base), nc, cs, gs, int(von_mises), _lib.ptr(out), stream),
                "fdk_gp_to_element",
            )
        elif convert_to == "GaussPoint":
            if von_mises:
                out = torch.empty((1, n_gp), dtype=torch.float64, device=dev)
                _lib.check(lib.fdk_gp_von_mises(n_gp, _lib.ptr(base), cs, gs, _lib.ptr(out), stream), "fdk_gp_von_mises")
            else:
                out = torch.stack([t.reshape(-1)[(c0 + c) * cs :][: (n_gp - 1) * gs + 1 : gs] for c in range(nc)])
        else:
            raise ValueError("convert_to must be 'Node', 'Element' or 'GaussPoint'")
        outs.append(out)
    res = outs[0] if len(outs) == 1 else torch.cat(outs, 0)
    return res[0] if scalar else res


class DataSet:
    """The part of fedoo.core.dataset.DataSet the tests read: dictionaries of NumPy arrays by field label."""

    def __init__(self, mesh=None):
        self.mesh = mesh
        self.node_data, self.element_data, self.gausspoint_data, self.scalar_data = {}, {}, {}, {}

    def _store(self, label, data, data_type):
        {"Node": self.node_data, "Element": self.element_data, "GaussPoint": self.gausspoint_data,
         "Scalar": self.scalar_data}[data_type][label] = data  # fmt: skip

    def get_data(self, field, data_type=None):
        for typ, d in (("Node", self.node_data), ("Element", self.element_data), ("GaussPoint", self.gausspoint_data)):
            if (data_type in (None, typ)) and field in d:
                return d[field]
        raise KeyError(field)


def get_results(pb, assemb, output_list, output_type=None):
    """fedoo/core/output.py:_get_results for the fields of the accelerated path: dof variables and vectors ('Disp',
    'DispX', 'Temp', ...), 'Strain', 'Stress', 'Stress_vm', and any Gauss-point array of ``assemb.sv``."""
    if isinstance(output_list, str):
        output_list = [output_list]
    if output_type is not None and output_type not in ("Node", "Element", "GaussPoint"):
        raise NameError("output_type should be either 'Node', 'Element' or 'GaussPoint'")
    result = DataSet(assemb.mesh)
    space = pb.space
    for res in output_list:
        if res in space.list_variables() or res in space.list_vectors():
            data, data_type = np.asarray(pb.get_dof_solution(res)), "Node"
            if output_type not in (None, "Node"):
                raise NotImplementedError("node fields are returned at the nodes")
        elif res in ("Strain", "Stress", "Stress_vm"):
            src = assemb.sv.get(res[:-3] if res.endswith("_vm") else res, 0)
            if np.isscalar(src):
                raise NameError(f'Field "{res}" not available')
            typ = output_type or "GaussPoint"
            data = convert_gp(assemb, src, typ, von_mises=res.endswith("_vm")).cpu().numpy()
            data_type = typ
        elif res in assemb.sv and not np.isscalar(assemb.sv[res]):
            typ = output_type or "GaussPoint"
            data = convert_gp(assemb, assemb.sv[res], typ).cpu().numpy()
            data_type = typ
        else:
            raise NameError(f'Field "{res}" not available')
        result._store(res, data, data_type)
    return result


class NodeTensor(list):
    """Six arrays in Voigt order (the reference's StrainTensorList / StressTensorList at the nodes or elements)."""

    def asarray(self):
        return np.array(self)
