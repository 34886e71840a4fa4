"""``Assembly``: drop-in mirror of fedoo's global-operator assembly
(fedoo/core/assembly.py:33-1663) whose ``assemble_global_mat`` runs the sm_100a cluster
kernels of libfdk instead of NumPy/SciPy.

Boundary kept:
  * ``Assembly.create(weakform, mesh="", elm_type="", name="", **kargs)`` (assembly.py:1566)
  * ``assemble_global_mat(compute)`` sets ``global_matrix`` / ``global_vector`` (:143-470);
    ``compute`` in {"all", "matrix", "vector", "none"}; the vector is the scalar 0 when the weak
    form has no vector term (:462-463)
  * ``get_global_matrix() / get_global_vector()`` (fedoo/core/base.py:138-148)
  * ``initialize / set_start / update / to_start / reset`` and the ``sv`` / ``sv_start`` dicts
    (:672-753), ``get_gp_results``-style state (``sv["Strain"]``, ``sv["Stress"]``,
    ``sv["DispGradient"]``, ``sv["TempGradient"]``, ``sv["Temp"]``)
The matrix is returned as a device-resident ``DeviceCSR`` (materialised to scipy on demand),
the vector as a NumPy array (device copy in ``global_vector_device``).
"""

from __future__ import annotations

import os

import ctypes as C

import numpy as np
import torch

from . import _lib, symbolic
from .constitutivelaw import ElasticAnisotropic, ElasticIsotrop, ElastoPlasticity
from .core import DeviceCSR, GaussPointTensor, Mesh, _Named, as_device_f64, device
from .weakform import WeakFormBase

_LAZY = object()  # a per-Gauss-point tangent exists in structured form; the (6,6,N) array is expanded only on demand
_DEFAULT_NGP = {"hex8": 8, "tet4": 4, "tet10": 15, "quad4": 4}  # fedoo/lib_elements/element_list.py:50-84


_RESIDUAL_KERNEL = os.environ.get("FDK_RESIDUAL_KERNEL", "1") != "0"  # 0: residual-only through the cluster kernels
_TET10_BIG = os.environ.get("FDK_TET10_BIG", "0") != "0"  # tet10 + isotropic law through the balanced 1024-thread kernel
_HEAT_TET4_ROWS = os.environ.get("FDK_HEAT_TET4_ROWS", "1") != "0"  # 0: tet4 heat through the cluster kernel


class Assembly(_Named):
    _dict = {}
    # class-level symbolic cache keyed by (mesh object, element type), like the reference's
    # _saved_elementary_operators (fedoo/core/assembly.py:56,928)
    _saved_plans = {}

    @staticmethod
    def create(weakform, mesh="", elm_type="", name="", **kargs):
        return Assembly(weakform, mesh, elm_type, name, **kargs)

    @staticmethod
    def delete_memory():
        """fedoo/core/assembly.py:755-774."""
        Assembly._saved_plans = {}

    def __init__(self, weakform, mesh="", elm_type="", name="", **kargs):
        if isinstance(weakform, str):
            weakform = WeakFormBase.get_all()[weakform]
        if isinstance(mesh, str):
            mesh = Mesh.get_all()[mesh]
        if not type(mesh) == Mesh:
            raise TypeError("mesh should refers to a fedoo.Mesh object")
        self.weakform = weakform
        self.space = weakform.space
        self.current = self
        self.meshChange = kargs.pop("MeshChange", False)
        # multi-GPU: bool mask of the nodes whose rows this rank assembles (fedoo_b200/dist.py)
        self.owned_nodes = kargs.pop("owned_nodes", None)
        # reuse the K / D device buffers (and a pinned host buffer for D) across assemblies
        # instead of replacing them (the reference replaces; useful for 15 GB matrices)
        self.reuse_buffers = kargs.pop("reuse_buffers", False)
        # leave the global vector in HBM (``global_vector`` is then the CUDA tensor): no D2H copy
        self.vector_on_device = kargs.pop("vector_on_device", False)
        # pipeline the host copies (pinned dof vector in, residual out) on two copy streams with rotating buffers, so that
        # step i + 1's H2D and step i's D2H overlap the kernels; ``global_vector`` is then a ``HostVector`` future
        self.async_copies = kargs.pop("async_copies", False)
        self._pipe = None
        # multi-GPU: a fedoo_b200.dist.PeerVector -> the residual exchange is fused into the assembly kernel
        self.peer_vector = kargs.pop("peer_vector", None)
        self._bufs = {}
        self.mesh = mesh
        if elm_type == "":
            elm_type = mesh.elm_type
        self.elm_type = elm_type.lower()
        if self.elm_type not in _DEFAULT_NGP:
            raise NotImplementedError(
                f"element type '{self.elm_type}' is not on the accelerated path (hex8, tet4, tet10, quad4)"
            )
        self.n_elm_gp = kargs.pop("n_elm_gp", None) or _DEFAULT_NGP[self.elm_type]
        if self.n_elm_gp != _DEFAULT_NGP[self.elm_type]:
            raise NotImplementedError(f"{self.elm_type} is integrated with {_DEFAULT_NGP[self.elm_type]} Gauss points")
        if mesh.elements.shape[1] != {"hex8": 8, "tet4": 4, "tet10": 10, "quad4": 4}[self.elm_type]:
            raise ValueError("mesh connectivity does not match the element type")
        if mesh.ndim != (2 if self.elm_type == "quad4" else 3) or self.space.ndim != mesh.ndim:
            raise ValueError("mesh / modeling space dimension mismatch")
        self.assume_sym = weakform.assembly_options.get("assume_sym", False)
        self.sv, self.sv_start, self.sv_type, self.sv_component = {}, {}, {}, {}
        self._nlgeom = None
        self._pb = None
        self.global_matrix = None
        self.global_vector = None
        self.global_vector_device = None
        self._saved_bloc_structure = None
        self._U_dev = None
        self._T_start_dev = None
        self._register(name)

    # ------------------------------------------------------------------ sizes
    @property
    def n_gauss_points(self):
        return self.mesh.n_elements * self.n_elm_gp

    @property
    def nvar(self):
        return 1 if self.weakform.operator == "heat" else self.space.ndim

    # ------------------------------------------------------------------ symbolic (one-time)
    def _symbolic(self):
        """Pattern, tiled CSR and cluster plan; cached per (mesh, element type) and per nvar."""
        # the entry keeps the mesh and the mask alive, so their ids cannot be recycled while it is cached (the
        # reference keys its caches on the mesh object itself); sizes are checked as well: a mesh edited in place
        # (new connectivity) must not reuse a stale pattern
        key = (id(self.mesh), self.elm_type, id(self.owned_nodes))
        entry = Assembly._saved_plans.get(key)
        if entry is not None and not (
            entry["mesh"] is self.mesh and entry["owned"] is self.owned_nodes
            and entry["sizes"] == (self.mesh.n_nodes, self.mesh.n_elements, id(self.mesh.elements))
        ):  # fmt: skip
            entry = None
        if entry is None:
            coords, conn = self.mesh.device_arrays()
            pattern = symbolic.build_pattern(conn, self.mesh.n_nodes)
            # the cluster plan is built on first use (``_plan``): the row-owner kernels do not need one
            entry = {"pattern": pattern, "plan": None, "csr": {}, "mesh": self.mesh, "owned": self.owned_nodes,
                     "sizes": (self.mesh.n_nodes, self.mesh.n_elements, id(self.mesh.elements))}  # fmt: skip
            Assembly._saved_plans[key] = entry
        nvar = self.nvar
        n_glob = 0 if self._pb is None else getattr(self._pb, "n_global_dof", 0)
        if (nvar, n_glob) not in entry["csr"]:
            entry["csr"][(nvar, n_glob)] = symbolic.expand_csr(entry["pattern"], nvar, n_glob)
        self._saved_bloc_structure = entry
        return entry, entry["csr"][(nvar, n_glob)], n_glob

    def _plan(self, entry):
        """Cluster plan of the pattern (fedoo_b200/plan.py), built on first use."""
        if entry["plan"] is None:
            coords, conn = self.mesh.device_arrays()
            owned = None if self.owned_nodes is None else torch.from_numpy(np.asarray(self.owned_nodes, dtype=bool))
            entry["plan"] = symbolic.build_plan(self.elm_type, coords, conn, entry["pattern"], owned=owned)
        return entry["plan"]

    def _row_positions(self, entry, node_ptr, node_inc):
        """One-time table of the row-owner kernels: for every incidence (node I, element e) the record {element * nne +
        local node, position of each column conn[e][j] inside row I of the block pattern packed 4 x u8} -- the row-local
        counterpart of the reference's ``Matrix_convertCOOtoCSR`` (fedoo/core/_sparsematrix.py:256-274)."""
        if "row_pos" not in entry:
            pattern = entry["pattern"]
            conn = self.mesh.device_arrays()[1].to(torch.int64)
            nne = conn.shape[1]
            assert nne == 4, "packed positions: four columns per incidence"
            n = self.mesh.n_nodes
            elem = node_inc.to(torch.int64) // nne
            node = torch.repeat_interleave(torch.arange(n, device=conn.device), (node_ptr[1:] - node_ptr[:-1]))
            keys = node[:, None] * n + conn[elem]  # (n_inc, 4) block keys I * n + J
            slot = torch.searchsorted(pattern.keys, keys.reshape(-1)).reshape(-1, nne)
            pos = slot - pattern.blk_indptr[node][:, None]
            deg = pattern.blk_indptr[1:] - pattern.blk_indptr[:-1]
            max_deg = int(deg.max()) if n else 0
            if max_deg > 255:
                raise _lib.FdkError(f"row degree {max_deg} exceeds the packed position format")
            packed = pos[:, 0] | (pos[:, 1] << 8) | (pos[:, 2] << 16) | (pos[:, 3] << 24)
            packed = torch.where(packed >= 2**31, packed - 2**32, packed).to(torch.int32)  # the same 32 bits, signed
            # one 8-byte record per incidence: {element * nne + local node, packed positions}
            entry["row_pos"] = (torch.stack([node_inc.to(torch.int32), packed], dim=1).contiguous(), max_deg)
        return entry["row_pos"]

    def _owned_rows(self, entry):
        """int32 list of the owned node rows on the device (None: every node is owned)."""
        if self.owned_nodes is None:
            return None
        if "owned_rows" not in entry:
            own = np.flatnonzero(np.asarray(self.owned_nodes, dtype=bool)).astype(np.int32)
            entry["owned_rows"] = torch.from_numpy(own).to(device())
        return entry["owned_rows"]

    def _heavy_rows(self, entry, plan, flags, coords, iso, lam, mu, C_h, tangent_dev, U_dev, stress_dev, K, D):
        """Rows of the nodes the cluster plan left out (more incident elements than one cluster holds): csrc/fdk_rows.cuh."""
        rows = plan.heavy_nodes
        if rows.numel() == 0 or not flags:
            return
        from .results import node_incidences

        pattern = entry["pattern"]
        node_ptr, node_inc = node_incidences(self.mesh)
        conn = self.mesh.device_arrays()[1]
        if "max_deg" not in entry:
            entry["max_deg"] = int((pattern.blk_indptr[1:] - pattern.blk_indptr[:-1]).max())
        rc = _lib.load().fdk_assemble_rows_elastic(
            _lib.ELEM_IDS[self.elm_type], int(rows.numel()), _lib.ptr(rows), self.mesh.n_nodes, self.mesh.n_elements,
            _lib.ptr(conn), _lib.ptr(coords), _lib.ptr(node_ptr), _lib.ptr(node_inc), _lib.ptr(pattern.blk_indptr),
            _lib.ptr(pattern.blk_indices), pattern.blk_nnz, entry["max_deg"], int(iso), float(lam), float(mu), _lib.ptr(C_h),
            _lib.ptr(tangent_dev), flags, _lib.ptr(U_dev), _lib.ptr(stress_dev), _lib.ptr(K), _lib.ptr(D),
            _lib.current_stream(),
        )  # fmt: skip
        _lib.check(rc, "fdk_assemble_rows_elastic")

    def _big_plan(self, entry):
        """Plan with clusters for 1024-thread CTAs (tet10 + isotropic law: csrc/fdk_assemble_iso.cuh with 5 threads per
        incidence), built on first use."""
        if "plan_big" not in entry:
            coords, conn = self.mesh.device_arrays()
            owned = None if self.owned_nodes is None else torch.from_numpy(np.asarray(self.owned_nodes, dtype=bool))
            entry["plan_big"] = symbolic.build_plan(self.elm_type, coords, conn, entry["pattern"], owned=owned, big=True)
        return entry["plan_big"]

    def _small_plan(self, entry):
        """Second cluster plan of the same pattern with half-size (16-node hex8) clusters, built on first use."""
        if "plan_small" not in entry:
            coords, conn = self.mesh.device_arrays()
            owned = None if self.owned_nodes is None else torch.from_numpy(np.asarray(self.owned_nodes, dtype=bool))
            entry["plan_small"] = symbolic.build_plan(self.elm_type, coords, conn, entry["pattern"], owned=owned, small=True)
        return entry["plan_small"]

    def _coords(self):
        if getattr(self, "_coords_override", None) is not None:
            return self._coords_override
        if self.meshChange:
            self.mesh.invalidate_device()
        return self.mesh.device_arrays()[0]

    def set_disp(self, disp):
        """Updated-Lagrangian geometry refresh (fedoo/core/assembly.py:1207-1229): ``self.current`` becomes an assembly
        on the node positions ``mesh.nodes + disp.T``.  The reference has to rebuild the Jacobians and the elementary
        operators of the moved mesh; here J, det J and grad N are recomputed inside the assembly kernel from the
        coordinates at every launch, so the refresh is one axpy on the coordinate array -- pattern, cluster plan and
        CSR structure (topology only) are shared with the undeformed assembly."""
        if np.isscalar(disp) and disp == 0:
            self.current = self
            return
        from copy import copy

        from .core import as_device_f64

        base = self.mesh.device_arrays()[0]
        d = as_device_f64(disp, base.device).reshape(-1, self.mesh.n_nodes)[: base.shape[1]]
        if self.current is self:
            cur = copy(self)
            cur.global_matrix = cur.global_vector = cur.global_vector_device = None
            cur._bufs = {}
            cur.current = cur
            self.current = cur
        self.current._coords_override = (base + d.T).contiguous()

    # ------------------------------------------------------------------ the hot path
    def assemble_global_mat(self, compute="all"):
        """fedoo/core/assembly.py:143-470."""
        if compute == "none":
            return
        if compute not in ("all", "matrix", "vector"):
            raise ValueError("compute must be 'all', 'matrix', 'vector' or 'none'")
        lib = _lib.load()
        entry, (indptr, indices), n_glob = self._symbolic()
        pattern = entry["pattern"]
        nvar = self.nvar
        n_nodes = self.mesh.n_nodes
        dev = device()
        coords = self._coords()
        want_mat = compute != "vector"
        want_vec = compute != "matrix"
        stream = _lib.current_stream()

        if self.weakform.operator == "elastic":
            law = self.weakform.constitutivelaw
            dimension = self.space.get_dimension()
            stress = self.sv.get("Stress", 0)
            has_vec = want_vec and not (np.isscalar(stress) and stress == 0)
            flags = (_lib.MATRIX if want_mat else 0) | (_lib.VECTOR if has_vec else 0)
            if self.mesh.n_elements == 0:
                flags = 0  # nothing to integrate: K has no entry, D stays zero
            K = self._buffer("K", nvar * nvar * pattern.blk_nnz, zero=self.owned_nodes is not None) if want_mat else None
            D = self._buffer(self._d_tag(), nvar * n_nodes + n_glob, zero=True) if has_vec else None
            U_dev = stress_dev = None
            if has_vec:
                if isinstance(stress, _FusedElasticStress):
                    U_dev = stress.U  # sigma = H eps(U) recomputed in the kernel, never materialised
                else:
                    stress_dev = stress.device_tensor
            if flags:
                if getattr(law, "tangent_r1_device", None) is not None and law.tangent_r1_device(self) is not None:
                    tangent_dev = _LAZY
                else:
                    tangent_dev = law.tangent_device(self) if hasattr(law, "tangent_device") else None
                peer = self.peer_vector
                fusable = isinstance(law, ElasticIsotrop) and tangent_dev is None and stress_dev is None
                split = (flags == _lib.ALL and not fusable and self.owned_nodes is None and peer is None
                         and _RESIDUAL_KERNEL)  # fmt: skip
                if split:
                    # K and D both wanted but the residual is not -K U (plastic stress, F-bar, general tangent):
                    # matrix through the cluster kernel without its B^T sigma pass, residual through its own kernels
                    flags = _lib.MATRIX
                done_vec = False
                if (flags == _lib.VECTOR or split) and self.owned_nodes is None and peer is None and _RESIDUAL_KERNEL:
                    done_vec = True
                    # residual alone (every Newton sub-iteration): element forces + per-node gather, no cluster plan
                    from .results import node_incidences

                    node_ptr, node_inc = node_incidences(self.mesh)
                    conn = self.mesh.device_arrays()[1]
                    fe = self._scratch("fe", self.mesh.n_elements * conn.shape[1] * self.space.ndim)
                    C_h = None
                    if stress_dev is None and tangent_dev is _LAZY:
                        tangent_dev = law.tangent_device(self)
                    if stress_dev is None and tangent_dev is None:
                        C_h = np.ascontiguousarray(self.sv["TangentMatrix"], dtype=np.float64)
                    rc = lib.fdk_residual_elastic(
                        _lib.ELEM_IDS[self.elm_type], n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                        _lib.ptr(C_h), _lib.ptr(tangent_dev if stress_dev is None else None), _lib.ptr(U_dev),
                        _lib.ptr(stress_dev), _lib.ptr(node_ptr), _lib.ptr(node_inc), _lib.ptr(fe), _lib.ptr(D), stream,
                    )  # fmt: skip
                    _lib.check(rc, "fdk_residual_elastic")
                if flags == _lib.VECTOR and done_vec:
                    pass
                elif (isinstance(law, ElasticIsotrop) and tangent_dev is None and peer is not None and flags == _lib.ALL
                        and U_dev is not None and stress_dev is None and getattr(peer, "mode", "fused") == "fused"):
                    # multi-GPU: the kernel stores the owned residual entries straight into every rank's global vector
                    lam, mu = law.lame(dimension)
                    dst, n_dst = peer.begin_step()  # the half of the double-buffered symmetric vector this step writes
                    rc = lib.fdk_assemble_elastic_iso_dist(
                        C.byref(self._plan(entry).struct(nvar)), flags, _lib.ptr(coords), lam, mu, _lib.ptr(U_dev), _lib.ptr(K),
                        _lib.ptr(D), dst, n_dst, _lib.ptr(peer.node_gid), peer.n_global, stream,
                    )  # fmt: skip
                    _lib.check(rc, "fdk_assemble_elastic_iso_dist")
                    if self._plan(entry).heavy_nodes.numel():
                        raise NotImplementedError("high-valence nodes with the fused multi-GPU exchange")
                    peer.barrier()
                elif isinstance(law, ElasticIsotrop) and tangent_dev is None:
                    lam, mu = law.lame(dimension)
                    plan = self._plan(entry)
                    if (self.elm_type == "tet10" and _TET10_BIG and want_mat and (U_dev is not None or not has_vec)
                            and stress_dev is None):
                        plan = self._big_plan(entry)  # 1024-thread clusters for the balanced kernel
                    rc = lib.fdk_assemble_elastic_iso(
                        C.byref(plan.struct(nvar)), flags, _lib.ptr(coords), lam, mu, _lib.ptr(U_dev),
                        _lib.ptr(stress_dev), _lib.ptr(K), _lib.ptr(D), stream,
                    )  # fmt: skip
                    _lib.check(rc, "fdk_assemble_elastic_iso")
                    self._heavy_rows(entry, plan, flags, coords, True, lam, mu, None, None, U_dev, stress_dev, K, D)
                    if peer is not None and has_vec:  # multi-GPU, "copy" mode: one coalesced copy of the owned slices
                        peer.publish(D)
                        peer.barrier()
                elif (getattr(law, "tangent_r1_device", None) is not None and law.tangent_r1_device(self) is not None
                        and self.elm_type == "hex8" and nvar == 3 and want_mat):
                    # J2 tangent in its structured form: the balanced kernel reads 10 doubles per (element, Gauss point)
                    plan = self._small_plan(entry)
                    rc = lib.fdk_assemble_elastic_r1(
                        C.byref(plan.struct(nvar)), flags, _lib.ptr(coords), _lib.ptr(law.tangent_r1_device(self)),
                        _lib.ptr(stress_dev), _lib.ptr(K), _lib.ptr(D), stream,
                    )  # fmt: skip
                    _lib.check(rc, "fdk_assemble_elastic_r1")
                    if plan.heavy_nodes.numel():
                        self._heavy_rows(entry, plan, flags, coords, False, 0.0, 0.0, None, law.tangent_device(self), U_dev,
                                         stress_dev, K, D)  # fmt: skip
                else:
                    if tangent_dev is _LAZY:
                        tangent_dev = law.tangent_device(self)
                    H = self.sv["TangentMatrix"]
                    C_h = None if tangent_dev is not None else np.ascontiguousarray(H, dtype=np.float64)
                    if self.elm_type == "hex8" and nvar == 3 and want_mat:
                        # general tangent on hex8: 16-node clusters, so that the 36 tangent entries of every (touched
                        # element, Gauss point) fit in shared memory next to the geometry (csrc/fdk_assemble_iso.cuh)
                        plan = self._small_plan(entry)
                    else:
                        plan = self._plan(entry)
                    rc = lib.fdk_assemble_elastic_general(
                        C.byref(plan.struct(nvar)), flags, _lib.ptr(coords), _lib.ptr(C_h), _lib.ptr(tangent_dev),
                        _lib.ptr(U_dev), _lib.ptr(stress_dev), _lib.ptr(K), _lib.ptr(D), stream,
                    )  # fmt: skip
                    _lib.check(rc, "fdk_assemble_elastic_general")
                    self._heavy_rows(entry, plan, flags, coords, False, 0.0, 0.0, C_h, tangent_dev, U_dev, stress_dev, K, D)
        elif self.weakform.operator == "heat":
            law = self.weakform.constitutivelaw
            cond = np.ascontiguousarray(np.asarray(law.thermal_conductivity, dtype=np.float64).reshape(3, 3))
            dtime = getattr(self._pb, "dtime", 0) if self._pb is not None else 0
            rcdt = float(law.density * law.specific_heat / dtime) if (self.weakform.transient and dtime != 0) else 0.0
            T_dev = self._U_dev
            has_vec = want_vec and T_dev is not None
            flags = (_lib.MATRIX if want_mat else 0) | (_lib.VECTOR if has_vec else 0)
            if self.mesh.n_elements == 0:
                flags = 0
            K = self._buffer("K", pattern.blk_nnz, zero=self.owned_nodes is not None) if want_mat else None
            D = self._buffer(self._d_tag(), n_nodes + n_glob, zero=True) if has_vec else None
            T_start = self._T_start_dev if rcdt != 0.0 else None
            if flags and self.elm_type == "tet4" and _HEAT_TET4_ROWS:
                # constant-gradient element: the row-owner kernel does K and D in one launch, without a cluster plan;
                # on a rank of a multi-GPU partition it computes the rows of the owned nodes only
                from .results import node_incidences

                node_ptr, node_inc = node_incidences(self.mesh)
                inc_rec, max_deg = self._row_positions(entry, node_ptr, node_inc)
                conn = self.mesh.device_arrays()[1]
                rows = self._owned_rows(entry)
                rc = lib.fdk_assemble_heat_tet4(
                    flags, n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords), _lib.ptr(cond), rcdt,
                    _lib.ptr(T_dev), _lib.ptr(T_start), _lib.ptr(node_ptr), _lib.ptr(inc_rec),
                    _lib.ptr(pattern.blk_indptr), max_deg, n_nodes if rows is None else int(rows.numel()), _lib.ptr(rows),
                    _lib.ptr(K), _lib.ptr(D), stream,
                )  # fmt: skip
                _lib.check(rc, "fdk_assemble_heat_tet4")
                flags = 0
            split = (flags & _lib.VECTOR) and self.owned_nodes is None and _RESIDUAL_KERNEL
            if split:
                # the residual through its own kernels; with K also wanted the cluster kernel then runs matrix-only
                from .results import node_incidences

                node_ptr, node_inc = node_incidences(self.mesh)
                conn = self.mesh.device_arrays()[1]
                fe = self._scratch("fe", self.mesh.n_elements * conn.shape[1])
                rc = lib.fdk_residual_heat(
                    _lib.ELEM_IDS[self.elm_type], n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                    _lib.ptr(cond), rcdt, _lib.ptr(T_dev), _lib.ptr(T_start),
                    _lib.ptr(node_ptr), _lib.ptr(node_inc), _lib.ptr(fe), _lib.ptr(D), stream,
                )  # fmt: skip
                _lib.check(rc, "fdk_residual_heat")
                flags &= ~_lib.VECTOR
            if flags:
                plan = self._plan(entry)
                if plan.heavy_nodes.numel():
                    raise NotImplementedError(
                        f"{int(plan.heavy_nodes.numel())} nodes of this mesh touch more elements than a cluster of the heat "
                        "kernel holds (the row-owner kernel covers tet4 meshes of any valence)"
                    )
                rc = lib.fdk_assemble_heat(
                    C.byref(plan.struct(1)), flags, _lib.ptr(coords), _lib.ptr(cond), rcdt, _lib.ptr(T_dev),
                    _lib.ptr(T_start), _lib.ptr(K), _lib.ptr(D), stream,
                )  # fmt: skip
                _lib.check(rc, "fdk_assemble_heat")
        else:
            raise NotImplementedError(f"weak form operator '{self.weakform.operator}'")

        if self._pipe is not None:
            self._pipe.kernel_finished()
        if want_mat:
            n_rows = nvar * n_nodes + n_glob
            block = (pattern.blk_indptr, pattern.blk_indices, nvar, n_nodes)
            self.global_matrix = DeviceCSR(indptr, indices, K, (n_rows, n_rows), block=block, n_glob=n_glob)
        if want_vec:
            if has_vec:
                self.global_vector_device = D
                if self.vector_on_device:
                    self.global_vector = D
                elif self.async_copies and self.reuse_buffers:
                    self.global_vector = self._copy_pipe().d2h(D)
                else:
                    self.global_vector = self._to_host(D)
            else:
                self.global_vector_device = None
                self.global_vector = 0

    def _copy_pipe(self):
        if self._pipe is None:
            self._pipe = _CopyPipe()
        return self._pipe

    def _d_tag(self):
        """Output buffer of the residual: two of them in rotation when the D2H copies are pipelined (the kernel of step
        i + 1 must not write the buffer step i's copy is still reading)."""
        if self.async_copies and self.reuse_buffers and not self.vector_on_device:
            return "D%d" % self._copy_pipe().begin_output()
        return "D"

    def _scratch(self, tag, n):
        """Device scratch kept on the assembly between calls (never handed out)."""
        b = self._bufs.get(tag)
        if b is None or b.numel() != n:
            b = torch.empty(n, dtype=torch.float64, device=device())
            self._bufs[tag] = b
        return b

    def _buffer(self, tag, n, zero=False):
        """Device output buffer.  Entries the kernels do not write (global dofs, halo nodes of a
        rank-local mesh) must read 0, hence ``zero`` on (first) allocation."""
        dev = device()
        if not self.reuse_buffers:
            return (torch.zeros if zero else torch.empty)(n, dtype=torch.float64, device=dev)
        b = self._bufs.get(tag)
        if b is None or b.numel() != n:
            b = (torch.zeros if zero else torch.empty)(n, dtype=torch.float64, device=dev)
            self._bufs[tag] = b
        return b

    def _to_host(self, D):
        if not self.reuse_buffers:
            return D.cpu().numpy()
        h = self._bufs.get("D_host")
        if h is None or h.numel() != D.numel():
            h = torch.empty(D.numel(), dtype=torch.float64, pin_memory=True)
            self._bufs["D_host"] = h
        h.copy_(D, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return h.numpy()

    def get_global_matrix(self):
        if self.global_matrix is None:
            self.assemble_global_mat()
        return self.global_matrix

    def get_global_vector(self):
        if self.global_vector is None:
            self.assemble_global_mat()
        return self.global_vector

    def delete_global_mat(self):
        self.global_matrix = self.global_vector = self.global_vector_device = None

    # ------------------------------------------------------------------ lifecycle (assembly.py:672-753)
    def initialize(self, pb):
        self._pb = pb
        self.weakform.initialize(self, pb)
        if self.weakform.constitutivelaw is not None:
            self.weakform.constitutivelaw.initialize(self, pb)
        self.sv_start = dict(self.sv)

    def set_start(self, pb):
        self.weakform.set_start(self, pb)
        if self.weakform.constitutivelaw is not None:
            self.weakform.constitutivelaw.set_start(self, pb)
        self.sv_start = dict(self.sv)
        self._assemble_current("all")

    def _assemble_current(self, compute):
        """fedoo/core/assembly.py:706,721,735: the lifecycle assembles ``self.current`` (the assembly on the deformed
        configuration once ``set_disp`` has been called), which reads the state of this assembly."""
        cur = self.current
        if cur is not self:
            cur.sv, cur.sv_start, cur._U_dev, cur._T_start_dev, cur._pb = (
                self.sv, self.sv_start, self._U_dev, self._T_start_dev, self._pb)  # fmt: skip
        cur.assemble_global_mat(compute)

    def update(self, pb, compute="all"):
        self.weakform.update(self, pb)
        if self.weakform.constitutivelaw is not None:
            self.weakform.constitutivelaw.update(self, pb)
        self.weakform.update_2(self, pb)
        self._assemble_current(compute)

    def to_start(self, pb):
        self.weakform.to_start(self, pb)
        if self.weakform.constitutivelaw is not None:
            self.weakform.constitutivelaw.to_start(self, pb)
        self.sv = dict(self.sv_start)
        self._assemble_current("all")

    def reset(self):
        self.weakform.reset()
        if self.weakform.constitutivelaw is not None:
            self.weakform.constitutivelaw.reset()
        self.delete_global_mat()
        self.sv, self.sv_start, self.sv_type = {}, {}, {}

    # ------------------------------------------------------------------ state update helpers
    def _strain_update(self, U):
        """StressEquilibrium.update (fedoo/weakform/stress_equilibrium.py:191-217): record the dof
        vector; sv['Strain'] / sv['DispGradient'] are produced lazily by ``fdk_gp_strain_stress``."""
        if self.async_copies and isinstance(U, torch.Tensor) and not U.is_cuda and U.is_pinned():
            self._U_dev = self._copy_pipe().h2d(U)
        else:
            self._U_dev = as_device_f64(U)
        fbar = bool(getattr(self.weakform, "fbar", False))
        if fbar and self.space.ndim != 3:
            raise NotImplementedError("F-bar is available for the 3D modeling space")
        self.sv["Strain"] = _LazyStrain(self, self._U_dev, fbar)
        self.sv["DispGradient"] = _LazyGrad(self, self._U_dev, fbar)

    def _elastic_stress_update(self, law):
        """ElasticAnisotropic.update (fedoo/constitutivelaw/elastic_anisotropic.py:36-56)."""
        strain = self.sv.get("Strain", 0)
        if np.isscalar(strain) and strain == 0:
            self.sv["Stress"] = 0
            return
        if getattr(strain, "fbar", False):
            # F-bar: sigma = H eps_bar is not -K U any more -> materialise it, the kernel integrates B^T sigma
            stress = self._gp_strain_stress(strain.U, want_stress=True, law=law, fbar=True)[2]
            self.sv["Stress"] = GaussPointTensor(stress, "stress")
            return
        self.sv["Stress"] = _FusedElasticStress(self, law, strain.U)

    def _gp_strain_stress(self, U_dev, want_grad=False, want_strain=False, want_stress=False, law=None, fbar=False):
        lib = _lib.load()
        coords, conn = self._coords(), self.mesh.device_arrays()[1]
        N = self.n_gauss_points
        dev = device()
        grad = torch.empty((9, N), dtype=torch.float64, device=dev) if want_grad else None
        strain = torch.empty((N, 6), dtype=torch.float64, device=dev) if want_strain else None
        stress = torch.empty((N, 6), dtype=torch.float64, device=dev) if want_stress else None
        C_h = tangent_dev = None
        if want_stress:
            tangent_dev = law.tangent_device(self)
            if tangent_dev is None:
                C_h = np.ascontiguousarray(self.sv["TangentMatrix"], dtype=np.float64)
        if fbar:  # small-strain F-bar (fedoo/weakform/stress_equilibrium.py:527-540)
            center = torch.empty(self.mesh.n_elements, dtype=torch.float64, device=dev)
            _lib.check(
                lib.fdk_gp_strain_stress_fbar(
                    _lib.ELEM_IDS[self.elm_type], self.mesh.n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                    _lib.ptr(U_dev), _lib.ptr(C_h), _lib.ptr(tangent_dev), _lib.ptr(center), _lib.ptr(grad), _lib.ptr(strain),
                    _lib.ptr(stress), _lib.current_stream(),
                ),
                "fdk_gp_strain_stress_fbar",
            )  # fmt: skip
            return grad, strain, stress
        _lib.check(
            lib.fdk_gp_strain_stress(
                _lib.ELEM_IDS[self.elm_type], self.mesh.n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                _lib.ptr(U_dev), _lib.ptr(C_h), _lib.ptr(tangent_dev), _lib.ptr(grad), _lib.ptr(strain),
                _lib.ptr(stress), _lib.current_stream(),
            ),
            "fdk_gp_strain_stress",
        )  # fmt: skip
        return grad, strain, stress

    def convert_data(self, data, convert_from="GaussPoint", convert_to="Node"):
        """Mesh.convert_data for Gauss-point fields (fedoo/core/mesh.py:1267-1308), on the device; returns NumPy."""
        if convert_from != "GaussPoint":
            raise NotImplementedError("only Gauss-point fields are converted on the accelerated path")
        from .results import convert_gp

        return convert_gp(self, data, convert_to).cpu().numpy()

    def get_strain(self, U, type_output="Node", nlgeom=False):
        """Legacy accessor used by the reference's cantilever test (fedoo/core/assembly.py get_strain): the small
        strain of the dof vector U at the nodes / elements / Gauss points, six arrays in Voigt order."""
        if nlgeom:
            raise NotImplementedError("nlgeom is outside the accelerated path")
        from .results import NodeTensor, convert_gp

        strain = self._gp_strain_stress(as_device_f64(U), want_strain=True)[1]  # (N, 6)
        return NodeTensor(convert_gp(self, GaussPointTensor(strain, "strain"), type_output).cpu().numpy())

    def get_grad_disp(self, U, type_output="GaussPoint"):
        """fedoo/core/assembly.py:1285-1336: 3x3 list of (N,) arrays, gp-major."""
        if type_output != "GaussPoint":
            raise NotImplementedError("only Gauss-point output is on the accelerated path")
        grad, _, _ = self._gp_strain_stress(as_device_f64(U), want_grad=True)
        g = grad.cpu().numpy()
        ndim = self.space.ndim
        return [[g[a * 3 + b] if (a < ndim and b < ndim) else 0 for b in range(3)] for a in range(3)]

    def get_deformation_gradient(self, U, fbar=None):
        """F = 1 + grad u at the Gauss points as a device tensor (N, 3, 3)[n, j, i] -- the memory of the reference's
        Fortran-ordered ``sv["F"]`` of shape (3, 3, N) -- or, with ``fbar`` (default: the weak form's ``fbar`` flag), the
        F-bar form F (J_mean / J)^(1/3) (fedoo/weakform/stress_equilibrium.py:542-586, ``_comp_F`` / ``_comp_Fbar``).
        ``result.permute(2, 1, 0)`` is the (3, 3, N) view.  These are the finite-strain kinematics the reference computes
        itself; the strain measures and objective rates that follow live in simcoon and are not on this path."""
        if fbar is None:
            fbar = bool(getattr(self.weakform, "fbar", False))
        coords, conn = self._coords(), self.mesh.device_arrays()[1]
        U_dev = as_device_f64(U)
        n_nodes = self.mesh.n_nodes
        if U_dev.numel() < self.space.ndim * n_nodes:
            raise ValueError("the dof vector is shorter than ndim * n_nodes")
        F = torch.empty((self.n_gauss_points, 3, 3), dtype=torch.float64, device=device())
        _lib.check(
            _lib.load().fdk_gp_deformation_gradient(
                _lib.ELEM_IDS[self.elm_type], n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                _lib.ptr(U_dev), 1 if fbar else 0, _lib.ptr(F), _lib.current_stream(),
            ),
            "fdk_gp_deformation_gradient",
        )  # fmt: skip
        return F

    def _thermal_state_update(self, pb, initialize=False):
        """SteadyHeatEquation.update / TemperatureTimeDerivative.update
        (fedoo/weakform/heat_equation.py:54-70,140-152)."""
        T = pb.get_dof_solution()
        if np.isscalar(T):
            self._U_dev = None
            self.sv["TempGradient"] = [0, 0, 0]
            self.sv["Temp"] = 0
            if initialize:
                self._T_start_dev = None
            return
        if self.async_copies and isinstance(T, torch.Tensor) and not T.is_cuda and T.is_pinned():
            self._U_dev = self._copy_pipe().h2d(T)
        else:
            self._U_dev = as_device_f64(T)
        if initialize:
            self._T_start_dev = self._U_dev.clone()
        self.sv["Temp"] = _LazyTemp(self, self._U_dev, 0)
        self.sv["TempGradient"] = _LazyTemp(self, self._U_dev, 1)

    def _thermal_set_start(self, pb):
        """TemperatureTimeDerivative.set_start (fedoo/weakform/heat_equation.py:164-165)."""
        self._T_start_dev = None if self._U_dev is None else self._U_dev.clone()

    def _gp_temperature(self, T_dev):
        lib = _lib.load()
        coords, conn = self._coords(), self.mesh.device_arrays()[1]
        N = self.n_gauss_points
        dev = device()
        temp = torch.empty(N, dtype=torch.float64, device=dev)
        grad = torch.empty((3, N), dtype=torch.float64, device=dev)
        _lib.check(
            lib.fdk_gp_temperature(
                _lib.ELEM_IDS[self.elm_type], self.mesh.n_nodes, self.mesh.n_elements, _lib.ptr(conn), _lib.ptr(coords),
                _lib.ptr(T_dev), _lib.ptr(temp), _lib.ptr(grad), _lib.current_stream(),
            ),
            "fdk_gp_temperature",
        )  # fmt: skip
        return temp, grad


class HostVector:
    """The residual on its way to (pinned) host memory: ``result()`` / ``np.asarray()`` wait for the copy.  The
    storage rotates: a result must be consumed before the assembly after next overwrites it."""

    def __init__(self, host, event):
        self._host, self._event = host, event

    def result(self):
        if self._event is not None:
            self._event.synchronize()
            self._event = None
        return self._host.numpy()

    def __array__(self, dtype=None, copy=None):
        a = self.result()
        return a if dtype is None else a.astype(dtype)

    def __len__(self):
        return self._host.numel()

    @property
    def size(self):
        return self._host.numel()


class _CopyPipe:
    """Two copy streams (H2D, D2H: PCIe is full duplex) and two-deep rotating buffers around the assembly kernel.
    Ordering, all by events: H2D(i) waits for the kernel that last read its device buffer (i - 2); kernel(i) waits for
    H2D(i) and for the D2H that last read its output buffer (i - 2); D2H(i) waits for kernel(i)."""

    def __init__(self):
        self.h2d_stream, self.d2h_stream = torch.cuda.Stream(), torch.cuda.Stream()
        self.k_in = self.k_out = 0
        self.U = [None, None]
        self.kernel_done = [None, None]  # per input buffer: the assembly that read it has finished
        self.host = [None, None]
        self.d2h_done = [None, None]  # per output buffer

    def h2d(self, U_host):
        k = self.k_in = self.k_in ^ 1
        if self.U[k] is None or self.U[k].numel() != U_host.numel():
            self.U[k] = torch.empty(U_host.numel(), dtype=torch.float64, device=device())
        if self.kernel_done[k] is not None:
            self.h2d_stream.wait_event(self.kernel_done[k])
        with torch.cuda.stream(self.h2d_stream):
            self.U[k].copy_(U_host.reshape(-1), non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.h2d_stream)
        torch.cuda.current_stream().wait_event(ev)
        return self.U[k]

    def begin_output(self):
        k = self.k_out = self.k_out ^ 1
        if self.d2h_done[k] is not None:
            torch.cuda.current_stream().wait_event(self.d2h_done[k])
        return k

    def kernel_finished(self):
        """Called once per assembly, after its last launch on the current stream."""
        done = torch.cuda.Event()
        done.record(torch.cuda.current_stream())
        self.kernel_done[self.k_in] = self.last_done = done

    def d2h(self, D):
        k = self.k_out
        done = self.last_done
        if self.host[k] is None or self.host[k].numel() != D.numel():
            self.host[k] = torch.empty(D.numel(), dtype=torch.float64, pin_memory=True)
        self.d2h_stream.wait_event(done)
        with torch.cuda.stream(self.d2h_stream):
            self.host[k].copy_(D, non_blocking=True)
            ev = torch.cuda.Event()
            ev.record(self.d2h_stream)
        self.d2h_done[k] = ev
        return HostVector(self.host[k], ev)


class _LazyStrain(GaussPointTensor):
    """sv['Strain'] of the current dof vector, computed on first access."""

    def __init__(self, assembly, U, fbar=False):
        self._asm, self.U, self.kind, self._host, self._dev, self.fbar = assembly, U, "strain", None, None, fbar

    @property
    def device_tensor(self):
        if self._dev is None:
            self._dev = self._asm._gp_strain_stress(self.U, want_strain=True, fbar=self.fbar)[1]
        return self._dev


class _FusedElasticStress(GaussPointTensor):
    """sv['Stress'] = H eps(U) of a linear elastic law.  The residual kernel recomputes it on the
    fly from U (no (6,N) array is written); it is materialised only if somebody reads it."""

    def __init__(self, assembly, law, U):
        self._asm, self._law, self.U, self.kind, self._host, self._dev = assembly, law, U, "stress", None, None

    @property
    def device_tensor(self):
        if self._dev is None:
            self._dev = self._asm._gp_strain_stress(self.U, want_stress=True, law=self._law)[2]
        return self._dev


class _LazyGrad:
    def __init__(self, assembly, U, fbar=False):
        self._asm, self.U, self._host, self.fbar = assembly, U, None, fbar

    def _get(self):
        if self._host is None:
            self._host = self._asm._gp_strain_stress(self.U, want_grad=True, fbar=self.fbar)[0].cpu().numpy()
        return self._host

    def __getitem__(self, a):
        g = self._get()
        return [g[a * 3 + b] for b in range(3)]


class _LazyTemp:
    """sv['Temp'] (which=0, (N,)) or sv['TempGradient'] (which=1, list of 3 (N,))."""

    def __init__(self, assembly, T, which):
        self._asm, self.T, self.which, self._host = assembly, T, which, None

    def _get(self):
        if self._host is None:
            temp, grad = self._asm._gp_temperature(self.T)
            self._host = (temp if self.which == 0 else grad).cpu().numpy()
        return self._host

    def __array__(self, dtype=None, copy=None):
        return self._get()

    def __getitem__(self, i):
        return self._get()[i]

    def __len__(self):
        return len(self._get())
