"""Host-side mirror of the fedoo objects the assembly path touches: name registries,
ModelingSpace, Mesh, and the device-resident containers returned at the boundary
(``DeviceCSR`` for ``get_global_matrix()``, ``GaussPointTensor`` for ``sv["Stress"]`` /
``sv["Strain"]``).

Mirrors fedoo/core/base.py:79-148,210-260 (registries), fedoo/core/modelingspace.py:7-120
(active space, variables), fedoo/core/mesh.py:70-200 (Mesh) and
fedoo/util/voigt_tensors.py:150-321 (tensor lists) -- API only, none of their code.
"""

from __future__ import annotations

import numpy as np
import torch

from . import _lib


def device():
    if not torch.cuda.is_available():
        raise _lib.FdkError("fedoo_b200 needs a CUDA device: there is no CPU fallback")
    return torch.device("cuda", torch.cuda.current_device())


class _Named:
    """Objects registered by name in a class-level dict (fedoo/core/base.py:79,109,213,276)."""

    _dict: dict

    def _register(self, name):
        self.name = name
        if name != "":
            type(self)._registry()[name] = self

    @classmethod
    def _registry(cls):
        for k in cls.__mro__:
            if "_dict" in k.__dict__:
                return k._dict
        raise TypeError("no registry")

    @classmethod
    def get_all(cls):
        return cls._registry()

    def __class_getitem__(cls, item):
        return cls._registry()[item]


class ModelingSpace(_Named):
    """fedoo/core/modelingspace.py:7-120: '3D', '2Dplane' (plane strain) or '2Dstress'."""

    _dict = {}
    _active = None

    def __init__(self, dimension, name="Main"):
        assert isinstance(dimension, str), "The dimension value must be a string"
        assert dimension in ("3D", "2Dplane", "2Dstress"), "Dimension must be '3D', '2Dplane' or '2Dstress'"
        self._dimension = dimension
        self.ndim = 3 if dimension == "3D" else 2
        self._variable = {}
        self._register(name)
        ModelingSpace._active = self

    def get_dimension(self):
        return self._dimension

    @staticmethod
    def get_active():
        assert ModelingSpace._active is not None, "Define a ModelingSpace before constructing any model object"
        return ModelingSpace._active

    def new_variable(self, name):
        if name not in self._variable:
            self._variable[name] = len(self._variable)

    def variable_rank(self, name):
        return self._variable[name]

    @property
    def nvar(self):
        return len(self._variable)

    def list_variables(self):
        return list(self._variable)

    def list_vectors(self):
        """fedoo/core/modelingspace.py: the displacement vector registered by StressEquilibrium."""
        return ["Disp"] if "DispX" in self._variable else []


class BoundingBox:
    def __init__(self, nodes):
        self.mins, self.maxs = nodes.min(axis=0), nodes.max(axis=0)
        self.xmin, self.ymin = self.mins[0], self.mins[1]
        self.xmax, self.ymax = self.maxs[0], self.maxs[1]
        if nodes.shape[1] > 2:
            self.zmin, self.zmax = self.mins[2], self.maxs[2]

    def __iter__(self):
        return iter((self.mins, self.maxs))

    @property
    def center(self):
        return (self.mins + self.maxs) / 2

    @property
    def size(self):
        return self.maxs - self.mins

    @property
    def volume(self):
        return float(np.prod(self.size))


class Mesh(_Named):
    """fedoo/core/mesh.py:70: nodes (n, dim) float64, elements (n_el, nne) ints, elm_type."""

    _dict = {}
    SUPPORTED = ("hex8", "tet4", "tet10", "quad4")

    def __init__(self, nodes, elements=None, elm_type=None, node_sets=None, element_sets=None, ndim=None, name=""):
        self.nodes = np.ascontiguousarray(nodes, dtype=float)
        self.elements = np.ascontiguousarray(elements)
        self.elm_type = elm_type.lower() if elm_type else elm_type
        self.node_sets = {} if node_sets is None else node_sets
        self.element_sets = {} if element_sets is None else element_sets
        self._dev = None
        self._register(name)

    @property
    def n_nodes(self):
        return self.nodes.shape[0]

    @property
    def n_elements(self):
        return self.elements.shape[0]

    @property
    def n_elm_nodes(self):
        return self.elements.shape[1]

    @property
    def ndim(self):
        return self.nodes.shape[1]

    @property
    def bounding_box(self):
        return BoundingBox(self.nodes)

    def find_nodes(self, selection_criterion, value=0, tol=1e-6):
        """Subset of fedoo/core/mesh.py find_nodes: 'X', 'Y', 'Z' planes."""
        axis = {"X": 0, "Y": 1, "Z": 2}[selection_criterion.upper()]
        return np.where(np.abs(self.nodes[:, axis] - value) < tol)[0]

    def device_arrays(self):
        """(coords float64 (n, dim), conn int32 (n_el, nne)) on the current CUDA device."""
        if self._dev is None:
            dev = device()
            coords = torch.from_numpy(self.nodes).to(dev)
            conn = torch.from_numpy(self.elements.astype(np.int32)).to(dev)
            self._dev = (coords, conn)
        return self._dev

    def invalidate_device(self):
        self._dev = None


class DeviceCSR:
    """Device-resident CSR matrix returned by ``Assembly.get_global_matrix()``.

    The reference returns a host ``scipy.sparse.csr_matrix`` (fedoo/core/assembly.py:452-460).
    At 8 M hex8 elements K is 15.6 GB of values + 7.8 GB of indices, so the matrix stays in HBM
    and the scipy object is materialised on first host access (``tocsr()`` or any scipy-like
    attribute); ``indptr`` / ``indices`` / ``data`` are the device tensors."""

    def __init__(self, indptr, indices, data, shape, block=None, n_glob=0):
        self.indptr, self.indices, self.data = indptr, indices, data
        self.shape = tuple(shape)
        self._host = None
        # (blk_indptr int64, blk_indices int32, nvar, n_nodes) of the tiled pattern when the matrix comes from the
        # assembly: matvec / pcg then read every block row's column list once.  n_glob: trailing empty rows / columns
        # of global (non-nodal) dofs (fedoo/core/assembly.py:192-197,459-460)
        self.block = block
        self.n_glob = int(n_glob)

    @property
    def n_nodal(self):
        return self.shape[0] - self.n_glob

    @property
    def nnz(self):
        return int(self.data.numel())

    def tocsr(self):
        if self._host is None:
            from scipy import sparse

            self._host = sparse.csr_matrix(
                (self.data.cpu().numpy(), self.indices.cpu().numpy(), self.indptr.cpu().numpy()), shape=self.shape
            )
        return self._host

    to_scipy = tocsr

    def resize(self, *shape):
        """In-place resize to account for global dofs (fedoo/core/problem.py:282-284): only
        trailing empty rows/columns can be added."""
        if len(shape) == 1:
            shape = shape[0]
        n = int(shape[0])
        assert n >= self.shape[0] and shape[0] == shape[1]
        extra = n - self.shape[0]
        if extra:
            self.indptr = torch.cat([self.indptr, self.indptr[-1:].expand(extra)])
            self.shape = (n, n)
            self._host = None
            self.n_glob += extra

    def _index_bytes(self):
        assert self.indptr.dtype == self.indices.dtype, "indptr and indices share one integer type (scipy convention)"
        return self.indptr.element_size()

    def matvec(self, x, free_mask=None, out=None):
        """y = A x on the device (csrc/fdk_solve.cuh); with ``free_mask`` (uint8 per dof, 0 = imposed) rows and
        columns of imposed dofs are left out: the Dirichlet elimination of fedoo/core/problem.py:277-298
        without forming MatCB^T A MatCB."""
        from . import _lib

        x = as_device_f64(x, self.data.device)
        assert x.numel() == self.shape[1]
        y = torch.empty(self.shape[0], dtype=torch.float64, device=self.data.device) if out is None else out
        if self.block is not None:
            bp, bi, nvar, n_nodes = self.block
            _lib.check(
                _lib.load().fdk_bcsr_spmv(n_nodes, nvar, int(bi.numel()), _lib.ptr(bp), _lib.ptr(bi), _lib.ptr(self.data),
                                          _lib.ptr(x), _lib.ptr(free_mask), _lib.ptr(y), _lib.current_stream()),
                "fdk_bcsr_spmv",
            )  # fmt: skip
            if self.n_glob:
                y[self.n_nodal :] = 0.0  # the rows of the global dofs are empty
            return y
        _lib.check(
            _lib.load().fdk_csr_spmv(
                self.shape[0], self.nnz, _lib.ptr(self.indptr), _lib.ptr(self.indices), self._index_bytes(),
                _lib.ptr(self.data), _lib.ptr(x), _lib.ptr(free_mask), _lib.ptr(y), _lib.current_stream(),
            ),
            "fdk_csr_spmv",
        )
        return y

    def diagonal_device(self):
        from . import _lib

        d = torch.empty(self.shape[0], dtype=torch.float64, device=self.data.device)
        _lib.check(
            _lib.load().fdk_csr_diagonal(
                self.shape[0], _lib.ptr(self.indptr), _lib.ptr(self.indices), self._index_bytes(), _lib.ptr(self.data),
                _lib.ptr(d), _lib.current_stream(),
            ),
            "fdk_csr_diagonal",
        )
        return d

    def pcg(self, b, free_mask=None, rtol=1e-8, maxiter=None, check_every=10, mpc=None):
        """Jacobi-preconditioned CG on the device (scipy.sparse.linalg.cg with M = diag(1 / A.diagonal()),
        fedoo/core/base.py:521-537).  Returns (x, iterations, ||r|| / ||b||), x a device tensor.
        With ``mpc`` (constraint.MpcMap) the system is T^T A T on n_nodal + n_glob entries: b already folded,
        free_mask 0 on imposed dofs and slaves, x = the independent dofs."""
        import ctypes as C

        from . import _lib

        lib = _lib.load()
        if mpc is not None:
            if self.block is None:
                raise NotImplementedError("the constrained solve runs on the tiled pattern of the assembly (nodal size)")
            bp, bi, nvar, n_nodes = self.block
            n = mpc.n_total
            assert self.n_nodal == mpc.n_nodal and free_mask is not None and free_mask.numel() == n
            b = as_device_f64(b, self.data.device)
            assert b.numel() == n
            x = torch.empty(n, dtype=torch.float64, device=self.data.device)
            work = torch.empty(int(lib.fdk_pcg_work_doubles(n)), dtype=torch.float64, device=self.data.device)
            it, rel = C.c_int(0), C.c_double(0.0)
            _lib.check(
                lib.fdk_bcsr_pcg_jacobi_mpc(
                    n_nodes, nvar, int(bi.numel()), _lib.ptr(bp), _lib.ptr(bi), _lib.ptr(self.indptr), _lib.ptr(self.indices),
                    self._index_bytes(), _lib.ptr(self.data), _lib.ptr(b), _lib.ptr(x), _lib.ptr(free_mask), float(rtol),
                    int(10 * n if maxiter is None else maxiter), int(check_every), _lib.ptr(work), mpc.struct(),
                    C.byref(it), C.byref(rel), _lib.current_stream(),
                ),
                "fdk_bcsr_pcg_jacobi_mpc",
            )  # fmt: skip
            return x, it.value, rel.value
        n = self.shape[0]
        b = as_device_f64(b, self.data.device)
        x = torch.empty(n, dtype=torch.float64, device=self.data.device)
        work = torch.empty(int(lib.fdk_pcg_work_doubles(n)), dtype=torch.float64, device=self.data.device)
        it, rel = C.c_int(0), C.c_double(0.0)
        if self.block is not None and self.n_glob == 0:
            bp, bi, nvar, n_nodes = self.block
            _lib.check(
                lib.fdk_bcsr_pcg_jacobi(
                    n_nodes, nvar, int(bi.numel()), _lib.ptr(bp), _lib.ptr(bi), _lib.ptr(self.indptr), _lib.ptr(self.indices),
                    self._index_bytes(), _lib.ptr(self.data), _lib.ptr(b), _lib.ptr(x), _lib.ptr(free_mask), float(rtol),
                    int(10 * n if maxiter is None else maxiter), int(check_every), _lib.ptr(work), C.byref(it), C.byref(rel),
                    _lib.current_stream(),
                ),
                "fdk_bcsr_pcg_jacobi",
            )  # fmt: skip
            return x, it.value, rel.value
        _lib.check(
            lib.fdk_pcg_jacobi(
                n, self.nnz, _lib.ptr(self.indptr), _lib.ptr(self.indices), self._index_bytes(), _lib.ptr(self.data),
                _lib.ptr(b), _lib.ptr(x), _lib.ptr(free_mask), float(rtol), int(10 * n if maxiter is None else maxiter),
                int(check_every), _lib.ptr(work), C.byref(it), C.byref(rel), _lib.current_stream(),
            ),
            "fdk_pcg_jacobi",
        )
        return x, it.value, rel.value

    def pcg_multi(self, B, free_mask=None, rtol=1e-8, maxiter=None, check_every=10, mpc=None):
        """R right-hand sides in lockstep (R = 3 or 6): B, X are (n, R) device tensors, K is read once per iteration
        for all of them (csrc/fdk_solve.cuh: k_bcsr_spmm + per-column CG recurrences).  With ``mpc`` B must be folded
        and X comes back expanded.  Returns (X, iterations, [||r_k|| / ||b_k||])."""
        import ctypes as C

        from . import _lib

        if self.block is None:
            raise NotImplementedError("the multi-right-hand-side solve runs on the tiled pattern of the assembly")
        lib = _lib.load()
        bp, bi, nvar, n_nodes = self.block
        if mpc is None and self.n_glob:
            raise NotImplementedError("global dofs without a constraint map")
        n = self.shape[0] if mpc is None else mpc.n_total
        B = as_device_f64(B, self.data.device)
        R = int(B.shape[1])
        assert B.shape[0] == n and B.is_contiguous()
        X = torch.empty((n, R), dtype=torch.float64, device=self.data.device)
        work = torch.empty(int(lib.fdk_pcg_multi_work_doubles(n, R)), dtype=torch.float64, device=self.data.device)
        it, rel = C.c_int(0), (C.c_double * R)()
        _lib.check(
            lib.fdk_bcsr_pcg_jacobi_multi(
                n_nodes, nvar, int(bi.numel()), _lib.ptr(bp), _lib.ptr(bi), _lib.ptr(self.indptr), _lib.ptr(self.indices),
                self._index_bytes(), _lib.ptr(self.data), R, _lib.ptr(B), _lib.ptr(X), _lib.ptr(free_mask), float(rtol),
                int(10 * n if maxiter is None else maxiter), int(check_every), _lib.ptr(work),
                None if mpc is None else mpc.struct(), C.byref(it), rel, _lib.current_stream(),
            ),
            "fdk_bcsr_pcg_jacobi_multi",
        )  # fmt: skip
        return X, it.value, list(rel)

    def __matmul__(self, x):
        if isinstance(x, torch.Tensor) and x.is_cuda:
            return self.matvec(x)
        return self.tocsr() @ x

    def __rmatmul__(self, x):
        return x @ self.tocsr()

    def diagonal(self):
        return self.tocsr().diagonal()

    def __getattr__(self, item):  # any other scipy attribute: materialise
        if item.startswith("_"):
            raise AttributeError(item)
        return getattr(self.tocsr(), item)


class GaussPointTensor:
    """Voigt tensor field at the Gauss points kept on the device: the (6, N) Fortran-ordered
    array of the reference's StressTensorList / StrainTensorList
    (fedoo/util/voigt_tensors.py:150-321) stored as a contiguous (N, 6) CUDA tensor."""

    def __init__(self, tensor, kind="stress"):
        self.device_tensor = tensor  # (N, 6)
        self.kind = kind
        self._host = None

    def asarray(self):
        if self._host is None:
            self._host = self.device_tensor.cpu().numpy().T  # (6, N), F-contiguous view
        return self._host

    @property
    def array(self):
        return self.asarray()

    def __getitem__(self, i):
        return self.asarray()[i]

    def __len__(self):
        return 6

    def von_mises(self):
        """fedoo/util/voigt_tensors.py:270-283 (stress convention)."""
        s = self.asarray()
        return np.sqrt(
            0.5 * ((s[0] - s[1]) ** 2 + (s[1] - s[2]) ** 2 + (s[0] - s[2]) ** 2 + 6 * (s[3] ** 2 + s[4] ** 2 + s[5] ** 2))
        )


def as_device_f64(x, dev=None):
    """numpy / torch (host or device) -> contiguous float64 CUDA tensor."""
    dev = dev or device()
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64, non_blocking=True).contiguous()
    a = np.ascontiguousarray(x, dtype=np.float64)
    return torch.from_numpy(a).to(dev, non_blocking=True)
