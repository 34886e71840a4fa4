"""One-time symbolic phase on the device: block pattern, tiled CSR, cluster plan.

Replaces the symbolic half of ``_BlocSparse.tocsr`` (fedoo/core/_sparsematrix.py:225-284)
and the ``scipy.sparse.bmat`` tiling (:310-315); the result is bit-exact with the pattern the
reference builds with NumPy/SciPy (int32 indices unless max(nnz, n_rows) > 2^31-1, scipy's
``get_index_dtype`` rule).
"""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from .plan import Pattern, Plan

INT32_MAX = 2**31 - 1


def build_pattern(conn: torch.Tensor, n_nodes: int) -> Pattern:
    """conn: (n_el, nne) int32 CUDA tensor."""
    lib = _lib.load()
    if not conn.is_cuda:
        raise _lib.FdkError("build_pattern needs CUDA tensors (there is no CPU path)")
    conn = conn.contiguous()
    n_el, nne = conn.shape
    total = n_el * nne * nne
    keys = torch.empty(max(total, 1), dtype=torch.int64, device=conn.device)
    nnz = C.c_int64(0)
    _lib.check(
        lib.fdk_sym_block_keys(n_nodes, n_el, nne, _lib.ptr(conn), _lib.ptr(keys), C.byref(nnz), _lib.current_stream()),
        "fdk_sym_block_keys",
    )
    blk_nnz = nnz.value
    keys = keys[:blk_nnz].clone()
    indptr = torch.empty(n_nodes + 1, dtype=torch.int64, device=conn.device)
    indices = torch.empty(max(blk_nnz, 1), dtype=torch.int32, device=conn.device)
    _lib.check(
        lib.fdk_sym_block_csr(n_nodes, blk_nnz, _lib.ptr(keys), _lib.ptr(indptr), _lib.ptr(indices), _lib.current_stream()),
        "fdk_sym_block_csr",
    )
    return Pattern(n_nodes, indptr, indices[:blk_nnz], keys)


def csr_index_dtype(nnz: int, n_rows: int):
    """scipy.sparse get_index_dtype rule used by bmat (scipy _construct.py:1017-1018,1073)."""
    return torch.int32 if max(nnz, n_rows) <= INT32_MAX else torch.int64


def expand_csr(pattern: Pattern, nvar: int, n_global_dof: int = 0):
    """Global (indptr, indices) device tensors of the nvar x nvar tiled matrix."""
    lib = _lib.load()
    n = pattern.n_nodes
    nnz = nvar * nvar * pattern.blk_nnz
    n_rows = nvar * n + n_global_dof
    dt = csr_index_dtype(nnz, n_rows)
    dev = pattern.blk_indptr.device
    indptr = torch.empty(n_rows + 1, dtype=dt, device=dev)
    indices = torch.empty(max(nnz, 1), dtype=dt, device=dev)
    _lib.check(
        lib.fdk_sym_expand_csr(
            n, nvar, n_global_dof, pattern.blk_nnz, _lib.ptr(pattern.blk_indptr), _lib.ptr(pattern.blk_indices),
            4 if dt == torch.int32 else 8, _lib.ptr(indptr), _lib.ptr(indices), _lib.current_stream(),
        ),
        "fdk_sym_expand_csr",
    )  # fmt: skip
    return indptr, indices[:nnz]


def build_plan(elem_type: str, coords: torch.Tensor, conn: torch.Tensor, pattern: Pattern | None = None, **kw) -> Plan:
    if pattern is None:
        pattern = build_pattern(conn, coords.shape[0])
    return Plan(elem_type, coords, conn, pattern, **kw)
