"""CPU-side checks of the C ABI: the shared library loads without a GPU, exports every symbol
include/fdk.h declares, and its element tables equal the oracle's (host-only calls)."""

import ctypes
import os
import re

import numpy as np
import pytest

from fedoo_b200 import _lib
from oracle import fedoo_oracle as fo

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        from fedoo_b200 import build

        build.build()
    return _lib.load()


def test_header_symbols_exported(lib):
    header = open(os.path.join(ROOT, "include", "fdk.h")).read()
    declared = set(re.findall(r"\b(fdk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    raw = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(raw, name), name
    assert lib.fdk_version() >= 100


def test_plan_struct_matches_header():
    header = open(os.path.join(ROOT, "include", "fdk.h")).read()
    body = header[header.index("typedef struct fdk_plan {") : header.index("} fdk_plan;")]
    fields = re.findall(r"\b([a-z_0-9]+)\s*(?:,|;)", re.sub(r"/\*.*?\*/", "", body, flags=re.S))
    fields = [f for f in fields if f not in ("fdk_plan",)]
    assert fields == [f[0] for f in _lib.PlanStruct._fields_]


@pytest.mark.parametrize("elem", ["hex8", "tet4", "tet10", "quad4"])
def test_element_tables_match_oracle(lib, elem):
    w, N, dN = _lib.element_table(elem)
    tab = fo.element_table(elem)
    assert w.shape == (tab.ngp,) and N.shape == (tab.ngp, tab.nne) and dN.shape == (tab.ngp, tab.dim, tab.nne)
    assert np.abs(w - tab.w_gp).max() <= 1e-16
    assert np.abs(N - tab.N).max() <= 2e-16
    assert np.abs(dN - tab.dN).max() <= 5e-16
    assert np.abs(N.sum(axis=1) - 1).max() <= 1e-15  # partition of unity
    assert np.abs(dN.sum(axis=2)).max() <= 1e-15


def test_error_string_and_bad_element(lib):
    nne, ngp, dim = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
    assert lib.fdk_element_info(7, ctypes.byref(nne), ctypes.byref(ngp), ctypes.byref(dim)) == -1
    assert b"unknown element type" in lib.fdk_last_error_string()


def test_no_cpu_fallback():
    """The product path refuses to run without a CUDA device."""
    import torch

    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import fedoo_b200 as fd

    fd.ModelingSpace("3D")
    m = fd.mesh.box_mesh(4, 4, 4, name="nofallback")
    law = fd.constitutivelaw.ElasticIsotrop(1.0, 0.3, name="nofallback_law")
    a = fd.Assembly.create(fd.weakform.StressEquilibrium(law, name="nofallback_wf"), m)
    with pytest.raises(fd.FdkError):
        a.assemble_global_mat()


def test_solver_callable_has_no_cpu_fallback():
    """fedoo_b200.solver.pcg (the callable handed to the reference's Problem.set_solver) validates its arguments on the
    host and then needs the device: without one it raises, it does not quietly solve on the CPU."""
    import numpy as np
    import torch
    from scipy import sparse

    import fedoo_b200 as fd

    A = sparse.identity(5, format="csr") * 2.0
    with pytest.raises(ValueError):
        fd.solver.pcg(sparse.csr_matrix(np.ones((3, 4))), np.ones(3))
    with pytest.raises(ValueError):
        fd.solver.pcg(A, np.ones(4))
    assert fd.solver.pcg(sparse.csr_matrix((0, 0)), np.zeros(0)).shape == (0,)
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    with pytest.raises(fd.FdkError):
        fd.solver.pcg(A, np.ones(5))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "fedoo_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("fedoo_oracle", "oracle") or f == "__none__", (
                    f"{f} mentions the oracle: the product path must not depend on it"
                )
