"""ctypes binding of libfdk (include/fdk.h).  There is no fallback: if the CUDA library is
missing or fails to load, every entry point raises."""

from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("FDK_LIB") or os.path.join(HERE, "_fdk.so")  # FDK_LIB: diagnostic builds (scripts/)

HEX8, TET4, TET10, QUAD4 = 0, 1, 2, 3
ELEM_IDS = {"hex8": HEX8, "tet4": TET4, "tet10": TET10, "quad4": QUAD4}
MATRIX, VECTOR, ALL = 1, 2, 3
COMPUTE_FLAGS = {"matrix": MATRIX, "vector": VECTOR, "all": ALL}

EXPORTS = [
    "fdk_last_error_string", "fdk_version", "fdk_set_option", "fdk_get_option", "fdk_debug_phase_clocks", "fdk_element_info", "fdk_element_table",
    "fdk_sym_block_keys", "fdk_sym_block_csr", "fdk_sym_expand_csr", "fdk_plan_color_blocks",
    "fdk_assemble_elastic_iso", "fdk_assemble_elastic_iso_dist", "fdk_assemble_elastic_general", "fdk_assemble_heat", "fdk_assemble_heat_tet4", "fdk_assemble_rows_elastic",
    "fdk_gp_strain_stress", "fdk_gp_strain_stress_fbar", "fdk_residual_elastic", "fdk_residual_heat", "fdk_residual_heat_gp", "fdk_gp_temperature", "fdk_gp_deformation_gradient", "fdk_j2_update", "fdk_j2_update_r1", "fdk_j2_update_from_dofs", "fdk_j2_tangent_expand", "fdk_assemble_elastic_r1",
    "fdk_gather_f64", "fdk_scatter_add_f64", "fdk_copy_segments",
    "fdk_csr_spmv", "fdk_csr_diagonal", "fdk_pcg_work_doubles", "fdk_pcg_jacobi", "fdk_bcsr_spmv", "fdk_bcsr_pcg_jacobi",
    "fdk_mpc_expand", "fdk_mpc_fold", "fdk_bcsr_pcg_jacobi_mpc", "fdk_pcg_multi_work_doubles", "fdk_bcsr_pcg_jacobi_multi",
    "fdk_gp_to_node", "fdk_gp_to_element", "fdk_gp_von_mises",
]  # fmt: skip


class FdkError(RuntimeError):
    pass


class MpcStruct(C.Structure):
    """Mirror of ``struct fdk_mpc`` (include/fdk.h)."""

    _fields_ = [
        ("n_nodal", C.c_int64),
        ("n_glob", C.c_int32),
        ("n_slave", C.c_int64),
        ("slave", C.c_void_p),
        ("master", C.c_void_p),
        ("coef", C.c_void_p),
        ("n_master", C.c_int64),
        ("mst_dof", C.c_void_p),
        ("mst_ptr", C.c_void_p),
        ("mst_slv", C.c_void_p),
        ("scratch", C.c_void_p),
    ]


MPC_SCRATCH_DOUBLES = 9 * 9 * 148


class PlanStruct(C.Structure):
    """Mirror of ``struct fdk_plan`` (include/fdk.h)."""

    _fields_ = [
        ("elem_type", C.c_int32),
        ("n_nodes", C.c_int32),
        ("n_elems", C.c_int64),
        ("n_clusters", C.c_int32),
        ("nvar", C.c_int32),
        ("blk_nnz", C.c_int64),
        ("cap_te", C.c_int32),
        ("cap_tn", C.c_int32),
        ("cap_inc", C.c_int32),
        ("cap_owned", C.c_int32),
        ("cap_slots", C.c_int32),
        ("cap_ent", C.c_int32),
        ("cap_heavy", C.c_int32),
        ("threads", C.c_int32),
        ("cl_hdr", C.c_void_p),
        ("cl_node_ptr", C.c_void_p),
        ("cl_node", C.c_void_p),
        ("cl_bptr", C.c_void_p),
        ("cl_slot_ptr", C.c_void_p),
        ("cl_finc_ptr", C.c_void_p),
        ("cl_slot_loc", C.c_void_p),
        ("cl_finc_loc", C.c_void_p),
        ("cl_inc_ptr", C.c_void_p),
        ("inc_desc", C.c_void_p),
        ("ent_src", C.c_void_p),
        ("inc_fdst", C.c_void_p),
        ("cl_te_ptr", C.c_void_p),
        ("cl_te_elem", C.c_void_p),
        ("cl_te_own", C.c_void_p),
        ("cl_lconn", C.c_void_p),
        ("te_desc", C.c_void_p),
        ("cl_tn_ptr", C.c_void_p),
        ("cl_tn_node", C.c_void_p),
        ("slot_rec", C.c_void_p),
        ("cl_heavy_ptr", C.c_void_p),
        ("heavy_slot", C.c_void_p),
        ("blk_slot", C.c_void_p),
        ("ent_pos", C.c_void_p),
    ]


_lib = None


def load():
    """Load libfdk; raises FdkError when the extension has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FdkError(
            f"{LIB_PATH} not found: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or python -m fedoo_b200.build)"
        )
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    lib.fdk_last_error_string.restype = C.c_char_p
    lib.fdk_last_error_string.argtypes = []
    lib.fdk_version.restype = i32
    lib.fdk_set_option.argtypes = [C.c_char_p, i32]
    lib.fdk_get_option.argtypes = [C.c_char_p, C.POINTER(i32)]
    lib.fdk_debug_phase_clocks.argtypes = [C.POINTER(C.c_ulonglong), i32, i32]
    lib.fdk_element_info.argtypes = [i32, C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
    lib.fdk_element_table.argtypes = [i32, vp, vp, vp]
    lib.fdk_sym_block_keys.argtypes = [i32, i64, i32, vp, vp, C.POINTER(i64), vp]
    lib.fdk_sym_block_csr.argtypes = [i32, i64, vp, vp, vp, vp]
    lib.fdk_sym_expand_csr.argtypes = [i32, i32, i32, i64, vp, vp, i32, vp, vp, vp]
    pp = C.POINTER(PlanStruct)
    lib.fdk_plan_color_blocks.argtypes = [pp, vp, vp, vp]
    lib.fdk_assemble_elastic_iso.argtypes = [pp, i32, vp, dbl, dbl, vp, vp, vp, vp, vp]
    lib.fdk_assemble_elastic_iso_dist.argtypes = [pp, i32, vp, dbl, dbl, vp, vp, vp, vp, i32, vp, i64, vp]
    lib.fdk_assemble_elastic_general.argtypes = [pp, i32, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_assemble_heat.argtypes = [pp, i32, vp, vp, dbl, vp, vp, vp, vp, vp]
    lib.fdk_assemble_heat_tet4.argtypes = [i32, i32, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp]
    lib.fdk_assemble_rows_elastic.argtypes = [i32, i32, vp, i32, i64, vp, vp, vp, vp, vp, vp, i64, i32, i32, dbl, dbl, vp, vp,
                                              i32, vp, vp, vp, vp, vp]
    lib.fdk_gp_strain_stress.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_residual_heat.argtypes = [i32, i32, i64, vp, vp, vp, dbl, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_residual_heat_gp.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_residual_elastic.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_gp_strain_stress_fbar.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_gp_temperature.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp]
    lib.fdk_gp_deformation_gradient.argtypes = [i32, i32, i64, vp, vp, vp, i32, vp, vp]
    lib.fdk_j2_update.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_j2_update_r1.argtypes = [i64, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_j2_tangent_expand.argtypes = [i64, vp, vp, vp]
    lib.fdk_j2_update_from_dofs.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_assemble_elastic_r1.argtypes = [pp, i32, vp, vp, vp, vp, vp, vp]
    lib.fdk_csr_spmv.argtypes = [i64, i64, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.fdk_csr_diagonal.argtypes = [i64, vp, vp, i32, vp, vp, vp]
    lib.fdk_bcsr_spmv.argtypes = [i32, i32, i64, vp, vp, vp, vp, vp, vp, vp]
    lib.fdk_bcsr_pcg_jacobi.argtypes = [i32, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, dbl, i32, i32, vp,
                                        C.POINTER(i32), C.POINTER(dbl), vp]
    lib.fdk_pcg_work_doubles.argtypes = [i64]
    mp = C.POINTER(MpcStruct)
    lib.fdk_mpc_expand.argtypes = [mp, vp, vp]
    lib.fdk_pcg_multi_work_doubles.argtypes = [i64, i32]
    lib.fdk_pcg_multi_work_doubles.restype = i64
    lib.fdk_bcsr_pcg_jacobi_multi.argtypes = [i32, i32, i64, vp, vp, vp, vp, i32, vp, i32, vp, vp, vp, dbl, i32, i32, vp, mp,
                                              C.POINTER(i32), C.POINTER(dbl), vp]
    lib.fdk_mpc_fold.argtypes = [mp, vp, vp]
    lib.fdk_bcsr_pcg_jacobi_mpc.argtypes = [i32, i32, i64, vp, vp, vp, vp, i32, vp, vp, vp, vp, dbl, i32, i32, vp, mp,
                                            C.POINTER(i32), C.POINTER(dbl), vp]
    lib.fdk_pcg_jacobi.argtypes = [i64, i64, vp, vp, i32, vp, vp, vp, vp, dbl, i32, i32, vp, C.POINTER(i32),
                                   C.POINTER(dbl), vp]
    lib.fdk_gp_to_node.argtypes = [i32, i32, i32, i64, vp, vp, vp, vp, i32, i64, i64, i32, vp, vp]
    lib.fdk_gp_to_element.argtypes = [i32, i64, vp, i32, i64, i64, i32, vp, vp]
    lib.fdk_gp_von_mises.argtypes = [i64, vp, i64, i64, vp, vp]
    lib.fdk_copy_segments.argtypes = [i32, vp, vp, vp, i64, vp, vp, vp]
    lib.fdk_gather_f64.argtypes = [i64, vp, vp, vp, vp]
    lib.fdk_scatter_add_f64.argtypes = [i64, vp, vp, vp, vp]
    for name in EXPORTS:
        if name not in ("fdk_last_error_string", "fdk_pcg_work_doubles", "fdk_pcg_multi_work_doubles"):
            getattr(lib, name).restype = i32
    lib.fdk_pcg_work_doubles.restype = i64
    _lib = lib
    return lib


def set_option(key, value):
    """Process-wide kernel option (include/fdk.h: fdk_set_option): 'fuse_ku', 'mma', 'iso4'."""
    check(load().fdk_set_option(key.encode(), int(value)), "fdk_set_option")


def get_option(key):
    v = C.c_int(0)
    check(load().fdk_get_option(key.encode(), C.byref(v)), "fdk_get_option")
    return v.value


def check(rc, what=""):
    if rc != 0:
        msg = load().fdk_last_error_string().decode()
        raise FdkError(f"{what} failed with code {rc}: {msg}")


def ptr(t):
    """Device (or host) pointer of a torch tensor / numpy array, NULL for None."""
    if t is None:
        return None
    if hasattr(t, "data_ptr"):
        assert t.is_contiguous(), "fdk buffers must be contiguous"
        return C.c_void_p(t.data_ptr())
    assert t.flags["C_CONTIGUOUS"] or t.flags["F_CONTIGUOUS"]
    return C.c_void_p(t.ctypes.data)


def current_stream():
    import torch

    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def element_table(elem):
    """(w, N, dN) of an element type as NumPy arrays, straight from the library."""
    import numpy as np

    lib = load()
    eid = ELEM_IDS[elem] if isinstance(elem, str) else elem
    nne, ngp, dim = C.c_int(), C.c_int(), C.c_int()
    check(lib.fdk_element_info(eid, C.byref(nne), C.byref(ngp), C.byref(dim)), "fdk_element_info")
    w = np.zeros(ngp.value)
    N = np.zeros((ngp.value, nne.value))
    dN = np.zeros((ngp.value, dim.value, nne.value))
    check(lib.fdk_element_table(eid, ptr(w), ptr(N), ptr(dN)), "fdk_element_table")
    return w, N, dN
