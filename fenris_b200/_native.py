"""ctypes binding of libfenris_b200.so (the C ABI declared in include/fenris_b200.h).

There is no CPU fallback: importing works without a GPU (so that host-side helpers and the
symbol-export check run anywhere), but creating a context without a CUDA device raises.
"""
from __future__ import annotations

import ctypes as C
import os
import re

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfenris_b200.so")
HEADER_PATH = os.path.join(os.path.dirname(_HERE), "include", "fenris_b200.h")

# element / operator / scatter ids (include/fenris_b200.h)
QUAD4, TET4, HEX8, HEX27, TET10, HEX20 = 1, 2, 3, 4, 5, 6
LAPLACE, LINEAR_ELASTIC, STVK, NEO_HOOKEAN = 1, 2, 3, 4
SCATTER_ATOMIC, SCATTER_COLORED, SCATTER_GATHER = 0, 1, 2

OK = 0
ERR_SINGULAR_JACOBIAN, ERR_SHAPE, ERR_INDEX_OOB, ERR_COLUMN_NOT_IN_PATTERN = 1, 2, 3, 4
ERR_UNSUPPORTED, ERR_CUDA, ERR_NCCL, ERR_STATE, ERR_COLORING = 5, 6, 7, 8, 9
ERR_NOT_CONVERGED, ERR_INDEFINITE = 10, 11


class Fb200Error(RuntimeError):
    def __init__(self, status: int, message: str, element_index: int = -1):
        super().__init__(f"fenris_b200 status {status}: {message}" + (f" (element {element_index})" if element_index >= 0 else ""))
        self.status = status
        self.element_index = element_index


class SingularJacobianError(Fb200Error):
    """eyre!("Singular element Jacobian encountered") - src/assembly/local/elliptic.rs:401-404."""


class Quadrature(C.Structure):
    _fields_ = [("num_points", C.c_int32), ("dim", C.c_int32), ("weights", C.POINTER(C.c_double)),
                ("points", C.POINTER(C.c_double)), ("data", C.POINTER(C.c_double))]


class Operator(C.Structure):
    _fields_ = [("kind", C.c_int32)]


_lib = None


def declared_symbols():
    """Every function name include/fenris_b200.h declares."""
    with open(HEADER_PATH) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(fb200_[a-z0-9_]+)\s*\(", src)))


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). fenris_b200 has no CPU fallback.")
    L = C.CDLL(LIB_PATH)
    vp, u64, i32, i64, dbl = C.c_void_p, C.c_uint64, C.c_int32, C.c_int64, C.c_double
    pu64, pdbl, pi32 = C.POINTER(u64), C.POINTER(dbl), C.POINTER(i32)
    sig = {
        "fb200_create": (i32, [i32, C.POINTER(vp)]),
        "fb200_destroy": (None, [vp]),
        "fb200_status_string": (C.c_char_p, [i32]),
        "fb200_last_error": (i32, [vp, C.c_char_p, C.c_size_t, C.POINTER(i64)]),
        "fb200_abi_version": (i32, []),
        "fb200_set_stream": (i32, [vp, vp]),
        "fb200_synchronize": (i32, [vp]),
        "fb200_timer_begin": (i32, [vp]),
        "fb200_timer_end": (i32, [vp, C.POINTER(C.c_float)]),
        "fb200_launch_count": (u64, [vp]),
        "fb200_set_tuning": (i32, [vp, C.c_char_p, i32]),
        "fb200_space_upload": (i32, [vp, i32, u64, vp, u64, vp]),
        "fb200_space_update_vertices": (i32, [vp, vp]),
        "fb200_connectivity_upload": (i32, [vp, u64, u64, vp, vp]),
        "fb200_set_num_owned_elements": (i32, [vp, u64]),
        "fb200_assemble_pattern": (i32, [vp, i32, pu64, pu64]),
        "fb200_pattern_download": (i32, [vp, vp, vp]),
        "fb200_pattern_adopt": (i32, [vp, i32, u64, vp, vp]),
        "fb200_color_nodes": (i32, [vp, pu64]),
        "fb200_colors_download": (i32, [vp, vp, vp]),
        "fb200_colors_adopt": (i32, [vp, u64, vp, vp]),
        "fb200_assemble_into_csr_device": (i32, [vp, C.POINTER(Operator), C.POINTER(Quadrature), vp, i32, i32]),
        "fb200_assemble_into_csr": (i32, [vp, C.POINTER(Operator), C.POINTER(Quadrature), vp, i32, i32, vp]),
        "fb200_assemble_into_csr_table_device": (i32, [vp, C.POINTER(Operator), C.c_uint32, C.POINTER(Quadrature), vp, vp, i32, i32]),
        "fb200_values_device": (i32, [vp, C.POINTER(vp), pu64]),
        "fb200_values_download": (i32, [vp, vp]),
        "fb200_values_upload": (i32, [vp, vp]),
        "fb200_element_matrices": (i32, [vp, C.POINTER(Operator), C.POINTER(Quadrature), u64, u64, vp]),
        "fb200_element_matrices_u": (i32, [vp, C.POINTER(Operator), C.POINTER(Quadrature), vp, u64, u64, vp]),
        "fb200_assemble_mass_into_csr_device": (i32, [vp, C.POINTER(Quadrature), i32, i32]),
        "fb200_assemble_mass_into_csr": (i32, [vp, C.POINTER(Quadrature), i32, i32, vp]),
        "fb200_assemble_vector": (i32, [vp, C.POINTER(Quadrature), i32, vp, i32, i32, i32, vp]),
        "fb200_physical_quadrature_points": (i32, [vp, C.POINTER(Quadrature), vp]),
        "fb200_assemble_elliptic_vector": (i32, [vp, C.POINTER(Operator), C.POINTER(Quadrature), vp, i32, i32, vp]),
        "fb200_assemble_elliptic_scalar": (i32, [vp, C.POINTER(Operator), C.POINTER(Quadrature), vp, pdbl]),
        "fb200_assemble_elliptic_vector_table": (i32, [vp, C.POINTER(Operator), C.c_uint32, C.POINTER(Quadrature), vp, vp, i32, i32, vp]),
        "fb200_assemble_elliptic_scalar_table": (i32, [vp, C.POINTER(Operator), C.c_uint32, C.POINTER(Quadrature), vp, vp, pdbl]),
        "fb200_apply_homogeneous_dirichlet_bc_csr": (i32, [vp, u64, vp, pdbl]),
        "fb200_spmv": (i32, [vp, vp, vp]),
        "fb200_cg_solve": (i32, [vp, vp, vp, dbl, u64, i32, pu64, pdbl]),
        "fb200_comm_unique_id": (i32, [C.c_char_p]),
        "fb200_comm_init": (i32, [vp, C.c_char_p, i32, i32]),
        "fb200_interface_set": (i32, [vp, u64, vp, vp, u64]),
        "fb200_interface_set_peers": (i32, [vp, u64, vp, vp, vp]),
        "fb200_interface_allreduce": (i32, [vp]),
        "fb200_interface_enable_p2p": (i32, [vp]),
        "fb200_gen_hex_mesh": (i32, [u64, u64, u64, dbl, pu64, pu64, vp, vp]),
        "fb200_gen_tet_mesh": (i32, [u64, u64, u64, dbl, pu64, pu64, vp, vp]),
        "fb200_gen_quad_mesh": (i32, [u64, u64, dbl, pu64, pu64, vp, vp]),
        "fb200_hex27_from_hex8": (i32, [u64, vp, u64, vp, pu64, vp, vp]),
        "fb200_hex20_from_hex8": (i32, [u64, vp, u64, vp, pu64, vp, vp]),
        "fb200_tet10_from_tet4": (i32, [u64, vp, u64, vp, pu64, vp, vp]),
        "fb200_canonical_quadrature": (i32, [i32, pi32, vp, vp]),
        "fb200_lame_from_young_poisson": (None, [dbl, dbl, pdbl, pdbl]),
        "fb200_tile_lists_selftest": (i32, [u64, vp, u64, vp, u64, pu64, pi32]),
        "fb200_tile_lists_selftest_ex": (i32, [u64, vp, u64, vp, u64, i32, pu64, pi32]),
        "fb200_chunk_lists_selftest": (i32, [u64, vp, u64, vp, u64, i32, i32, pu64, pi32]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(L, name)  # AttributeError = header/library mismatch: fail loudly
        fn.restype = res
        fn.argtypes = args
    _lib = L
    return L


def ptr(a):
    """void* of a C-contiguous numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(C.c_void_p)


def as_u64(a):
    return np.ascontiguousarray(a, dtype=np.uint64)


def as_f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)
