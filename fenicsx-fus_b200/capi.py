"""ctypes binding of the C ABI in include/fus_b200.h.

The shared library must exist (built by build.py / __graft_entry__.build()); there is no Python
or CPU fallback: a missing library or a missing GPU raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libfus_b200.so")

FUS_OK = 0
FUS_LINEAR, FUS_LOSSY, FUS_WESTERVELT = 0, 1, 2
KINDS = {"linear": FUS_LINEAR, "lossy": FUS_LOSSY, "westervelt": FUS_WESTERVELT}

_i32 = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_f64 = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_f32 = np.ctypeslib.ndpointer(dtype=np.float32, flags="C_CONTIGUOUS")
_i64 = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")
_p = C.c_void_p
_ll = C.c_int64
_int = C.c_int
_dbl = C.c_double

# name -> (restype, argtypes); exactly the symbols include/fus_b200.h declares
SIGNATURES = {
    "fus_last_error": (C.c_char_p, []),
    "fus_version": (_int, []),
    "fus_device_count": (_int, []),
    "fus_gll": (_int, [_int, _f64, _f64]),
    "fus_tabulate_dphi": (_int, [_int, _f64]),
    "fus_box_mesh": (_int, [_i32, _f64, _f64, _f64, _i32]),
    "fus_box_dofmap": (_int, [_int, _i32, _int, _i32]),
    "fus_box_num_dofs": (_ll, [_int, _i32]),
    "fus_box_facets": (_ll, [_i32, _p]),
    "fus_boundary_vectors": (_int, [_int, _int, _ll, _ll, _f64, _i32, _i32, _ll, _i32, _f64, _f64,
                                    _p, _p, _p, _p, _p]),
    "fus_trilinear_coeffs": (_int, [_ll, _f64, _i32, _f64]),
    "fus_trilinear_geometry": (_int, [_int, _ll, _f64, _p, _p]),
    "fus_ctx_create": (_int, [_int, _ll, _ll, _ll, _i32, _p, _p, _f64, _int, C.POINTER(_p)]),
    "fus_ctx_create_from_mesh": (_int, [_int, _ll, _ll, _ll, _i32, _ll, _f64, _i32, _int,
                                        C.POINTER(_p)]),
    "fus_ctx_create_from_mesh_lean": (_int, [_int, _ll, _ll, _ll, _i32, _ll, _f64, _i32, _int,
                                             C.POINTER(_p)]),
    "fus_rect_mesh": (_int, [_i32, _f64, _f64, _f64, _i32]),
    "fus_rect_dofmap": (_int, [_int, _i32, _i32]),
    "fus_rect_num_dofs": (_ll, [_int, _i32]),
    "fus_rect_facets": (_ll, [_i32, _p]),
    "fus_boundary_vectors_2d": (_int, [_int, _int, _ll, _ll, _f64, _i32, _i32, _ll, _i32, _f64,
                                       _f64, _p, _p, _p, _p, _p]),
    "fus_ctx_create_2d": (_int, [_int, _ll, _ll, _ll, _i32, _p, _p, _f64, _int, C.POINTER(_p)]),
    "fus_ctx_create_from_mesh_2d": (_int, [_int, _ll, _ll, _ll, _i32, _ll, _f64, _i32, _int,
                                           C.POINTER(_p)]),
    "fus_ctx_destroy": (_int, [_p]),
    "fus_ctx_set_stream": (_int, [_p, _p]),
    "fus_ctx_set_option": (_int, [_p, C.c_char_p, _int]),
    "fus_ctx_get_option": (_int, [_p, C.c_char_p, C.POINTER(_int)]),
    "fus_ctx_sync": (_int, [_p]),
    "fus_ctx_get_geometry": (_int, [_p, _p, _p]),
    "fus_stiffness_apply_dev": (_int, [_p, _p, _p, _p]),
    "fus_stiffness_apply_host": (_int, [_p, _f64, _f64, _f64]),
    "fus_mass_apply_dev": (_int, [_p, _p, _p, _p]),
    "fus_mass_apply_host": (_int, [_p, _f64, _f64, _f64]),
    "fus_stiffness_apply_f32_dev": (_int, [_p, _p, _p, _p]),
    "fus_stiffness_apply_f32_host": (_int, [_p, _f32, _f32, _f32]),
    "fus_mass_apply_f32_dev": (_int, [_p, _p, _p, _p]),
    "fus_mass_apply_f32_host": (_int, [_p, _f32, _f32, _f32]),
    "fus_dev_alloc": (_int, [_p, C.c_size_t, C.POINTER(_p)]),
    "fus_dev_free": (_int, [_p, _p]),
    "fus_dev_upload": (_int, [_p, _p, _p, C.c_size_t]),
    "fus_dev_download": (_int, [_p, _p, _p, C.c_size_t]),
    "fus_dev_memset": (_int, [_p, _p, _int, C.c_size_t]),
    "fus_model_create": (_int, [_p, _int, _f64, _f64, _p, _p, _p, _p, _p, _p, _dbl, _dbl, _dbl,
                                C.POINTER(_p)]),
    "fus_model_destroy": (_int, [_p]),
    "fus_model_set_state": (_int, [_p, _p, _p]),
    "fus_model_get_state": (_int, [_p, _p, _p]),
    "fus_model_state_dev": (_int, [_p, C.POINTER(_p), C.POINTER(_p)]),
    "fus_model_get_mass": (_int, [_p, _f64]),
    "fus_model_f1": (_int, [_p, _dbl, _f64, _f64, _f64]),
    "fus_model_rk4": (_int, [_p, _dbl, _dbl, _dbl, C.POINTER(_int)]),
    "fus_launch_count": (_ll, []),
    "fus_ctx_profile": (_int, [_p, C.c_char_p, C.POINTER(_ll), C.POINTER(_dbl)]),
    "fus_comm_unique_id": (_int, [_p]),
    "fus_halo_setup": (_int, [_p, _int, _int, _p, _int, _p, _p, _p, _p, _p, _ll]),
    "fus_halo_peer_export": (_int, [_p, _p, _p, C.POINTER(_p)]),
    "fus_halo_mailbox_layout": (_int, [_ll, _ll, _int, _i64]),
    "fus_halo_peer_offsets": (_int, [_i64, _i64, _i64, _int, _i64]),
    "fus_halo_peer_connect": (_int, [_p, _p, _p]),
    "fus_halo_peer_connect_local": (_int, [_p, _p, _p, _p]),
    "fus_box_partition_create": (_int, [_int, _i32, _i32, _int, _int, C.POINTER(_p)]),
    "fus_box_partition_info": (_int, [_p, _i64, _i32, _i32]),
    "fus_box_partition_arrays": (_int, [_p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
    "fus_box_partition_destroy": (_int, [_p]),
    "fus_scatter_fwd_dev": (_int, [_p, _p]),
    "fus_scatter_rev_dev": (_int, [_p, _p]),
}

_lib = None


class FusError(RuntimeError):
    pass


def load():
    """Load libfus_b200.so.  Raises if it has not been built -- never falls back."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FusError(
            f"{LIB_PATH} is missing: build it with `python fenicsx-fus_b200/build.py` "
            "(or __graft_entry__.build()); there is no fallback implementation")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc, what=""):
    if rc != FUS_OK:
        msg = load().fus_last_error().decode(errors="replace")
        raise FusError(f"{what} failed with code {rc}: {msg}")


def optional(arr):
    """void* of a contiguous float64 array, or NULL."""
    if arr is None:
        return None
    assert arr.dtype == np.float64 and arr.flags.c_contiguous
    return arr.ctypes.data_as(_p)
