"""ctypes binding of libgsb200.so (include/gsb200.h).  No torch types cross this boundary.

The library is the product: if it is missing, or no CUDA device is visible, calls raise -- there
is no CPU path behind this module.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# GSB_LIB_PATH: development aid -- load an experimental build of the same library (tools/exp/)
LIB_PATH = os.environ.get("GSB_LIB_PATH") or os.path.join(_HERE, "libgsb200.so")

GSB_F64, GSB_I32 = 0, 1
ORDER_AUTO, ORDER_REDBLACK, ORDER_MULTICOLOR, ORDER_USER = 0, 1, 2, 3
UNIQUE_ID_BYTES = 128

STATUS = {0: "GSB_OK", 1: "GSB_ERR_ARG", 2: "GSB_ERR_SHAPE", 3: "GSB_ERR_UNSORTED", 4: "GSB_ERR_CUDA",
          5: "GSB_ERR_NCCL", 6: "GSB_ERR_NO_DEVICE", 7: "GSB_ERR_ALLOC", 8: "GSB_ERR_COLORING",
          9: "GSB_ERR_OVERFLOW", 10: "GSB_ERR_STATE"}


class GsbError(RuntimeError):
    def __init__(self, status, where, text):
        self.status = status
        super().__init__("%s: %s (%d): %s" % (where, STATUS.get(status, "?"), status, text))


class GsOptions(C.Structure):
    _fields_ = [("ordering", C.c_int), ("check_every", C.c_int), ("batch_sweeps", C.c_int), ("use_graph", C.c_int),
                ("kernel", C.c_int), ("compute_residual", C.c_int), ("fused_lead", C.c_int),
                ("reserved", C.c_int * 1)]


class GsStats(C.Structure):
    _fields_ = [("sweeps", C.c_int), ("n_colors", C.c_int), ("ordering_used", C.c_int), ("kernel_used", C.c_int),
                ("kernel_launches", C.c_int64), ("last_eps", C.c_double * 4), ("residual_l2", C.c_double * 4),
                ("solve_ms", C.c_double), ("setup_ms", C.c_double)]


class GdfOptions(C.Structure):
    _fields_ = [("solver", C.c_int), ("max_iteration", C.c_int), ("epsilon", C.c_double), ("gs", GsOptions),
                ("reserved", C.c_int * 4)]


class GdfStats(C.Structure):
    _fields_ = [("iterations", C.c_int * 4), ("last_eps", C.c_double * 4), ("residual_l2", C.c_double * 4),
                ("solve_ms", C.c_double), ("total_ms", C.c_double)]


GDF_GS, GDF_CG = 0, 1

_vp, _i, _i64, _d = C.c_void_p, C.c_int, C.c_int64, C.c_double
_ip, _i64p, _dp = C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_double)

# name -> argtypes; every symbol include/gsb200.h declares is listed (tests/test_abi.py checks both ways)
SIGNATURES = {
    "gsb_last_error": [],
    "gsb_version": [],
    "gsb_device_count": [_ip],
    "gsb_set_device": [_i],
    "gsb_stream": [],
    "gsb_matrix_create": [C.POINTER(_vp), _i],
    "gsb_matrix_destroy": [_vp],
    "gsb_matrix_assemble_sorted_coo": [_vp, _vp, _vp, _vp, _i64],
    "gsb_matrix_assemble_coo": [_vp, _vp, _vp, _vp, _i64, _i, _i],
    "gsb_matrix_import_csr": [_vp, _vp, _i, _vp, _i, _vp, _i, _vp, _i],
    "gsb_matrix_upload": [_vp, _vp, _vp, _i64, _vp, _vp, _vp, _i, _i],
    "gsb_matrix_shape": [_vp, _i64p, _ip, _ip, _i64p],
    "gsb_matrix_download": [_vp, _vp, _vp, _vp, _vp, _vp],
    "gsb_csr_from_sorted_coo": [_i, _vp, _vp, _vp, _i64, _vp, _vp, _vp, _vp, _vp, _ip, _ip],
    "gsb_csr_import": [_i, _vp, _i, _vp, _i, _vp, _i, _vp, _i, _vp, _vp, _vp, _vp, _vp],
    "gsb_matrix_at": [_vp, _vp, _vp, _i64, _vp],
    "gsb_matrix_analyze": [_vp, _i, _vp],
    "gsb_matrix_coloring": [_vp, _ip, _ip, _ip],
    "gsb_matrix_ordering": [_vp, _vp, _vp],
    "gsb_gs_default_options": [C.POINTER(GsOptions)],
    "gsb_gauss_seidel": [_vp, _vp, _i, _d, _i, C.POINTER(GsOptions), _vp, C.POINTER(GsStats)],
    "gsb_gauss_seidel_dev": [_vp, _vp, _i, _d, _i, C.POINTER(GsOptions), _vp, C.POINTER(GsStats)],
    "gsb_gauss_seidel_x0": [_vp, _vp, _vp, _i, _d, _i, C.POINTER(GsOptions), _vp, C.POINTER(GsStats)],
    "gsb_spmv": [_vp, _vp, _vp],
    "gsb_spmv_dev": [_vp, _vp, _vp],
    "gsb_residual_l2": [_vp, _vp, _vp, _dp],
    "gsb_residual_l2_dev": [_vp, _vp, _vp, _dp],
    "gsb_l1_dist": [_vp, _vp, _i64, _dp],
    "gsb_dot": [_vp, _vp, _i64, _dp],
    "gsb_axpy": [_vp, _vp, _d, _i64, _vp],
    "gsb_vecmul": [_vp, _vp, _i64, _vp],
    "gsb_conjugate_gradient": [_vp, _vp, _d, _i, _vp, _vp, _ip],
    "gsb_conjugate_gradient_multi": [_vp, _vp, _i, _d, _i, _vp, _vp, _vp],
    "gsb_conjugate_gradient_jacobi": [_vp, _vp, _d, _i, _vp, _ip],
    "gsb_poisson_matrix": [_vp, _i, _i],
    "gsb_poisson_rhs": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "gsb_poisson_rhs_dev": [_i, _i, _i, _vp, _vp, _vp, _vp],
    "gsb_poisson_rhs_rows": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "gsb_poisson_rhs_rows_dev": [_i, _i, _i, _i, _i, _vp, _vp, _vp, _vp],
    "gsb_writeback_u8": [_vp, _i64, _vp],
    "gsb_writeback_u8_dev": [_vp, _i64, _vp],
    "gsb_gdf_default_options": [C.POINTER(GdfOptions)],
    "gsb_gdf_gradients": [_vp, _i, _vp, _i, _i, _vp, _vp],
    "gsb_gdf_composite": [_vp, _i, _vp, _i, _i, _vp],
    "gsb_gdf_solve": [_i, _i, _vp, _vp, _vp, _vp, C.POINTER(GdfOptions), _vp, C.POINTER(GdfStats)],
    "gsb_gdf_fuse": [_vp, _i, _vp, _i, _i, _i, C.POINTER(GdfOptions), _vp, C.POINTER(GdfStats)],
    "gsb_gdf_release": [],
    "gsb_pano_mask_image": [_vp, _vp, _i, _i, _vp],
    "gsb_pano_gradients": [_vp, _i, _i, _vp, _vp],
    "gsb_pano_gradients_masked": [_vp, _vp, _i, _i, _vp, _vp],
    "gsb_pano_merge2_f32": [_vp, _vp, _vp, _vp, _vp, _i, _i],
    "gsb_pano_merge_u8": [_vp, _vp, _vp, _vp, _i, _d, _i, _i],
    "gsb_pano_enforce_gradient_bound": [_vp, _vp, _vp, _vp, _i, _i],
    "gsb_pano_merge_step": [_vp, _vp, _vp, _vp, _vp, _vp, _vp, _i, _i],
    "gsb_pano_split_planes_f32": [_vp, _i, _i, _vp],
    "gsb_dist_unique_id": [_vp],
    "gsb_dist_init": [C.POINTER(_vp), _vp, _i, _i, _i],
    "gsb_dist_finalize": [_vp],
    "gsb_dist_poisson_strip": [_vp, _i, _i, _i, _i],
    "gsb_dist_matrix_rows": [_vp, _vp, _vp, _vp, _i64, _i, _i64, _i],
    "gsb_dist_gauss_seidel_dev": [_vp, _vp, _i, _d, _i, C.POINTER(GsOptions), _vp, C.POINTER(GsStats)],
    "gsb_dist_residual_l2_dev": [_vp, _vp, _vp, _dp],
    "gsb_dist_set_colors": [_vp, _vp, _i64],
    "gsb_set_devices": [_vp, _i],
    "gsb_get_devices": [_vp, _i],
    "gsb_dist_init_local": [C.POINTER(_vp), _vp, _i],
    "gsb_dist_group_finalize": [_vp],
    "gsb_dist_group_size": [_vp],
    "gsb_dist_group_matrix": [_vp, _vp],
    "gsb_dist_group_poisson": [_vp, _i, _i],
    "gsb_dist_group_gauss_seidel": [_vp, _vp, _i, _d, _i, C.POINTER(GsOptions), _vp, C.POINTER(GsStats)],
    "gsb_dist_group_residual_l2": [_vp, _vp, _vp, _dp],
    "gsb_host_alloc": [C.POINTER(_vp), _i64],
    "gsb_alloc_counters": [C.POINTER(_i64), C.POINTER(_i64)],
    "gsb_host_free": [_vp],
}
_RESTYPE = {"gsb_last_error": C.c_char_p, "gsb_stream": C.c_void_p, "gsb_gs_default_options": None,
            "gsb_gdf_default_options": None}

_lib = None


def load():
    """Load libgsb200.so; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError("libgsb200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "or `make -C coursecomputationalphotography_b200/csrc`. There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    missing = []
    for name, args in SIGNATURES.items():
        try:
            f = getattr(L, name)
        except AttributeError:
            missing.append(name)
            continue
        f.argtypes = args
        f.restype = _RESTYPE.get(name, C.c_int)
    if missing:
        raise ImportError("libgsb200.so is stale: it lacks %s -- rebuild it (make -C %s/csrc)" %
                          (", ".join(missing), _HERE))
    _lib = L
    return L


def check(status, where):
    if status != 0:
        raise GsbError(status, where, load().gsb_last_error().decode("utf-8", "replace"))


def device_count():
    n = C.c_int(0)
    load().gsb_device_count(C.byref(n))
    return n.value


def ptr(a):
    """Raw pointer of a numpy array (or an int device pointer, passed through)."""
    if a is None:
        return None
    if isinstance(a, (int, np.integer)):
        return C.c_void_p(int(a))
    return C.c_void_p(a.ctypes.data)
