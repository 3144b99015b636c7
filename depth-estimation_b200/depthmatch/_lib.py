"""ctypes binding of libdepthmatch.so (include/depthmatch.h).

The library is the product; this module only declares its signatures.  If the
shared object is missing it is built with nvcc (build.py); if there is no CUDA
device dm_create fails and DepthMatchError is raised -- there is no CPU path.
"""
import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(HERE)
SO_PATH = os.path.join(PKG, "csrc", "libdepthmatch.so")

DM_OK, DM_ERR_INVALID, DM_ERR_CUDA, DM_ERR_UNSUPPORTED, DM_ERR_NOMEM = 0, -1, -2, -3, -4
DM_VOLUME_SSD, DM_VOLUME_NEG_SOFTMAX, DM_VOLUME_EXACT = 0, 1, 0x100
DM_FLAG_TIE_MIDDLE, DM_FLAG_EXACT_SSD, DM_FLAG_ASYNC, DM_FLAG_DIFF_SSD = 1, 2, 4, 8


class DepthMatchError(RuntimeError):
    def __init__(self, status, message):
        super().__init__("libdepthmatch error %d: %s" % (status, message))
        self.status = status


class dm_pair(C.Structure):
    _fields_ = [("in1", C.c_void_p), ("in2", C.c_void_p),
                ("n_pairs", C.c_int32), ("channels", C.c_int32),
                ("h1", C.c_int32), ("w1", C.c_int32), ("h2", C.c_int32), ("w2", C.c_int32),
                ("in1_stride_n", C.c_int64), ("in1_stride_c", C.c_int64), ("in1_stride_y", C.c_int64),
                ("in2_stride_n", C.c_int64), ("in2_stride_c", C.c_int64), ("in2_stride_y", C.c_int64)]


class dm_extract_out(C.Structure):
    _fields_ = [("index", C.c_void_p), ("min_ssd", C.c_void_p), ("pmax", C.c_void_p),
                ("flow_full", C.c_void_p), ("index_thr", C.c_void_p), ("score_thr", C.c_void_p),
                ("soft_yx", C.c_void_p), ("n_untouched", C.c_void_p), ("conf_marginal", C.c_void_p)]


class dm_conv_layer(C.Structure):
    _fields_ = [("n_in", C.c_int32), ("n_out", C.c_int32), ("kh", C.c_int32), ("kw", C.c_int32),
                ("n_conn", C.c_int32), ("tanh_after", C.c_int32),
                ("conn", C.c_void_p), ("weight", C.c_void_p), ("bias", C.c_void_p)]


# name -> (restype, argtypes); mirrors include/depthmatch.h declaration by declaration
_P, _I, _D, _L, _F = C.c_void_p, C.c_int, C.c_double, C.c_int64, C.c_float
SIGNATURES = {
    "dm_version": (_I, []),
    "dm_last_error": (C.c_char_p, []),
    "dm_create": (_I, [_I, C.POINTER(_P)]),
    "dm_destroy": (_I, [_P]),
    "dm_synchronize": (_I, [_P]),
    "dm_set_stream": (_I, [_P, _P]),
    "dm_reset_stream": (_I, [_P]),
    "dm_get_stream": (_P, [_P]),
    "dm_host_alloc": (_I, [C.POINTER(_P), C.c_size_t]),
    "dm_host_free": (_I, [_P]),
    "dm_launch_count": (_L, [_P]),
    "dm_set_profiling": (_I, [_P, _I]),
    "dm_last_kernel_ms": (_I, [_P, C.POINTER(_F)]),
    "dm_set_option": (_I, [_P, C.c_char_p, C.c_char_p]),
    "dm_last_counts": (_I, [_P, C.POINTER(_L), C.POINTER(_L)]),
    "dm_match_volume": (_I, [_P, C.POINTER(dm_pair), _I, _I, _I, _P]),
    "dm_match_extract": (_I, [_P, C.POINTER(dm_pair), _I, _I, C.c_uint, _D, _I, _I,
                              C.POINTER(dm_extract_out)]),
    "dm_radial_match_extract": (_I, [_P, C.POINTER(dm_pair), _I, _P, _P]),
    "dm_neg_softmax": (_I, [_P, _P, _L, _I, _P]),
    "dm_argmax_tie": (_I, [_P, _P, _L, _I, _I, _I, _P, _P]),
    "dm_extract_output": (_I, [_P, _P, _I, _I, _I, _D, _P, _P, C.POINTER(_L)]),
    "dm_match_extract_raw_ssd": (_I, [_P, _P, _I, _I, _D, _P, _P, _P]),
    "dm_extract_output_marginalized": (_I, [_P, _P, _I, _I, _I, _D, _D, _P, _P]),
    "dm_soft_mean": (_I, [_P, _P, _L, _I, _I, _P, _P]),
    "dm_marginal_x": (_I, [_P, _P, _L, _I, _I, _P]),
    "dm_flow_canvas": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _P]),
    "dm_x2yx_multi": (_I, [_P, _P, _I, _I, _I, _I, C.POINTER(_I), _I, _I, _P, _P]),
    "dm_cascade_add": (_I, [_P, _P, _L, _I, _I, C.POINTER(_I), _I, _P]),
    "dm_multiscale_extract": (_I, [_P, C.POINTER(_P), C.POINTER(_P), _I, _I, _I, _I, _I,
                                   C.POINTER(_I), _I, _P, _P, _P]),
    "dm_downsample_avg": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "dm_polar_remap": (_I, [_P, _P, _I, _I, _I, _D, _D, _D, _D, _I, _I, _P, _I, _I]),
    "dm_polar_unmap": (_I, [_P, _P, _I, _I, _I, _D, _D, _D, _D, _P, _I, _I]),
    "dm_warp_bilinear": (_I, [_P, _P, _I, _I, _I, _P, _I, _I, _P]),
    "dm_c2p_mask": (_I, [_P, _I, _I, _D, _D, _I, _I, _D, _D, _P]),
    "dm_p2c_mask": (_I, [_P, _I, _I, _I, _I, _D, _D, _D, _D, _P]),
    "dm_flow2depth": (_I, [_P, _P, _I, _I, _F, _F, _F, _P, _P]),
    "dm_warp_homography": (_I, [_P, _P, _I, _I, _I, C.POINTER(C.c_double), _I, _I, _P, _P]),
    "dm_post_process_image": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "dm_enlarge_mask": (_I, [_P, _P, _I, _I, _I, _I]),
    "dm_radial_depth": (_I, [_P, _P, _I, _I, _F, _F, _F, _P, _P]),
    "dm_depth_from_xflow": (_I, [_P, _P, _P, _I, _I, _F, _P, _P]),
    "dm_filter_create": (_I, [_P, C.POINTER(dm_conv_layer), _I, C.POINTER(_P)]),
    "dm_filter_destroy": (_I, [_P]),
    "dm_filter_output_size": (_I, [_P, _I, _I, _I, _I, _I, _I, C.POINTER(_I), C.POINTER(_I), C.POINTER(_I)]),
    "dm_filter_forward": (_I, [_P, _P, _P, _I, _I, _I, _I, _I, _I, _I, _P]),
}

_lib = None


def load(build_if_missing=True):
    """Load libdepthmatch.so; build it first when it is not there."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        if not build_if_missing:
            raise DepthMatchError(DM_ERR_UNSUPPORTED, "libdepthmatch.so is not built: run "
                                  "python depth-estimation_b200/build.py")
        import importlib.util
        spec = importlib.util.spec_from_file_location("_dm_build", os.path.join(PKG, "build.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        mod.build()
    lib = C.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(status):
    if status != DM_OK:
        raise DepthMatchError(status, load().dm_last_error().decode("utf-8", "replace"))
