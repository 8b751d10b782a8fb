"""ctypes binding of liblkflow_b200.so (C ABI declared in include/lkflow_b200.h).

There is NO fallback: if the shared library is missing or a CUDA device is absent the product path
raises.  (`python __graft_entry__.py build` / `csrc/build.py` compiles it in-tree with nvcc.)
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# LKFLOW_B200_SO: developer switch for A/B timing of two builds of the same library (tools/exp_round.py)
SO_PATH = os.environ.get("LKFLOW_B200_SO") or os.path.join(_HERE, "csrc", "liblkflow_b200.so")
ABI_VERSION = 2          # LKF_ABI_VERSION of include/lkflow_b200.h

LKF_MAX_LEVELS = 8


class LkfError(RuntimeError):
    """Raised when a C-ABI call returns a negative status (the analogue of cv2.error)."""

    def __init__(self, status, where, detail=""):
        self.status = status
        super().__init__(f"{where}: {detail}" if detail else where)


class LkfImage(ctypes.Structure):
    _fields_ = [("data", ctypes.c_void_p), ("batch", ctypes.c_int32), ("height", ctypes.c_int32),
                ("width", ctypes.c_int32), ("channels", ctypes.c_int32), ("row_pitch", ctypes.c_int64),
                ("image_pitch", ctypes.c_int64)]


class LkfLkParams(ctypes.Structure):
    _fields_ = [("win_w", ctypes.c_int32), ("win_h", ctypes.c_int32), ("max_level", ctypes.c_int32),
                ("crit_type", ctypes.c_int32), ("max_count", ctypes.c_int32), ("flags", ctypes.c_int32),
                ("epsilon", ctypes.c_double), ("min_eig_threshold", ctypes.c_double)]


class LkfGfttParams(ctypes.Structure):
    _fields_ = [("max_corners", ctypes.c_int32), ("block_size", ctypes.c_int32),
                ("dy_tail_cols", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("quality_level", ctypes.c_double), ("min_distance", ctypes.c_double)]


LKF_VP_RING = 32


class LkfVpParams(ctypes.Structure):
    _fields_ = [("cp_gate_x", ctypes.c_double), ("cp_gate_y", ctypes.c_double), ("min_ang_dif", ctypes.c_double),
                ("max_cp_std", ctypes.c_double), ("vp_update_rate", ctypes.c_double),
                ("fl_update_rate", ctypes.c_double), ("min_fl_len", ctypes.c_double),
                ("tp_refresh_below", ctypes.c_double), ("start_x_sep", ctypes.c_float),
                ("vp_ref_num", ctypes.c_int32), ("hide_vp_thold", ctypes.c_int32),
                ("reset_avg_on_hide", ctypes.c_int32), ("tp_update_time", ctypes.c_int32),
                ("vl_ref", ctypes.c_int32)]


class LkfVpState(ctypes.Structure):
    _fields_ = [("vp_x", ctypes.c_double), ("vp_y", ctypes.c_double),
                ("ring_x", ctypes.c_double * LKF_VP_RING), ("ring_y", ctypes.c_double * LKF_VP_RING),
                ("avg_len", ctypes.c_double * 2), ("ring_alias", ctypes.c_int32 * LKF_VP_RING),
                ("vp_init", ctypes.c_int32), ("ring_count", ctypes.c_int32), ("ring_head", ctypes.c_int32),
                ("vp_ult", ctypes.c_int32), ("tp_ult", ctypes.c_int32), ("n_rows", ctypes.c_int32),
                ("row_overflow", ctypes.c_int32), ("refresh", ctypes.c_int32),
                ("n_lines", ctypes.c_int64), ("n_cps", ctypes.c_int64),
                ("vp_moved", ctypes.c_int32), ("vl_init", ctypes.c_int32),
                ("vl_m", ctypes.c_double), ("vl_b", ctypes.c_double), ("vl_mv", ctypes.c_double), ("vl_bv", ctypes.c_double),
                ("vl_bp", ctypes.c_double * 2), ("vl_lp_y", ctypes.c_double), ("vl_rp_y", ctypes.c_double),
                ("vl_up_x", ctypes.c_double), ("vl_dp_x", ctypes.c_double)]


_P = ctypes.POINTER
_vp = ctypes.c_void_p
_i = ctypes.c_int

# name -> (restype, argtypes); every symbol include/lkflow_b200.h declares must be listed here
SIGNATURES = {
    "lkf_abi_version": (_i, []),
    "lkf_struct_size": (_i, [_i]),
    "lkf_strerror": (ctypes.c_char_p, [_i]),
    "lkf_last_error": (ctypes.c_char_p, []),
    "lkf_ctx_create": (_i, [_i, _P(_vp)]),
    "lkf_ctx_destroy": (None, [_vp]),
    "lkf_ctx_launch_count": (ctypes.c_int64, [_vp]),
    "lkf_bgr2gray_u8": (_i, [_P(LkfImage), _P(LkfImage), _vp]),
    "lkf_gaussian_blur3_u8": (_i, [_P(LkfImage), _P(LkfImage), _vp]),
    "lkf_process_img_u8": (_i, [_P(LkfImage), _P(LkfImage), _vp]),
    "lkf_pyr_down_u8": (_i, [_P(LkfImage), _P(LkfImage), _vp]),
    "lkf_process_pyramid_u8": (_i, [_P(LkfImage), _P(LkfImage), ctypes.c_int, _vp]),
    "lkf_scharr_s16": (_i, [_P(LkfImage), _P(LkfImage), _vp]),
    "lkf_effective_max_level": (_i, [_i, _i, _i, _i, _i]),
    "lkf_pyrlk_sparse": (_i, [_vp, _P(LkfLkParams), _i, _P(LkfImage), _P(LkfImage), _vp, _vp, _i,
                              _vp, _vp, _vp, _vp, _vp]),
    "lkf_pyrlk_dense": (_i, [_vp, _P(LkfLkParams), _i, _P(LkfImage), _P(LkfImage), _vp, _vp, _vp, _vp, _vp]),
    "lkf_corner_min_eigen_f32": (_i, [_P(LkfImage), _P(LkfImage), _i, _i, _vp]),
    "lkf_probe_fp32_fma": (_i, [_i, _i, _vp, _vp]),
    "lkf_probe_pipe": (_i, [_i, _i, _i, _vp, _vp]),
    "lkf_gftt_detect": (_i, [_vp, _P(LkfGfttParams), _P(LkfImage), _P(LkfImage), _i, _vp, _vp, _vp, _i, _vp]),
    "lkf_resize_area_u8": (_i, [_P(LkfImage), _P(LkfImage), _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "lkf_vp_frame": (_i, [_P(LkfVpParams), _vp, _i, _vp, _vp, _vp, _vp, _i, _P(LkfImage), _vp, _i, _vp, _vp, _vp]),
    "lkf_vp_apply_refresh": (_i, [_vp, _i, _vp, _vp, _i, _vp, _vp, _i, _i, _vp]),
    "lkf_gftt_select": (_i, [_vp, _P(LkfGfttParams), _P(LkfImage), _P(LkfImage), _i, _vp, _vp, _i, _vp]),
}

_lib = None


def load():
    """dlopen the library and type every entry point.  Raises ImportError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for this path.")
    lib = ctypes.CDLL(SO_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)      # AttributeError if the .so lacks a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.lkf_abi_version() != ABI_VERSION:
        raise ImportError("liblkflow_b200.so ABI version mismatch; rebuild")
    for which, cls in enumerate((LkfImage, LkfLkParams, LkfGfttParams, LkfVpParams, LkfVpState)):
        if lib.lkf_struct_size(which) != ctypes.sizeof(cls):
            raise ImportError(f"{cls.__name__}: ctypes mirror is {ctypes.sizeof(cls)} bytes, the library's struct "
                              f"{lib.lkf_struct_size(which)}; rebuild")
    _lib = lib
    return lib


def check(status, where):
    if status < 0:
        lib = load()
        msg = lib.lkf_strerror(status).decode()
        if status in (-2, -3) and lib.lkf_last_error():
            msg += ": " + lib.lkf_last_error().decode()
        raise LkfError(status, where, msg)
    return status
