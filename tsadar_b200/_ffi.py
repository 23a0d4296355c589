"""ctypes binding of libtsff.so (include/tsff.h).  There is NO CPU fallback: importing the product on a box
without the built extension, or creating a context without an sm_100 GPU, raises."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TSFF_LIB_PATH") or os.path.join(_HERE, "_lib", "libtsff.so")   # TSFF_LIB_PATH: A/B builds of the same ABI

TSFF_ABI_VERSION = 3
TSFF_MODE_TABLE, TSFF_MODE_DIRECT, TSFF_MODE_2V = 0, 1, 2
TSFF_F32, TSFF_F64 = 0, 1
TSFF_PV_FP32, TSFF_PV_FP64 = 0, 1
P_TE, P_NE, P_LAM, P_VA, P_UD, P_NE_GRAD, P_TE_GRAD, P_AMP1, P_AMP2, P_AMP3, P_ION0 = range(11)
ION_A, ION_Z, ION_TI, ION_FRACT, ION_STRIDE = range(5)

EXPORTS = [
    "tsff_ctx_create", "tsff_ctx_destroy", "tsff_last_error", "tsff_abi_version", "tsff_ctx_set_profile_events",
    "tsff_ff_cells_bytes", "tsff_ctx_set_frozen_cells", "tsff_ff_saved_bytes", "tsff_ff_workspace_bytes", "tsff_ff_fwd", "tsff_ff_bwd", "tsff_ff_pair_fwd", "tsff_ff_pair_bwd", "tsff_chi2v_fwd",
    "tsff_pv_workspace_bytes", "tsff_pv_fwd", "tsff_pv_bwd", "tsff_microbench",
    "tsff_irf_workspace_bytes", "tsff_irf_saved_bytes", "tsff_irf_fwd", "tsff_irf_bwd", "tsff_loss_fwd_bwd",
    "tsff_ats_saved_bytes", "tsff_ats_workspace_bytes", "tsff_ats_fwd", "tsff_ats_bwd",
    "tsff_arts_weights_fwd", "tsff_arts_weights_bwd", "tsff_params_fwd", "tsff_params_bwd", "tsff_adam_step",
    "tsff_lineouts_fwd",
]


class StaticCfg(C.Structure):
    _fields_ = [
        ("abi_version", C.c_int32), ("mode", C.c_int32),
        ("W", C.c_int32), ("A", C.c_int32), ("G", C.c_int32), ("I", C.c_int32), ("V", C.c_int32),
        ("pv_precision", C.c_int32),
        ("lam_min", C.c_double), ("lam_max", C.c_double), ("lam_shift", C.c_double),
        ("v0", C.c_double), ("dv", C.c_double),
        ("sa_deg", C.POINTER(C.c_double)), ("weights", C.POINTER(C.c_double)), ("jmul", C.POINTER(C.c_double)),
        ("zp_x", C.POINTER(C.c_double)), ("zp_re", C.POINTER(C.c_double)), ("zp_im", C.POINTER(C.c_double)),
        ("zp_n", C.c_int32), ("reserved", C.c_int32),
        ("ud_angle_deg", C.c_double), ("va_angle_deg", C.c_double),
        ("W_total", C.c_int32), ("w_offset", C.c_int32),
    ]


class IrfCfg(C.Structure):
    _fields_ = [
        ("W", C.c_int32), ("nbins", C.c_int32), ("norm", C.c_int32), ("kind", C.c_int32),
        ("lam_min", C.c_double), ("lam_max", C.c_double), ("stddev", C.c_double), ("cut_sigma", C.c_double),
    ]


class AtsCfg(C.Structure):
    _fields_ = [
        ("NA", C.c_int32), ("W", C.c_int32), ("lam_step", C.c_int32), ("ang_step", C.c_int32),
        ("row_start", C.c_int32), ("row_end", C.c_int32), ("norm", C.c_int32), ("reserved", C.c_int32),
        ("ang_t0", C.c_int32), ("ang_t1", C.c_int32), ("lam_t0", C.c_int32), ("lam_t1", C.c_int32),
        ("lam_min", C.c_double), ("lam_max", C.c_double),
        ("taps_ang", C.c_void_p), ("taps_lam", C.c_void_p),
    ]


MAX_LEAVES, MAX_IONS = 24, 4


class ParamsCfg(C.Structure):
    _fields_ = [
        ("I", C.c_int32), ("V", C.c_int32), ("nm", C.c_int32), ("fe_dtype", C.c_int32), ("NLA", C.c_int32), ("reserved", C.c_int32),
        ("dv", C.c_double), ("m_offset", C.c_double), ("m0", C.c_double), ("dm", C.c_double),
        ("active_slot", C.c_int32 * MAX_LEAVES), ("scale", C.c_double * MAX_LEAVES), ("shift", C.c_double * MAX_LEAVES),
        ("ionA", C.c_double * MAX_IONS), ("ti_same", C.c_int32 * MAX_IONS), ("f_vx_m", C.c_void_p),
    ]


_lib = None


def lib():
    """Load libtsff.so (built in-tree by tsadar_b200/build.py).  Fails loudly when it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: build it with `python -m tsadar_b200.build` (nvcc, sm_100a). "
            "tsadar_b200 has no CPU or PyTorch fallback.")
    L = C.CDLL(LIB_PATH)
    vp, i64, dp = C.c_void_p, C.c_int64, C.c_void_p
    L.tsff_ctx_create.argtypes = [C.c_int, C.POINTER(StaticCfg), C.POINTER(vp)]
    L.tsff_ctx_create.restype = C.c_int
    L.tsff_ctx_destroy.argtypes = [vp]
    L.tsff_ctx_destroy.restype = None
    L.tsff_last_error.restype = C.c_char_p
    L.tsff_ctx_set_profile_events.argtypes = [vp, vp, vp, vp, vp]
    L.tsff_ctx_set_profile_events.restype = C.c_int
    L.tsff_abi_version.restype = C.c_int
    L.tsff_ff_cells_bytes.argtypes = [vp, i64]
    L.tsff_ff_cells_bytes.restype = C.c_size_t
    L.tsff_ctx_set_frozen_cells.argtypes = [vp, C.c_int, vp, i64]
    L.tsff_ctx_set_frozen_cells.restype = C.c_int
    L.tsff_ff_saved_bytes.argtypes = [vp, i64]
    L.tsff_ff_saved_bytes.restype = C.c_size_t
    L.tsff_ff_workspace_bytes.argtypes = [vp, i64]
    L.tsff_ff_workspace_bytes.restype = C.c_size_t
    L.tsff_ff_fwd.argtypes = [vp, i64, dp, vp, C.c_int, dp, dp, vp, vp, vp]
    L.tsff_ff_fwd.restype = C.c_int
    L.tsff_ff_bwd.argtypes = [vp, i64, dp, vp, C.c_int, vp, dp, dp, dp, vp, vp, vp]
    L.tsff_ff_bwd.restype = C.c_int
    L.tsff_ff_pair_fwd.argtypes = [vp, vp, i64, dp, vp, C.c_int, dp, dp, vp, vp, vp, vp]
    L.tsff_ff_pair_fwd.restype = C.c_int
    L.tsff_ff_pair_bwd.argtypes = [vp, vp, i64, dp, vp, C.c_int, vp, vp, dp, dp, dp, vp, vp, vp, vp]
    L.tsff_ff_pair_bwd.restype = C.c_int
    L.tsff_chi2v_fwd.argtypes = [vp, dp, dp, dp, dp, i64, dp, vp]
    L.tsff_chi2v_fwd.restype = C.c_int
    L.tsff_pv_workspace_bytes.argtypes = [i64, i64, i64]
    L.tsff_pv_workspace_bytes.restype = C.c_size_t
    L.tsff_pv_fwd.argtypes = [i64, i64, i64, dp, C.c_double, C.c_double, dp, dp, dp, C.c_int, vp, vp]
    L.tsff_pv_fwd.restype = C.c_int
    L.tsff_pv_bwd.argtypes = [i64, i64, i64, dp, C.c_double, C.c_double, dp, dp, dp, dp, vp, vp]
    L.tsff_pv_bwd.restype = C.c_int
    L.tsff_microbench.argtypes = [C.c_int, i64, C.POINTER(C.c_double), vp, vp]
    L.tsff_microbench.restype = C.c_int
    L.tsff_irf_workspace_bytes.argtypes = [C.POINTER(IrfCfg), i64]
    L.tsff_irf_workspace_bytes.restype = C.c_size_t
    L.tsff_irf_saved_bytes.argtypes = [C.POINTER(IrfCfg), i64]
    L.tsff_irf_saved_bytes.restype = C.c_size_t
    L.tsff_irf_fwd.argtypes = [C.POINTER(IrfCfg), i64, dp, dp, C.c_int32, dp, dp, dp, vp, vp, vp]
    L.tsff_irf_fwd.restype = C.c_int
    L.tsff_irf_bwd.argtypes = [C.POINTER(IrfCfg), i64, dp, C.c_int32, dp, vp, dp, dp, dp, vp, vp]
    L.tsff_irf_bwd.restype = C.c_int
    L.tsff_ats_saved_bytes.argtypes = [C.POINTER(AtsCfg)]
    L.tsff_ats_saved_bytes.restype = C.c_size_t
    L.tsff_ats_workspace_bytes.argtypes = [C.POINTER(AtsCfg)]
    L.tsff_ats_workspace_bytes.restype = C.c_size_t
    L.tsff_ats_fwd.argtypes = [C.POINTER(AtsCfg), dp, dp, dp, dp, dp, vp, vp, vp]
    L.tsff_ats_fwd.restype = C.c_int
    L.tsff_ats_bwd.argtypes = [C.POINTER(AtsCfg), dp, dp, vp, dp, dp, dp, vp, vp]
    L.tsff_ats_bwd.restype = C.c_int
    L.tsff_arts_weights_fwd.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, dp, C.c_int32, dp, dp, vp]
    L.tsff_arts_weights_fwd.restype = C.c_int
    L.tsff_arts_weights_bwd.argtypes = [dp, C.c_int32, C.c_int32, C.c_int32, dp, C.c_int32, dp, dp, vp]
    L.tsff_arts_weights_bwd.restype = C.c_int
    L.tsff_params_fwd.argtypes = [C.POINTER(ParamsCfg), i64, dp, dp, dp, vp, vp]
    L.tsff_params_fwd.restype = C.c_int
    L.tsff_params_bwd.argtypes = [C.POINTER(ParamsCfg), i64, dp, dp, dp, vp, dp, vp]
    L.tsff_params_bwd.restype = C.c_int
    L.tsff_adam_step.argtypes = [i64, C.c_int32, dp, dp, dp, dp, dp, C.c_double, C.c_double, C.c_double, C.c_double, vp]
    L.tsff_adam_step.restype = C.c_int
    L.tsff_lineouts_fwd.argtypes = [dp, C.c_int32, C.c_int32, vp, vp, C.c_int32, C.c_int32, C.c_double, vp, dp, dp, vp]
    L.tsff_lineouts_fwd.restype = C.c_int
    L.tsff_loss_fwd_bwd.argtypes = [i64, C.c_int32, dp, dp, dp, C.c_double, C.c_double, C.c_int, dp, dp, vp]
    L.tsff_loss_fwd_bwd.restype = C.c_int
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise RuntimeError(f"libtsff error {rc}: {lib().tsff_last_error().decode()}")
