"""ctypes binding of ``libgparml_b200.so`` (C ABI: ``include/gparml_b200.h``).

There is deliberately no fallback: if the shared library is missing or no B200 is
visible, using the package raises.
"""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GPARML_B200_LIB") or os.path.join(HERE, "libgparml_b200.so")   # env override: kernel-variant tuning only

# return codes (include/gparml_b200.h)
OK, ERR_CUDA, ERR_ARG, ERR_NOT_PD, ERR_STATE, ERR_NO_DEVICE, ERR_RANGE = 0, -1, -2, -3, -4, -5, -6
FLAG_FP32_MAP, FLAG_FIXED_EMBEDDINGS, FLAG_FIXED_BETA = 1, 2, 4
VARIANCE_UNCONSTRAINED, VARIANCE_POSITIVE = 0, 1
MAX_PEERS = 16      # GPARML_MAX_PEERS

(A_X_MU, A_X_S, A_GRAD_D, A_GRAD_LATEST, A_GRAD_NEW, A_GRAD_OLD, A_STATS, A_KMM, A_KMM_INV, A_A_INV,
 A_DF_DKMM, A_DF_DPSI1Y, A_DF_DPSI2, A_PSI1, A_GRAD_X_MU, A_GRAD_X_S, A_Y, A_GRAD_GLOBAL, A_GS_EXTRA) = range(19)

_dp = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p
_i64 = ctypes.c_int64

STAT_FIELDS = (
    "sum_YYT", "sum_exp_K_ii", "sum_exp_K_mi_K_im", "sum_exp_K_miY", "sum_KL",
    "sum_d_exp_K_miY_d_Z", "sum_d_exp_K_mi_K_im_d_Z", "sum_d_exp_K_miY_d_alpha",
    "sum_d_exp_K_mi_K_im_d_alpha", "sum_d_exp_K_ii_d_sf2", "sum_d_exp_K_miY_d_sf2",
    "sum_d_exp_K_mi_K_im_d_sf2",
)


class NamedStats(ctypes.Structure):
    _fields_ = [(name, _dp) for name in STAT_FIELDS]


# every symbol include/gparml_b200.h declares: name -> (restype, argtypes)
PROTOTYPES = {
    "gparml_abi_version": (ctypes.c_int, []),
    "gparml_device_count": (ctypes.c_int, []),
    "gparml_last_error": (ctypes.c_char_p, []),
    "gparml_create": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, _i64, ctypes.c_int]),
    "gparml_destroy": (ctypes.c_int, [_vp]),
    "gparml_set_stream": (ctypes.c_int, [_vp, _vp]),
    "gparml_use_own_stream": (ctypes.c_int, [_vp]),
    "gparml_synchronize": (ctypes.c_int, [_vp]),
    "gparml_set_n_total": (ctypes.c_int, [_vp, _i64]),
    "gparml_upload_shard": (ctypes.c_int, [_vp, _vp, _vp, _vp, _i64, ctypes.c_int]),
    "gparml_n_local": (_i64, [_vp]),
    "gparml_jitter_events": (_i64, [_vp]),
    "gparml_set_globals": (ctypes.c_int, [_vp, _vp, ctypes.c_double, _vp, ctypes.c_double]),
    "gparml_set_step": (ctypes.c_int, [_vp, ctypes.c_double]),
    "gparml_statistics": (ctypes.c_int, [_vp]),
    "gparml_statistics_launch": (ctypes.c_int, [_vp]),
    "gparml_status": (ctypes.c_int, [_vp]),
    "gparml_stats_count": (_i64, [_vp]),
    "gparml_stats_device_ptr": (ctypes.c_int, [_vp, ctypes.POINTER(_vp)]),
    "gparml_stats_add": (ctypes.c_int, [_vp, _vp, ctypes.c_double]),
    "gparml_stats_copy": (ctypes.c_int, [_vp, _vp]),
    "gparml_stats_expand": (ctypes.c_int, [_vp, ctypes.POINTER(NamedStats)]),
    "gparml_stats_set_named": (ctypes.c_int, [_vp, ctypes.POINTER(NamedStats)]),
    "gparml_global_step": (ctypes.c_int, [_vp, _vp, _vp]),
    "gparml_global_step_begin": (ctypes.c_int, [_vp]),
    "gparml_global_step_end": (ctypes.c_int, [_vp, _vp, _vp]),
    "gparml_update_global_statistics": (ctypes.c_int, [_vp]),
    "gparml_embedding_grads": (ctypes.c_int, [_vp]),
    "gparml_embedding_grads_download": (ctypes.c_int, [_vp, _vp, ctypes.c_int]),
    "gparml_kmm_derivative": (ctypes.c_int, [_vp, ctypes.c_int, _vp]),
    "gparml_grad_contract": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _vp, _vp, _vp, _vp, _vp, _vp]),
    "gparml_stats_add_peer": (ctypes.c_int, [_vp, _vp, ctypes.c_double]),
    "gparml_stats_copy_peer": (ctypes.c_int, [_vp, _vp]),
    "gparml_stats_allreduce_peers": (ctypes.c_int, [ctypes.POINTER(_vp), ctypes.c_int, ctypes.c_double]),
    "gparml_array_count": (_i64, [_vp, ctypes.c_int]),
    "gparml_download": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _i64]),
    "gparml_upload": (ctypes.c_int, [_vp, ctypes.c_int, _vp, _i64]),
    "gparml_array_device_ptr": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.POINTER(_vp)]),
    "gparml_scg_set_grads": (ctypes.c_int, [_vp]),
    "gparml_scg_get_mu": (ctypes.c_int, [_vp, _dp]),
    "gparml_scg_get_kappa": (ctypes.c_int, [_vp, _dp]),
    "gparml_scg_get_theta": (ctypes.c_int, [_vp, _dp]),
    "gparml_scg_get_current_grad": (ctypes.c_int, [_vp, _dp]),
    "gparml_scg_get_gamma": (ctypes.c_int, [_vp, _dp]),
    "gparml_scg_get_max_d": (ctypes.c_int, [_vp, ctypes.c_double, _dp]),
    "gparml_scg_reset_d": (ctypes.c_int, [_vp]),
    "gparml_scg_update_d": (ctypes.c_int, [_vp, ctypes.c_double]),
    "gparml_scg_update_X": (ctypes.c_int, [_vp, ctypes.c_double]),
    "gparml_scg_update_grad_old": (ctypes.c_int, [_vp]),
    "gparml_scg_update_grad_new": (ctypes.c_int, [_vp]),
    "gparml_launch_count": (_i64, [_vp]),
    "gparml_enable_timing": (ctypes.c_int, [_vp, ctypes.c_int]),
    "gparml_phase_times": (ctypes.c_int, [_vp, _dp]),
    "gparml_measure_dfma_peak": (ctypes.c_int, [_vp, _dp]),
    "gparml_upload_outputs": (ctypes.c_int, [_vp, _vp, _i64]),
    "gparml_init_column_sums": (ctypes.c_int, [_vp, _vp]),
    "gparml_init_scatter": (ctypes.c_int, [_vp, _vp, _vp]),
    "gparml_init_project": (ctypes.c_int, [_vp, _vp, _vp]),
    "gparml_init_random": (ctypes.c_int, [_vp, ctypes.c_int, ctypes.c_uint64, _i64]),
    "gparml_kmeans_step": (ctypes.c_int, [_vp, _vp, ctypes.c_int, _vp]),
}

_lib = None


class GparmlError(RuntimeError):
    pass


def load():
    """Load the shared library (raises if it has not been built)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise GparmlError("%s not found: build it with `python -m gparml_b200.build` "
                              "(there is no CPU fallback)" % LIB_PATH)
        lib = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(lib, name)          # AttributeError if the symbol is missing
            fn.restype = res
            fn.argtypes = args
        _lib = lib
    return _lib


def last_error():
    return load().gparml_last_error().decode("utf-8", "replace")


def check(code):
    """Map C return codes onto the exception types the reference's optimiser wrapper
    survives (scg_adapted.py:55: LinAlgError, ZeroDivisionError, ValueError, Warning,
    AssertionError)."""
    if code == OK:
        return
    msg = last_error()
    if code == ERR_NOT_PD:
        raise np.linalg.LinAlgError(msg)
    if code == ERR_RANGE:
        raise AssertionError(msg)
    if code in (ERR_ARG, ERR_STATE):
        raise ValueError(msg)
    raise GparmlError("gparml_b200 error %d: %s" % (code, msg))


def as_f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and tuple(a.shape) != tuple(shape):
        raise ValueError("expected array of shape %r, got %r" % (tuple(shape), a.shape))
    return a


def ptr(a):
    return ctypes.c_void_p(a.ctypes.data)
