"""ctypes binding of libproxb200.so (declared in include/proxb200.h).

The library is the product: there is no Python / torch / CPU fallback for any entry point.  If the shared object is
missing (not built) the import of this module raises immediately with the build command.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "lib", "libproxb200.so")

PB_F32, PB_F64 = 0, 1
PB_PROX_ZERO, PB_PROX_L1, PB_PROX_BOX, PB_PROX_SCALE, PB_PROX_L21, PB_PROX_SQRL2, PB_PROX_BALL = 0, 1, 2, 3, 4, 5, 6
PB_OPT_CTAS_PER_SM, PB_OPT_STREAM_HINTS, PB_OPT_UNROLL, PB_OPT_STEP_IMPL, PB_OPT_FUSED_EXCHANGE, PB_OPT_PERSISTENT, PB_OPT_MULTI_ITER, PB_OPT_GEMV_SCALAR, PB_OPT_LSQ_FUSED, PB_OPT_LSQ_FISTA = 0, 1, 2, 3, 4, 5, 6, 7, 8, 9
PB_IPC_HANDLE_BYTES, PB_MAX_WORLD = 64, 8
PB_S_GSUM, PB_S_RESSQ, PB_S_GDR, PB_S_RESINF, PB_S_AUX, PB_S_AUXINF, PB_S_AUX2, PB_S_AUX3, PB_NSCALARS = 0, 2, 4, 6, 8, 10, 12, 14, 16


class ProxB200Error(RuntimeError):
    """Raised when a libproxb200 call returns a non-zero status (message from pb_last_error)."""


class pb_prox(C.Structure):
    _fields_ = [
        ("kind", C.c_int32),
        ("group", C.c_int32),
        ("p0", C.c_double),
        ("p1", C.c_double),
        ("v0", C.c_void_p),
        ("v1", C.c_void_p),
    ]


class pb_smooth(C.Structure):
    _fields_ = [("kind", C.c_int32), ("pad", C.c_int32), ("m", C.c_int64), ("n", C.c_int64), ("lda", C.c_int64),
                ("nblk", C.c_int64), ("mb", C.c_int64), ("nb", C.c_int64), ("A", C.c_void_p), ("b", C.c_void_p), ("r", C.c_void_p)]


class pb_solve_opts(C.Structure):
    _fields_ = [("algorithm", C.c_int32), ("adaptive", C.c_int32), ("sequence", C.c_int32), ("profile", C.c_int32),
                ("maxit", C.c_int64), ("n_global", C.c_int64), ("tol", C.c_double), ("gamma", C.c_double), ("mf", C.c_double),
                ("constant_beta", C.c_double), ("minimum_gamma", C.c_double), ("reduce_gamma", C.c_double), ("increase_gamma", C.c_double),
                ("spare_x", C.c_void_p), ("spare_z", C.c_void_p), ("spare_grad", C.c_void_p)]


class pb_solve_result(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("backtracks", C.c_int64), ("gamma", C.c_double), ("f_x", C.c_double), ("g_z", C.c_double),
                ("res_inf", C.c_double), ("warned_small_gamma", C.c_int32), ("persistent_ctas", C.c_int32), ("x", C.c_void_p), ("grad", C.c_void_p),
                ("z", C.c_void_p), ("z_prev", C.c_void_p), ("loop_ms", C.c_double), ("step_kernel_ms", C.c_double),
                ("step_kernel_launches", C.c_int64), ("res_sq", C.c_double), ("gdr", C.c_double), ("gsum", C.c_double),
                ("multi_iter_kernel", C.c_int32), ("pad", C.c_int32)]


class pb_panoc_opts(C.Structure):
    _fields_ = [("maxit", C.c_int64), ("tol", C.c_double), ("alpha", C.c_double), ("beta", C.c_double), ("gamma", C.c_double),
                ("minimum_gamma", C.c_double), ("adaptive", C.c_int32), ("max_backtracks", C.c_int32), ("lbfgs_mem", C.c_int32),
                ("quadratic", C.c_int32), ("Am", C.c_int64), ("An", C.c_int64), ("A", C.c_void_p)]


class pb_panoc_result(C.Structure):
    _fields_ = [("iterations", C.c_int64), ("gamma_backtracks", C.c_int64), ("tau_backtracks", C.c_int64), ("gamma", C.c_double),
                ("f_Ax", C.c_double), ("g_z", C.c_double), ("res_inf", C.c_double), ("tau", C.c_double), ("warned_small_gamma", C.c_int32),
                ("pad", C.c_int32)]


PB_F_LSQ_DENSE, PB_F_LSQ_BLOCKDIAG, PB_F_SQDIST, PB_F_LINEAR = 0, 1, 2, 3
PB_ALG_FB, PB_ALG_FFB = 0, 1
PB_SEQ_ADAPTIVE, PB_SEQ_FIXED, PB_SEQ_SIMPLE, PB_SEQ_CONSTANT = 0, 1, 2, 3

_vp, _i, _i64, _d, _sz = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
_pp = C.POINTER(pb_prox)

# name -> (restype, argtypes); mirrors include/proxb200.h one to one (tests/test_abi.py checks the header against this table)
SIGNATURES = {
    "pb_version": (C.c_char_p, []),
    "pb_last_error": (C.c_char_p, []),
    "pb_device_count": (_i, [C.POINTER(C.c_int)]),
    "pb_ctx_create": (_i, [_i, _vp, _i, C.POINTER(_vp)]),
    "pb_ctx_destroy": (_i, [_vp]),
    "pb_ctx_make_current": (_i, [_vp]),
    "pb_ctx_stream": (_vp, [_vp]),
    "pb_ctx_sync": (_i, [_vp]),
    "pb_ctx_scalars_dev": (_vp, [_vp]),
    "pb_ctx_set_scalars_dev": (_i, [_vp, _vp]),
    "pb_ctx_set_option": (_i, [_vp, _i, _i]),
    "pb_ctx_launch_count": (_i64, [_vp]),
    "pb_xchg_init": (_i, [_vp, _i, _i, _vp]),
    "pb_xchg_connect": (_i, [_vp, _vp]),
    "pb_xchg_connect_local": (_i, [C.POINTER(_vp), _i]),
    "pb_xchg_shutdown": (_i, [_vp]),
    "pb_exchange": (_i, [_vp]),
    "pb_exchange_wait": (_i, [_vp, C.POINTER(C.c_double), _d]),
    "pb_malloc": (_i, [_vp, _sz, C.POINTER(_vp)]),
    "pb_free": (_i, [_vp, _vp]),
    "pb_host_alloc": (_i, [_sz, C.POINTER(_vp)]),
    "pb_host_free": (_i, [_vp]),
    "pb_upload": (_i, [_vp, _vp, _vp, _sz]),
    "pb_download": (_i, [_vp, _vp, _vp, _sz]),
    "pb_copy": (_i, [_vp, _vp, _vp, _sz]),
    "pb_memset_zero": (_i, [_vp, _vp, _sz]),
    "pb_fill_counter": (_i, [_vp, _i, _i64, _i64, C.c_uint64, _d, _vp]),
    "pb_read_scalars": (_i, [_vp, C.POINTER(C.c_double)]),
    "pb_fb_step": (_i, [_vp, _i, _i64, _vp, _vp, _d, _pp, _vp, _vp, _vp]),
    "pb_ffb_step": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _d, _d, _pp, _vp, _vp, _vp, _vp]),
    "pb_prox_apply": (_i, [_vp, _i, _i64, _vp, _d, _pp, _vp]),
    "pb_forward": (_i, [_vp, _i, _i64, _vp, _vp, _d, _vp]),
    "pb_extrapolate": (_i, [_vp, _i, _i64, _vp, _vp, _d, _vp]),
    "pb_residual": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _vp]),
    "pb_add_scalar": (_i, [_vp, _i, _i64, _vp, _d, _vp]),
    "pb_sub": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "pb_nrm2sq": (_i, [_vp, _i, _i64, _vp]),
    "pb_dot": (_i, [_vp, _i, _i64, _vp, _vp]),
    "pb_lsq_dense_residual": (_i, [_vp, _i, _i64, _i64, _vp, _i64, _vp, _vp, _vp]),
    "pb_lsq_dense_gradient": (_i, [_vp, _i, _i64, _i64, _vp, _i64, _vp, _vp]),
    "pb_lsq_dense_residual_sharded": (_i, [_vp, _i, _i64, _i64, _vp, _i64, _vp, _vp, _vp, _i64, _i64, _i]),
    "pb_lsq_dense_chunk_cols": (_i64, [_i, _i64, _i64]),
    "pb_lsq_blockdiag_residual": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _vp, _vp, _vp]),
    "pb_lsq_blockdiag_gradient": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _vp, _vp]),
    "pb_lsq_blockdiag_value_and_gradient": (_i, [_vp, _i, _i64, _i64, _i64, _vp, _vp, _vp, _vp, _vp]),
    "pb_sqdist": (_i, [_vp, _i, _i64, _vp, _vp, _vp]),
    "pb_lincomb2": (_i, [_vp, _i, _i64, _d, _vp, _d, _vp, _vp]),
    "pb_scale": (_i, [_vp, _i, _i64, _d, _vp, _vp]),
    "pb_lbfgs_create": (_i, [_vp, _i, _i64, _i, C.POINTER(_vp)]),
    "pb_lbfgs_destroy": (_i, [_vp]),
    "pb_lbfgs_reset": (_i, [_vp]),
    "pb_lbfgs_info": (_i, [_vp, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_double)]),
    "pb_lbfgs_pair": (_i, [_vp, _i, C.POINTER(_vp), C.POINTER(_vp), C.POINTER(C.c_double)]),
    "pb_lbfgs_update": (_i, [_vp, _vp, _vp, _vp, _vp, _vp]),
    "pb_lbfgs_commit": (_i, [_vp, _d, _d, C.POINTER(C.c_int)]),
    "pb_lbfgs_apply": (_i, [_vp, _vp, _vp, _d, _vp, _vp, _vp]),
    "pb_conj_prox": (_i, [_vp, _i, _i64, _vp, _d, _pp, _vp]),
    "pb_dr_tv_step": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _d, _d, _vp, _vp, _vp, _i64, _i64, _vp, _vp]),
    "pb_ipc_export": (_i, [_vp, _vp, _vp]),
    "pb_ipc_open": (_i, [_vp, _vp, C.POINTER(_vp)]),
    "pb_ipc_close": (_i, [_vp, _vp]),
    "pb_fd2d_forward": (_i, [_vp, _i, _i64, _i64, _vp, _vp]),
    "pb_fd2d_adjoint": (_i, [_vp, _i, _i64, _i64, _vp, _vp]),
    "pb_lsq_prox_create": (_i, [_vp, _i, _i64, _i64, _vp, _vp, _d, C.POINTER(_vp)]),
    "pb_lsq_prox_destroy": (_i, [_vp]),
    "pb_lsq_prox_apply": (_i, [_vp, _vp, _vp, _d, _vp]),
    "pb_dr_step": (_i, [_vp, _i, _i64, _vp, _d, _pp, _pp, _vp, _vp, _vp, _vp, _vp]),
    "pb_solve": (_i, [_vp, _i, _i64, C.POINTER(pb_smooth), _pp, C.POINTER(pb_solve_opts), _vp, _vp, _vp, _vp, _vp, _vp, _vp,
                      C.POINTER(pb_solve_result)]),
    "pb_panoc_solve": (_i, [_vp, _i, _i64, C.POINTER(pb_smooth), _pp, C.POINTER(pb_panoc_opts), _vp, _vp, C.POINTER(pb_panoc_result)]),
    "pb_persist_phase_cycles": (_i, [_vp, C.POINTER(C.c_int64)]),
    "pb_ffb_step_host": (_i, [_vp, _i, _i64, _vp, _vp, _vp, _d, _d, _pp, _vp, _vp, C.POINTER(C.c_double)]),
}


def load(path: str = LIB_PATH) -> C.CDLL:
    if not os.path.exists(path):
        raise ProxB200Error(
            f"{path} not found: the CUDA library is the product and has no fallback. "
            "Build it with `python proximalalgorithms.jl_b200/build.py` (needs nvcc)."
        )
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _lib = load()
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise ProxB200Error(f"libproxb200 error {rc}: {lib().pb_last_error().decode(errors='replace')}")
