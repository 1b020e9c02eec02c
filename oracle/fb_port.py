"""ctypes wrapper of the C port (oracle/c/fb_port.c).  TEST / BASELINE INFRASTRUCTURE ONLY."""
import ctypes as C
import os

import numpy as np

from . import build_port

PROX_L1, PROX_BOX = 1, 2
_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build_port.OUT
        if not os.path.exists(path):
            path = build_port.build()
        _lib = C.CDLL(path)
        for suf, ct in (("f32", C.c_float), ("f64", C.c_double)):
            f = getattr(_lib, "port_ffb_iteration_" + suf)
            f.restype = C.c_double
            f.argtypes = [C.c_int64] + [C.c_void_p] * 8 + [ct, ct, C.c_int, ct, ct, C.POINTER(C.c_double)]
            f = getattr(_lib, "port_fb_iteration_" + suf)
            f.restype = C.c_double
            f.argtypes = [C.c_int64] + [C.c_void_p] * 6 + [ct, C.c_int, ct, ct, C.POINTER(C.c_double)]
            f = getattr(_lib, "port_fill_" + suf)
            f.restype = None
            f.argtypes = [C.c_int64, C.c_void_p, C.c_uint64, ct]
        _lib.port_num_threads.restype = C.c_int
        _lib.port_set_threads.restype = None
        _lib.port_set_threads.argtypes = [C.c_int]
    return _lib


def num_threads():
    return int(lib().port_num_threads())


def set_threads(n):
    """Explicit OpenMP thread count (torchrun exports OMP_NUM_THREADS=1 to its workers)."""
    lib().port_set_threads(int(n))


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _suf(a):
    return "f32" if a.dtype == np.float32 else "f64"


def fill(v, seed, scale=1.0):
    getattr(lib(), "port_fill_" + _suf(v))(v.size, _p(v), seed, scale)


class FistaPort:
    """Holds the 8 work vectors of the unfused iteration and steps it (fixed gamma, beta supplied)."""

    def __init__(self, z, z_prev, grad, prox, p0, p1=0.0):
        self.T = z.dtype.type
        self.z, self.z_prev, self.grad = z, z_prev, grad
        n = z.size
        self.x, self.grad_f_x, self.y, self.res = (np.empty(n, z.dtype) for _ in range(4))
        self.z_new = np.empty(n, z.dtype)
        self.prox, self.p0, self.p1 = prox, p0, p1
        self.fn = getattr(lib(), "port_ffb_iteration_" + _suf(z))
        self.g_z = C.c_double()

    def step(self, gamma, beta):
        rinf = self.fn(self.z.size, _p(self.x), _p(self.z), _p(self.z_prev), _p(self.grad), _p(self.grad_f_x), _p(self.y),
                       _p(self.z_new), _p(self.res), gamma, beta, self.prox, self.p0, self.p1, C.byref(self.g_z))
        # swap(z_prev, z) then the new z: z_prev <- z, z <- z_new, recycle the old z_prev buffer
        self.z_prev, self.z, self.z_new = self.z, self.z_new, self.z_prev
        return rinf, self.g_z.value
