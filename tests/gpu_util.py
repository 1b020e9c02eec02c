"""Helpers shared by the GPU parity tests: direct C-ABI calls on torch-owned device memory."""
import ctypes as C
import math

import numpy as np
import torch

import proxb200 as pa
from proxb200 import _lib as L
from proxb200.host import Context, ptr


def ctx():
    return Context.get()


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a)).cuda()


def dt(T):
    return L.PB_F32 if T == np.float32 else L.PB_F64


def prox_desc(kind, p0=0.0, p1=0.0, group=0, v0=None, v1=None):
    return L.pb_prox(kind, group, float(p0), float(p1), v0.data_ptr() if v0 is not None else None,
                     v1.data_ptr() if v1 is not None else None)


def fb_step(T, x, g, gamma, desc, want_y=True, want_res=True):
    c = ctx()
    z = torch.empty_like(x)
    y = torch.empty_like(x) if want_y else None
    r = torch.empty_like(x) if want_res else None
    L.check(c.lib.pb_fb_step(c.h, dt(T), x.numel(), ptr(x), ptr(g), float(gamma), C.byref(desc), ptr(y), ptr(z), ptr(r)))
    return y, z, r, c.read_scalars()


def ffb_step(T, x, g, zp, gamma, beta, desc, want_y=True, want_res=True):
    c = ctx()
    z = torch.empty_like(x)
    xn = torch.empty_like(x)
    y = torch.empty_like(x) if want_y else None
    r = torch.empty_like(x) if want_res else None
    L.check(c.lib.pb_ffb_step(c.h, dt(T), x.numel(), ptr(x), ptr(g), ptr(zp), float(gamma), float(beta), C.byref(desc),
                              ptr(y), ptr(z), ptr(r), ptr(xn)))
    return y, z, r, xn, c.read_scalars()


def pair(row, slot):
    return float(row[slot]) + float(row[slot + 1])


def fsum_sq(a):
    return math.fsum((np.asarray(a, dtype=np.float64) ** 2).tolist()) if np.asarray(a).dtype == np.float32 else fsum_prod(a, a)


def fsum_prod(a, b):
    """Exactly rounded sum of products of two float arrays (products split error-free for float64)."""
    a = np.asarray(a)
    b = np.asarray(b)
    if a.dtype == np.float32:
        return math.fsum((a.astype(np.float64) * b.astype(np.float64)).tolist())
    from fractions import Fraction

    # float64: exact rational arithmetic on a bounded number of terms
    tot = Fraction(0)
    for u, v in zip(a.tolist(), b.tolist()):
        tot += Fraction(u) * Fraction(v)
    return float(tot)


def ulps(a, b):
    if a == b:
        return 0.0
    return abs(a - b) / np.spacing(max(abs(a), abs(b)))
