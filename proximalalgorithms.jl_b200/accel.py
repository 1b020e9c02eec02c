"""Acceleration (direction) strategies of the line-search methods: L-BFGS and NoAcceleration.

Reference: src/accel/lbfgs.jl (LBFGS, LBFGSOperator, update!, reset!, mul!), src/accel/noaccel.jl, src/accel/traits.jl.
The operator's memory (ring of M pairs, one spare slot) and the whole two-loop recursion live in libproxb200
(csrc/qn_kernels.cu): `mul` is a chain of 2*currmem + 2 kernel launches with no host read-back, `update` is one fused pass.
"""
from __future__ import annotations

import ctypes as C

from . import _lib as L
from .host import Context, LocalComm, check_vec, pb_dtype, ptr, real_type, torch


class QuasiNewtonStyle:
    """traits.jl:7"""


class NoAccelerationStyle:
    """traits.jl:5"""


class NoAcceleration:
    """noaccel.jl:1-5: direction = -res, no state."""

    acceleration_style = NoAccelerationStyle

    def initialize(self, x):
        return None


class LBFGSOperator:
    """LBFGSOperator{M} (lbfgs.jl:5-28) over device vectors.  `currmem`, `curridx`, `H` read the library's state."""

    def __init__(self, M, x, comm=None):
        check_vec(x)
        self.ctx = Context.get(x.device)
        self.R = real_type(x.dtype)
        self.M = int(M)
        self.n = x.numel()
        self.comm = comm or LocalComm()
        if self.comm.size != 1:
            from .host import DeviceExchangeComm

            # row shards: every dot product of the two-loop chain is summed over the ranks INSIDE its kernel (NVLink exchange of the
            # reducing CTA, csrc/qn_kernels.cu), which needs the device exchange attached to this context
            if not (isinstance(self.comm, DeviceExchangeComm) and self.comm.ctx is self.ctx):
                raise L.ProxB200Error("the device L-BFGS on row shards needs the device exchange (DeviceExchangeComm) of its context")
        h = C.c_void_p()
        L.check(self.ctx.lib.pb_lbfgs_create(self.ctx.h, pb_dtype(self.R), self.n, self.M, C.byref(h)))
        self.h = h
        self._dtype = x.dtype

    def __del__(self):
        h, self.h = getattr(self, "h", None), None
        if h is not None:
            try:
                self.ctx.lib.pb_lbfgs_destroy(h)
            except Exception:
                pass

    def _info(self):
        cm, ci, hh = C.c_int(), C.c_int(), C.c_double()
        L.check(self.ctx.lib.pb_lbfgs_info(self.h, C.byref(cm), C.byref(ci), C.byref(hh)))
        return cm.value, ci.value, self.R(hh.value)

    @property
    def currmem(self):
        return self._info()[0]

    @property
    def curridx(self):
        return self._info()[1]

    @property
    def H(self):
        return self._info()[2]

    # ---- update!(L, s, y), lbfgs.jl:30-51 -----------------------------------------------------------------------------
    def enqueue_update(self, a, a_prev, b, b_prev):
        """Enqueue s = a - a_prev, y = b - b_prev (None: s = a / y = b) into the spare slot with AUX2 = <s,y>, AUX3 = <y,y>."""
        L.check(self.ctx.lib.pb_lbfgs_update(self.ctx.h, self.h, ptr(a), ptr(a_prev), ptr(b), ptr(b_prev)))

    def commit(self, sc):
        """Host half of update!: `sc` is the combined Scalars of an exchange made after `enqueue_update`."""
        acc = C.c_int()
        L.check(self.ctx.lib.pb_lbfgs_commit(self.h, float(sc.aux2), float(sc.aux3), C.byref(acc)))
        return bool(acc.value)

    def update(self, s, y):
        """Reference-shaped update!(L, s, y): one fused pass + one read-back."""
        self.enqueue_update(s, None, y, None)
        return self.commit(self.comm.exchange(self.ctx))

    def reset(self):
        """reset!, lbfgs.jl:53-56."""
        L.check(self.ctx.lib.pb_lbfgs_reset(self.h))

    # ---- mul!(d, L, v), lbfgs.jl:66-95 --------------------------------------------------------------------------------
    def mul_into(self, d, v, scale=1.0, x=None, x_d=None):
        """d = scale * (H v) by the two-loop recursion; optionally x_d = x + d in the last launch.  Asynchronous."""
        L.check(self.ctx.lib.pb_lbfgs_apply(self.ctx.h, self.h, ptr(v), float(scale), ptr(d), ptr(x), ptr(x_d)))
        return d

    def mul(self, v):
        return self.mul_into(torch().empty_like(v), v)

    __mul__ = mul
    __matmul__ = mul

    def pair(self, pos):
        """(s, y, ys) stored at ring position `pos` (1-based like the reference), as tensors viewing the ring."""
        s, y, ys = C.c_void_p(), C.c_void_p(), C.c_double()
        L.check(self.ctx.lib.pb_lbfgs_pair(self.h, int(pos), C.byref(s), C.byref(y), C.byref(ys)))
        return s.value, y.value, self.R(ys.value)


class LBFGS:
    """LBFGS(M) (lbfgs.jl:97-105)."""

    acceleration_style = QuasiNewtonStyle

    def __init__(self, M=5):
        if int(M) < 1:
            raise ValueError("LBFGS memory must be positive")
        self.M = int(M)

    def initialize(self, x, comm=None):
        return LBFGSOperator(self.M, x, comm=comm)


def acceleration_style(directions):
    """traits.jl:11 + the per-type methods: UnknownStyle is an error for PANOC (no `set_next_direction!` method)."""
    style = getattr(directions, "acceleration_style", None)
    if style is None:
        raise TypeError(f"unsupported directions {type(directions).__name__}: expected LBFGS(M) or NoAcceleration()")
    return style


__all__ = ["LBFGS", "LBFGSOperator", "NoAcceleration", "acceleration_style", "QuasiNewtonStyle", "NoAccelerationStyle"]
