"""Douglas-Rachford splitting on the device.

Reference: src/algorithms/douglas_rachford.jl:31-42 (parameters), :46-52 (state), :54-63 (the one `iterate` method),
:65-72 (stop rule, solution = state.y, display), :100-119 (constructor).  Names, defaults and semantics are the same.

When both proximable terms are element-wise kinds of the library (Zero, NormL1, IndBox, SqrNormL2 with or without a
translation) the whole iteration -- prox_f, reflection, prox_g, residual, x update and the stop norm -- is ONE kernel pass
(`pb_dr_step`, csrc/dr_kernels.cu): read x, write x.  `state.y` (the solution), `r`, `z`, `res` are materialised on demand
from the pre-update x, which is kept by ping-ponging two x buffers.  Any other pair (user `prox_` callbacks, NormL21,
IndBallL2, LeastSquares) runs the reference's five-operation sequence with the library's kernels.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L
from .algorithms import IterativeAlgorithm, _Engine, _to_device_copy
from .functions import Zero
from .host import pb_dtype, ptr, real_type, torch
from .tv import IndConsensus, TVDouglasRachfordEngine

_ELEMENTWISE = (L.PB_PROX_ZERO, L.PB_PROX_L1, L.PB_PROX_BOX, L.PB_PROX_SQRL2)


def _elementwise(g):
    return getattr(g, "fused", False) and getattr(g, "kind", None) in _ELEMENTWISE


class DouglasRachfordState:
    """douglas_rachford.jl:46-52: x, y, r, z, res."""

    def __init__(self):
        self._mat = None          # (y, r, z, res) buffers of the fused path, filled on demand
        self._mat_valid = False

    def release(self):
        """Called by the solver driver once the solution has been copied out: a row-sharded TV run owns pb_malloc'ed ping-pong buffers
        and cudaIpc mappings of its neighbours (outside torch's allocator).  `state.x` is a view of those buffers, so it is copied
        first; y / z were materialised into ordinary tensors.  Collective in the sense that every rank does it after its last step."""
        eng = getattr(self, "_tv", None)
        if eng is not None and getattr(eng, "bufs", None):
            self._materialise()                  # y, z of the final state (needs the buffers and the halo mappings)
            self.x = self.x.clone()
            if eng.sharded:
                torch().cuda.synchronize(self.x.device)
                self._it.f.comm.dist.barrier(group=getattr(self._it.f.comm, "group", None))   # nobody reads my rows any more
            eng.close()

    def _materialise(self):
        if self._tv is not None:
            return self._materialise_tv()
        if self._fused and not self._mat_valid:
            t = torch()
            if self._mat is None:
                self._mat = tuple(t.empty_like(self.x) for _ in range(4))
            it, e = self._it, self._engine
            y, r, z, res = self._mat
            # re-run the pass on the pre-update x: same arithmetic, same x (written to the current x again)
            L.check(e.lib.pb_dr_step(e.ctx.h, pb_dtype(it.R), self.x.numel(), ptr(self._x_in), float(it.gamma), C.byref(it._fd),
                                     C.byref(it._gd), ptr(self.x), ptr(y), ptr(r), ptr(z), ptr(res)))
            self._mat_valid = True
        return self._mat

    def _materialise_tv(self):
        """TV consensus form: y (five stacked copies) and the consensus image z are written by re-running the last pass.
        Row-sharded runs: a COLLECTIVE operation (every rank must ask at the same iteration), fenced by two barriers so that
        no neighbour overwrites the buffer the halo rows are read from."""
        if not self._mat_valid:
            t = torch()
            eng, it = self._tv, self._it
            if self._mat is None:
                n = it.f.H * it.f.W
                self._mat = (t.empty(5 * n, dtype=self.x.dtype, device=self.x.device), None,
                             t.empty(n, dtype=self.x.dtype, device=self.x.device), None)
            if eng.sharded:
                it.f.comm.dist.barrier(group=getattr(it.f.comm, "group", None))
            eng.step(it.gamma, y=self._mat[0], z=self._mat[2], redo=True)
            if eng.sharded:
                t.cuda.synchronize(self.x.device)
                it.f.comm.dist.barrier(group=getattr(it.f.comm, "group", None))
            self._mat_valid = True
        return self._mat

    def _field(self, k, name):
        if self._fused or self._tv is not None:
            v = self._materialise()[k]
            if v is None:
                raise AttributeError(f"state.{name} is not materialised by the fused TV iteration (y and z are)")
            return v
        return getattr(self, "_" + name)

    y = property(lambda self: self._field(0, "y"))
    r = property(lambda self: self._field(1, "r"))
    z = property(lambda self: self._field(2, "z"))
    res = property(lambda self: self._field(3, "res"))

    @property
    def res_norm_inf(self):
        return self._R(self._sc.res_inf)


class DouglasRachfordIteration:
    """douglas_rachford.jl:31-42.  `gamma` has no default, as in the reference."""

    def __init__(self, x0, f=None, g=None, gamma=None, comm=None):
        if gamma is None:
            raise TypeError("DouglasRachfordIteration: keyword argument `gamma` not assigned")
        self.R = real_type(x0.dtype)
        self.x0 = x0
        self.f = f if f is not None else Zero()
        self.g = g if g is not None else Zero()
        self.gamma = self.R(gamma)
        if comm is None and getattr(self.f, "tv_split", False) and self.f.comm.size > 1:
            comm = self.f.comm                        # the row-sharded TV term carries the communicator of the run
        self.comm = comm

    def _prox(self, e, term, out, inp):
        """prox!(out, term, inp, gamma) with the value discarded (:58, :60)."""
        if hasattr(term, "prox_enqueue"):
            term.prox_enqueue(e.ctx, out, inp, self.gamma, comm=e.comm)
        else:
            term.prox_(out, inp, self.gamma)

    def step(self, st=None):                                                                # :54-63
        t = torch()
        R = self.R
        if st is None:
            st = DouglasRachfordState()
            e = _Engine(self, self.x0)
            st._engine, st._R, st._it = e, R, self
            st.x = _to_device_copy(self.x0, e.ctx)                                          # :56
            st._tv = None
            if getattr(self.f, "tv_split", False) and isinstance(self.g, IndConsensus):
                if self.g.K != self.f.ncopies:
                    raise ValueError("IndConsensus(K) must match the five copies of TVSplit")
                st._tv = TVDouglasRachfordEngine(self.f, st.x)
                st._fused = False
            else:
                st._fused = _elementwise(self.f) and _elementwise(self.g)
            if st._tv is not None:
                pass
            elif st._fused:
                self._fd, self._gd = self.f.descriptor(R), self.g.descriptor(R)
                st._x_in = t.empty_like(st.x)
            else:
                st._y, st._r, st._z, st._res = (t.empty_like(st.x) for _ in range(4))
        e = st._engine
        dt, n = pb_dtype(R), st.x.numel()
        if st._tv is not None:
            st.x = st._tv.step(self.gamma)                                                  # :58-62 in one pass (K10)
            st._mat_valid = False
        elif st._fused:
            st._x_in, st.x = st.x, st._x_in
            L.check(e.lib.pb_dr_step(e.ctx.h, dt, n, ptr(st._x_in), float(self.gamma), C.byref(self._fd), C.byref(self._gd),
                                     ptr(st.x), None, None, None, None))
            st._mat_valid = False
        else:
            self._prox(e, self.f, st._y, st.x)                                              # :58
            L.check(e.lib.pb_lincomb2(e.ctx.h, dt, n, 2.0, ptr(st._y), -1.0, ptr(st.x), ptr(st._r)))   # :59
            self._prox(e, self.g, st._z, st._r)                                             # :60
            L.check(e.lib.pb_residual(e.ctx.h, dt, n, ptr(st._y), ptr(st._z), None, ptr(st._res)))     # :61 (+ stop norm)
            L.check(e.lib.pb_lincomb2(e.ctx.h, dt, n, 1.0, ptr(st.x), -1.0, ptr(st._res), ptr(st.x)))  # :62
        _, st._sc = e.read()
        return st

    init = step

    def __iter__(self):
        st = self.step(None)
        while True:
            yield st
            st = self.step(st)


def default_stopping_criterion(tol, it, state):
    """douglas_rachford.jl:65-69: norm(res, Inf) / iter.gamma <= tol."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(state.res_norm_inf / it.gamma) <= float(tol)


def default_solution(it, state):
    return state.y                                                                          # :70


def default_display(k, it, state):
    with np.errstate(divide="ignore", invalid="ignore"):
        print("%5d | %.3e" % (k, float(state.res_norm_inf / it.gamma)))                      # :71-72


def DouglasRachford(maxit=1_000, tol=1e-8, stop=None, solution=default_solution, verbose=False, freq=100,
                    display=default_display, **kwargs):
    """douglas_rachford.jl:100-119."""
    if stop is None:
        def stop(it, state, _tol=tol):
            return default_stopping_criterion(_tol, it, state)
    return IterativeAlgorithm(DouglasRachfordIteration, maxit, stop, solution, verbose, freq, display, driver="python", **kwargs)
