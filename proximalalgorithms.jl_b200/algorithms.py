"""ForwardBackward / FastForwardBackward iterators, states and the IterativeAlgorithm driver, host side.

Same names, keyword arguments, defaults, state fields and control flow as the reference
(src/algorithms/forward_backward.jl, src/algorithms/fast_forward_backward.jl, src/utilities/fb_tools.jl,
src/ProximalAlgorithms.jl:58-123); every n-length vector lives on the device and every pass over one is a libproxb200
kernel.  What differs from a transliteration (and why):

  * one fused kernel (K1 `pb_fb_step` / K2 `pb_ffb_step`) replaces the reference's y / prox! / res broadcasts, the
    f_model reductions and the stop norm; `state.y` and `state.res` are materialised lazily on first access;
  * with a fixed stepsize the NEXT iteration's extrapolation x = z + beta*(z - z_prev) is computed inside the step
    kernel into a spare buffer and swapped in by the next `step` (beta only depends on host scalars);
  * there is exactly one host synchronisation per iteration (the scalar-block read-back that the driver's stop test
    needs anyway), two on the adaptive path (f(z) must be known before the line-search decision).
"""
from __future__ import annotations

import ctypes as C
import warnings

import numpy as np

from . import _lib as L
from .functions import Deferred, IndBallL2, Zero
from .host import Context, LocalComm, check_vec, pb_dtype, ptr, real_type, torch
from .nesterov import AdaptiveNesterovSequence


# ---------------------------------------------------------------------------------------------------------------------
# helpers
# ---------------------------------------------------------------------------------------------------------------------


def _to_device_copy(x0, ctx):
    """copy(x0) onto the device: x0 itself is never mutated (asserted by every reference test, e.g.
    test/problems/test_lasso_small.jl:54)."""
    t = torch()
    if isinstance(x0, t.Tensor):
        out = x0.detach().to(ctx.device, copy=True).contiguous().view(-1)
    else:
        arr = np.ascontiguousarray(x0).reshape(-1)
        out = t.as_tensor(arr).to(ctx.device, copy=True)
    real_type(out.dtype)
    return out


def _like_input(x0, z):
    """Return the solution in the container type of x0 (numpy in -> numpy out; CUDA tensor in -> CUDA tensor out)."""
    t = torch()
    if isinstance(x0, t.Tensor):
        if not x0.is_cuda and x0.is_pinned():
            out = t.empty(x0.shape, dtype=z.dtype, pin_memory=True)
            out.view(-1).copy_(z)
            return out
        return z.to(x0.device).view(x0.shape)
    return z.cpu().numpy().reshape(np.shape(x0))


def _resolve(v, R, row, comb):
    return R(v.resolve(row, comb)) if isinstance(v, Deferred) else R(v)


class _Engine:
    """Shared device machinery of the two iterations."""

    def __init__(self, it, x0):
        t = torch()
        dev = x0.device if isinstance(x0, t.Tensor) and x0.is_cuda else None
        self.ctx = Context.get(dev)
        self.comm = it.comm if it.comm is not None else self.ctx.default_comm()
        self.lib = self.ctx.lib

    # ---- smooth term ---------------------------------------------------------------------------------------------
    def eval_f(self, f, x, grad_out):
        """(value, gradient) at x, gradient written into grad_out.  Built-in terms run their kernels and return a Deferred
        value; user terms follow the reference contract `value_and_gradient(f, x) -> (f_x, fresh grad)` and the
        gradient is copied in, exactly like `state.grad_f_x .= grad_f_x` (fast_forward_backward.jl:139)."""
        if hasattr(f, "value_and_gradient_into"):
            return f.value_and_gradient_into(self.ctx, x, grad_out)
        val, g = f.value_and_gradient(x)
        if g.data_ptr() != grad_out.data_ptr():
            grad_out.copy_(g)
        return val

    def eval_f_value(self, f, x, grad_scratch):
        if hasattr(f, "value_into"):
            return f.value_into(self.ctx, x)
        return self.eval_f(f, x, grad_scratch)

    # ---- K1 / K2 -------------------------------------------------------------------------------------------------
    def fb_step(self, R, g, x, grad, gamma, z, z_prev=None, beta=None, x_next=None, y_scratch=None, res_out=None):
        """Enqueue y = x - gamma*grad, z = prox(y), res = x - z (+ optional x_next = z + beta*(z - z_prev)).
        `res_out`: materialise res (the line-search methods keep it as a state vector).  Returns `g_of(scalars)` -> g(z)."""
        lib, ctx, dt, n = self.lib, self.ctx, pb_dtype(R), x.numel()
        extrap = x_next is not None
        if getattr(g, "fused", False):
            d = g.descriptor(R)
            if extrap:
                L.check(lib.pb_ffb_step(ctx.h, dt, n, ptr(x), ptr(grad), ptr(z_prev), float(gamma), float(beta), C.byref(d),
                                        None, ptr(z), ptr(res_out), ptr(x_next)))
            else:
                L.check(lib.pb_fb_step(ctx.h, dt, n, ptr(x), ptr(grad), float(gamma), C.byref(d), None, ptr(z), ptr(res_out)))
            return lambda sc: g.value_from(R, sc.gsum)
        if isinstance(g, IndBallL2) and self.comm.size == 1:
            # one GPU: both phases are enqueued by the library (norm pass -> AUX3 slot, scale factor formed on the device)
            d = g.ball_descriptor(R)
            if extrap:
                L.check(lib.pb_ffb_step(ctx.h, dt, n, ptr(x), ptr(grad), ptr(z_prev), float(gamma), float(beta), C.byref(d),
                                        None, ptr(z), ptr(res_out), ptr(x_next)))
            else:
                L.check(lib.pb_fb_step(ctx.h, dt, n, ptr(x), ptr(grad), float(gamma), C.byref(d), None, ptr(z), ptr(res_out)))
            return lambda sc_: R(0)
        if isinstance(g, IndBallL2):
            # row shards: phase 1: y and ||y||^2 (combined over shards); phase 2: fused step with the scale factor
            L.check(lib.pb_forward(ctx.h, dt, n, ptr(x), ptr(grad), float(gamma), ptr(y_scratch)))
            sc = self.comm.exchange(ctx)
            d = g.scale_descriptor(R, sc.aux)
            if extrap:
                L.check(lib.pb_ffb_step(ctx.h, dt, n, ptr(x), ptr(grad), ptr(z_prev), float(gamma), float(beta), C.byref(d),
                                        None, ptr(z), ptr(res_out), ptr(x_next)))
            else:
                L.check(lib.pb_fb_step(ctx.h, dt, n, ptr(x), ptr(grad), float(gamma), C.byref(d), None, ptr(z), ptr(res_out)))
            return lambda sc_: R(0)
        # user-supplied proximable term: the reference's unfused sequence with its prox! callback in the middle
        L.check(lib.pb_forward(ctx.h, dt, n, ptr(x), ptr(grad), float(gamma), ptr(y_scratch)))
        g_z = g.prox_(z, y_scratch, R(gamma))
        L.check(lib.pb_residual(ctx.h, dt, n, ptr(x), ptr(z), ptr(grad), ptr(res_out)))
        if extrap:
            L.check(lib.pb_extrapolate(ctx.h, dt, n, ptr(z), ptr(z_prev), float(beta), ptr(x_next)))
        return lambda sc_: R(g_z)

    def pre_resolve(self, R, g, fx):
        """A Deferred f value lives in the AUX slot until the iteration's read-back.  The two-phase (IndBallL2) and user-prox
        forms of the step run `pb_forward` first, which reuses that slot for ||y||^2: fetch the value before they do.  Fused
        prox kinds never touch AUX, so for them this is a no-op (no extra synchronisation)."""
        if getattr(g, "fused", False) or not isinstance(fx, Deferred) or (isinstance(g, IndBallL2) and self.comm.size == 1):
            return fx
        row, sc = self.read()
        return _resolve(fx, R, row, sc)

    def read(self):
        """The one host synchronisation of an iteration: returns (this rank's raw row, rank-combined Scalars)."""
        sc = self.comm.exchange(self.ctx)
        return sc.parts[self.comm.rank], sc


def f_model(R, f_x, gdr, res_sq, Lc):
    """fb_tools.jl:3-5: f_x - real(dot(grad, res)) + (L/2)*norm(res)^2, from the step kernel's reductions.
    `norm(res)^2` is sqrt-then-square in R, as in the reference."""
    nr = R(np.sqrt(np.float64(res_sq)))
    return R(R(R(f_x) - R(gdr)) + R(R(R(Lc) / R(2)) * R(nr * nr)))


# ---------------------------------------------------------------------------------------------------------------------
# states
# ---------------------------------------------------------------------------------------------------------------------


class _LazyState:
    """Common part of the two state structs.  Field names are API (docs/src/guide/getting_started.jl:146-152)."""

    def __init__(self):
        self._y = None
        self._res = None
        self._y_valid = False
        self._res_valid = False

    def _invalidate(self):
        self._y_valid = False
        self._res_valid = False

    @property
    def y(self):
        """forward point x - gamma*grad_f_x (materialised on demand with the same two roundings as the fused kernel)."""
        if not self._y_valid:
            t = torch()
            if self._y is None:
                self._y = t.empty_like(self.x)
            e = self._engine
            L.check(e.lib.pb_forward(e.ctx.h, pb_dtype(self._R), self.x.numel(), ptr(self.x), ptr(self.grad_f_x), float(self.gamma), ptr(self._y)))
            self._y_valid = True
        return self._y

    @property
    def res(self):
        """fixed-point residual x - z (materialised on demand)."""
        if not self._res_valid:
            t = torch()
            if self._res is None:
                self._res = t.empty_like(self.x)
            e = self._engine
            L.check(e.lib.pb_residual(e.ctx.h, pb_dtype(self._R), self.x.numel(), ptr(self.x), ptr(self.z), None, ptr(self._res)))
            self._res_valid = True
        return self._res

    @property
    def res_norm_inf(self):
        """norm(res, Inf), from the fused kernel's reduction (no extra pass)."""
        return self._R(self._sc.res_inf)


class ForwardBackwardState(_LazyState):
    """forward_backward.jl:52-63: x, f_x, grad_f_x, gamma, y, z, g_z, res, (Az), grad_f_z."""


class FastForwardBackwardState(_LazyState):
    """fast_forward_backward.jl:60-71: x, f_x, grad_f_x, gamma, y, z, g_z, res, z_prev, extrapolation_sequence."""


# ---------------------------------------------------------------------------------------------------------------------
# iterations
# ---------------------------------------------------------------------------------------------------------------------


class ForwardBackwardIteration:
    """forward_backward.jl:38-48.  Keyword arguments and defaults as in the reference; `comm` is the only addition
    (a `TorchDistComm` when x0 is this rank's shard of a row-sharded iterate)."""

    def __init__(self, x0, f=None, g=None, Lf=None, gamma=None, adaptive=None, minimum_gamma=1e-7, reduce_gamma=0.5,
                 increase_gamma=1.0, comm=None, n_global=None):
        self.f = f if f is not None else Zero()
        self.g = g if g is not None else Zero()
        self.x0 = x0
        self.Lf = Lf
        self.gamma = (None if Lf is None else 1 / Lf) if gamma is None else gamma          # :43
        self.adaptive = (self.gamma is None) if adaptive is None else bool(adaptive)        # :44
        R = real_type(x0.dtype)
        self.R = R
        self.minimum_gamma = R(minimum_gamma)
        self.reduce_gamma = R(reduce_gamma)
        self.increase_gamma = R(increase_gamma)
        self.comm = comm
        self.n_global = n_global
        self.backtracks = 0

    # -- pieces shared with the fast variant ------------------------------------------------------------------------
    def _init_state(self, st):
        R = self.R
        e = _Engine(self, self.x0)
        t = torch()
        st._engine, st._R = e, R
        st.x = _to_device_copy(self.x0, e.ctx)                                              # :66 copy(x0)
        n = st.x.numel()
        n_glob = self.n_global if self.n_global is not None else n * e.comm.size if e.comm.size > 1 else n
        if hasattr(self.f, "gradient_buffer"):
            st.grad_f_x = self.f.gradient_buffer()
            check_vec(st.grad_f_x, n, st.x.dtype)
        else:
            st.grad_f_x = t.empty_like(st.x)
        fx = e.eval_f(self.f, st.x, st.grad_f_x)                                           # :67
        if self.gamma is None:                                                              # :68-70, fb_tools.jl:7-12
            row, sc = (e.read() if isinstance(fx, Deferred) else (None, None))
            st.f_x = _resolve(fx, R, row, sc)
            xeps = t.empty_like(st.x)
            L.check(e.lib.pb_add_scalar(e.ctx.h, pb_dtype(R), n, ptr(st.x), 1.0, ptr(xeps)))
            geps = t.empty_like(st.x)
            e.eval_f(self.f, xeps, geps)
            L.check(e.lib.pb_sub(e.ctx.h, pb_dtype(R), n, ptr(geps), ptr(st.grad_f_x), ptr(geps)))
            _, sc2 = e.read()
            lower = R(R(np.sqrt(np.float64(sc2.aux))) / R(np.sqrt(np.float64(n_glob))))
            with np.errstate(divide="ignore"):
                st.gamma = R(R(1) / lower)
            fx = st.f_x
        else:
            st.gamma = R(self.gamma)
        st.z = t.empty_like(st.x)
        st._y_scratch = None if getattr(self.g, "fused", False) else t.empty_like(st.x)
        return e, fx

    def _finish(self, st, fx, g_of):
        row, sc = st._engine.read()
        st._sc = sc
        st.f_x = _resolve(fx, self.R, row, sc)
        st.g_z = g_of(sc)
        st._invalidate()

    def init(self):
        st = ForwardBackwardState()
        e, fx = self._init_state(st)
        fx = e.pre_resolve(self.R, self.g, fx)
        g_of = e.fb_step(self.R, self.g, st.x, st.grad_f_x, st.gamma, st.z, y_scratch=st._y_scratch)   # :71-72, :82
        st.grad_f_z = torch().empty_like(st.x)                                                          # :62
        self._finish(st, fx, g_of)
        return st

    def _backtrack(self, st, want_grad):
        """fb_tools.jl:24-63 with A = nothing and Az aliased to z.  Returns (f_z, g_of)."""
        R, e = self.R, st._engine
        eps = R(np.finfo(R).eps)
        sc = st._sc
        g_z = st.g_z
        f_upp = f_model(R, st.f_x, sc.gdr, sc.res_sq, R(R(1) / st.gamma))                    # :42
        grad_buf = st.grad_f_z if want_grad else getattr(st, "_grad_scratch", None)
        if want_grad:
            fz = e.eval_f(self.f, st.z, grad_buf)                                           # :43-44
        else:
            if grad_buf is None and not hasattr(self.f, "value_into"):
                grad_buf = st._grad_scratch = torch().empty_like(st.x)
            fz = e.eval_f_value(self.f, st.z, grad_buf)
        row, sc_f = (e.read() if isinstance(fz, Deferred) else (None, None))
        f_z = _resolve(fz, R, row, sc_f)
        tol = R(R(10) * eps * R(R(1) + abs(f_z)))                                           # :45
        while f_z > R(f_upp + tol) and st.gamma >= self.minimum_gamma:                      # :46
            st.gamma = R(st.gamma * self.reduce_gamma)                                      # :47
            g_of = e.fb_step(R, self.g, st.x, st.grad_f_x, st.gamma, st.z, y_scratch=st._y_scratch)   # :48-50
            row, sc = e.read()
            st._sc = sc
            g_z = g_of(sc)
            st._invalidate()
            f_upp = f_model(R, st.f_x, sc.gdr, sc.res_sq, R(R(1) / st.gamma))               # :51
            fz = e.eval_f(self.f, st.z, grad_buf) if want_grad else e.eval_f_value(self.f, st.z, grad_buf)  # :52-53
            row, sc_f = (e.read() if isinstance(fz, Deferred) else (None, None))
            f_z = _resolve(fz, R, row, sc_f)
            tol = R(R(10) * eps * R(R(1) + abs(f_z)))                                       # :54
            self.backtracks += 1
        if st.gamma < self.minimum_gamma:                                                   # :59-61
            warnings.warn(f"stepsize `gamma` became too small ({st.gamma})")
        st.g_z = g_z
        return f_z

    def step(self, st):
        R, e = self.R, st._engine
        if self.adaptive:                                                                   # :90-110
            st.gamma = R(st.gamma * self.increase_gamma)
            st.f_x = self._backtrack(st, want_grad=True)
            st.x, st.z = st.z, st.x
            st.grad_f_x, st.grad_f_z = st.grad_f_z, st.grad_f_x
            fx = st.f_x
        else:                                                                               # :111-115
            st.x, st.z = st.z, st.x
            fx = e.pre_resolve(R, self.g, e.eval_f(self.f, st.x, st.grad_f_x))
        g_of = e.fb_step(R, self.g, st.x, st.grad_f_x, st.gamma, st.z, y_scratch=st._y_scratch)       # :117-120
        self._finish(st, fx, g_of)
        return st

    def __iter__(self):
        st = self.init()
        while True:
            yield st
            st = self.step(st)


class FastForwardBackwardIteration(ForwardBackwardIteration):
    """fast_forward_backward.jl:44-56."""

    def __init__(self, x0, f=None, g=None, mf=0, Lf=None, gamma=None, adaptive=None, minimum_gamma=1e-7,
                 reduce_gamma=0.5, increase_gamma=1.0, extrapolation_sequence=None, comm=None, n_global=None):
        super().__init__(x0, f, g, Lf, gamma, adaptive, minimum_gamma, reduce_gamma, increase_gamma, comm, n_global)
        self.mf = self.R(mf)
        self.extrapolation_sequence = extrapolation_sequence

    def _next_beta(self, st):
        """fast_forward_backward.jl:99-104."""
        seq = st.extrapolation_sequence
        if isinstance(seq, AdaptiveNesterovSequence):
            return seq.next(st.gamma)
        return self.R(next(seq))

    def _next_beta_lookahead(self, st):
        """The fixed-stepsize path fuses the NEXT iteration's extrapolation into the current pass, so it draws beta one iteration
        before the reference does (fast_forward_backward.jl:134).  A finite user sequence may therefore run out here although the
        driver would have stopped first: remember it and fail only if that next iteration is really taken (the value is irrelevant
        until then: the speculative x_next is discarded)."""
        try:
            return self._next_beta(st)
        except StopIteration:
            st._sequence_exhausted = True
            return self.R(0)

    def init(self):
        st = FastForwardBackwardState()
        e, fx = self._init_state(st)                                                        # :74-78
        fx = e.pre_resolve(self.R, self.g, fx)
        t = torch()
        st.z_prev = st.x.clone()                                                            # :69 default z_prev = copy(x)
        if self.extrapolation_sequence is not None:                                         # :90-94
            st.extrapolation_sequence = iter(self.extrapolation_sequence)
        else:
            st.extrapolation_sequence = AdaptiveNesterovSequence(self.mf)
        st._x_next = None
        st._beta_next = None
        if not self.adaptive:
            # fixed stepsize: beta of the next step is a pure host scalar -> fuse its extrapolation into this pass
            st._x_next = t.empty_like(st.x)
            st._beta_next = self._next_beta_lookahead(st)
            g_of = e.fb_step(self.R, self.g, st.x, st.grad_f_x, st.gamma, st.z, z_prev=st.z_prev, beta=st._beta_next,
                             x_next=st._x_next, y_scratch=st._y_scratch)                    # :79-80, :89 (+ :135 of step 2)
        else:
            g_of = e.fb_step(self.R, self.g, st.x, st.grad_f_x, st.gamma, st.z, y_scratch=st._y_scratch)
        self._finish(st, fx, g_of)
        return st

    def step(self, st):
        R, e = self.R, st._engine
        dt, n = pb_dtype(R), st.x.numel()
        if self.adaptive:                                                                   # :110-129
            st.gamma = R(st.gamma * self.increase_gamma)
            self._backtrack(st, want_grad=False)
            beta = self._next_beta(st)                                                      # :134
            st.beta = beta
            L.check(e.lib.pb_extrapolate(e.ctx.h, dt, n, ptr(st.z), ptr(st.z_prev), float(beta), ptr(st.x)))   # :135
            st.z_prev, st.z = st.z, st.z_prev                                               # :136
            fx = e.pre_resolve(R, self.g, e.eval_f(self.f, st.x, st.grad_f_x))              # :138-139
            g_of = e.fb_step(R, self.g, st.x, st.grad_f_x, st.gamma, st.z, y_scratch=st._y_scratch)    # :140-142
        else:
            if getattr(st, "_sequence_exhausted", False):
                raise RuntimeError("extrapolation_sequence is exhausted (fast_forward_backward.jl:99-104 draws one coefficient per iteration)")
            st.gamma = R(self.gamma)                                                        # :130-132
            st.beta = st._beta_next
            st.x, st._x_next = st._x_next, st.x                                             # :135 (computed by the previous pass)
            st.z_prev, st.z = st.z, st.z_prev                                               # :136
            fx = e.pre_resolve(R, self.g, e.eval_f(self.f, st.x, st.grad_f_x))              # :138-139
            st._beta_next = self._next_beta_lookahead(st)
            g_of = e.fb_step(R, self.g, st.x, st.grad_f_x, st.gamma, st.z, z_prev=st.z_prev, beta=st._beta_next,
                             x_next=st._x_next, y_scratch=st._y_scratch)                    # :140-142 (+ next :135)
        self._finish(st, fx, g_of)
        return st


# ---------------------------------------------------------------------------------------------------------------------
# defaults (forward_backward.jl:125-129, fast_forward_backward.jl:147-154) and driver (ProximalAlgorithms.jl:58-123)
# ---------------------------------------------------------------------------------------------------------------------


def default_stopping_criterion(tol, it, state):
    """norm(state.res, Inf) / state.gamma <= tol, with the norm taken from the fused kernel's reduction."""
    with np.errstate(divide="ignore", invalid="ignore"):
        return float(state.res_norm_inf / state.gamma) <= float(tol)


def default_solution(it, state):
    return state.z


def default_display(k, it, state):
    with np.errstate(divide="ignore", invalid="ignore"):
        print("%5d | %.3e | %.3e" % (k, float(state.gamma), float(state.res_norm_inf / state.gamma)))


class IterativeAlgorithm:
    """src/ProximalAlgorithms.jl:58-66, :103-123: partial application of an iterator type plus the driver loop.

    `driver` (our addition): "auto" runs the loop inside the library (`pb_solve`, csrc/solve.cu) when every ingredient is
    built in -- default stop / solution, not verbose, library smooth term, single-pass prox, known extrapolation sequence,
    local or device exchange -- and this Python loop otherwise (user callbacks, custom stop, verbose, IndBallL2, NCCL
    exchange).  "python" / "native" force one of them.  Both issue the same kernels and the same scalar arithmetic."""

    def __init__(self, iterator_type, maxit, stop, solution, verbose, freq, display, driver="auto", **kwargs):
        self.iterator_type = iterator_type
        self.maxit, self.stop, self.solution = int(maxit), stop, solution
        self.verbose, self.freq, self.display = bool(verbose), int(freq), display
        if driver not in ("auto", "python", "native"):
            raise ValueError("driver must be 'auto', 'python' or 'native'")
        self.driver = driver
        self.kwargs = kwargs
        self.last_driver = None

    def __call__(self, **kwargs):
        it = self.iterator_type(**{**self.kwargs, **kwargs})                                # :115
        if self.driver != "python":
            out = _native_solve(self, it)
            if out is not None:
                self.last_driver = "native"
                return out
            if self.driver == "native":
                raise L.ProxB200Error("driver='native' requested but the problem needs the Python loop "
                                      "(user callbacks, custom stop/solution, verbose, IndBallL2 or an NCCL exchange)")
        self.last_driver = "python"
        for k, state in enumerate(it, start=1):                                             # :116
            if k >= self.maxit or self.stop(it, state):                                     # :117
                if self.verbose:
                    self.display(k, it, state)
                self.last_iteration, self.last_state = it, state
                sol = _like_input(it.x0, self.solution(it, state))                          # :119
                if hasattr(state, "release"):     # states that own resources outside torch's allocator (row-sharded TV engine)
                    state.release()
                return sol, k
            if self.verbose and k % self.freq == 0:                                         # :121
                self.display(k, it, state)


def _native_sequence(it, R):
    """Map the iteration's extrapolation sequence to (PB_SEQ_*, constant_beta) or None if only Python can run it."""
    import itertools

    from .nesterov import FixedNesterovSequence, SimpleNesterovSequence

    seq = getattr(it, "extrapolation_sequence", None)
    if seq is None:
        return L.PB_SEQ_ADAPTIVE, 0.0
    if isinstance(seq, FixedNesterovSequence) and seq.R is R:
        return L.PB_SEQ_FIXED, 0.0
    if isinstance(seq, SimpleNesterovSequence) and seq.R is R:
        return L.PB_SEQ_SIMPLE, 0.0
    if isinstance(seq, itertools.repeat):
        try:
            seq.__length_hint__()              # repeat(x, times) is finite: only the Python loop reproduces its exhaustion
            return None
        except TypeError:                      # "len() of unsized object": the unbounded repeat(x)
            return L.PB_SEQ_CONSTANT, float(R(next(seq)))
    return None


def _native_solve(alg, it):
    """Run the whole solve in `pb_solve` if possible; returns (solution, iterations) or None."""
    from .host import DeviceExchangeComm

    if type(it) not in (ForwardBackwardIteration, FastForwardBackwardIteration):
        return None           # pb_solve implements the two forward-backward loops only
    tol = getattr(alg.stop, "_default_tol", None)
    if tol is None or alg.solution is not default_solution or alg.verbose:
        return None
    f, g, R = it.f, it.g, it.R
    comm = it.comm
    ball = isinstance(g, IndBallL2) and (comm is None or comm.size == 1)         # one GPU: PB_PROX_BALL, both phases on the device
    if not hasattr(f, "native_descriptor") or not (getattr(g, "fused", False) or ball):
        return None
    if not ball and g.kind not in (L.PB_PROX_ZERO, L.PB_PROX_L1, L.PB_PROX_BOX, L.PB_PROX_L21):
        return None
    if comm is not None and not isinstance(comm, (LocalComm, DeviceExchangeComm)):
        return None
    fdesc = f.native_descriptor()
    if fdesc is None:
        return None
    fast = isinstance(it, FastForwardBackwardIteration)
    seq = _native_sequence(it, R) if fast else (L.PB_SEQ_ADAPTIVE, 0.0)
    if seq is None or (it.gamma is None and not it.adaptive):
        return None
    import time

    t = torch()
    e = _Engine(it, it.x0)
    if isinstance(e.comm, DeviceExchangeComm) and e.comm.ctx is not e.ctx:
        return None
    if isinstance(e.comm, LocalComm) and getattr(e.ctx, "_xchg_comm", None) is not None:
        return None           # pb_solve reads through the attached exchange; an explicit memcpy read-back needs the Python loop
    t0 = time.perf_counter()
    # Work vectors.  When x0 lives on the host the solution is copied out at the end, so the device vectors are private to this call and
    # are kept on the solver object for the next call of the same size (a repeated solve then allocates nothing: the allocator's
    # occasional cudaMalloc was the largest and least predictable part of the host-buffer call).  A device x0 gets fresh vectors: the
    # returned solution and `last_state` alias them.
    host_in = not (isinstance(it.x0, t.Tensor) and it.x0.is_cuda)
    pipelined = fast and not it.adaptive and isinstance(e.comm, DeviceExchangeComm) and getattr(alg, "pipeline", True)
    # block-diagonal least squares with a fixed stepsize: the one-sweep kernel (csrc/lsq_fista.cu) is launched one iteration ahead of the
    # stop decision whenever it has spare vectors, with or without the device exchange
    lookahead = pipelined or (fast and not it.adaptive and getattr(alg, "pipeline", True) and fdesc.kind == L.PB_F_LSQ_BLOCKDIAG)
    n0 = int(np.prod(np.shape(it.x0))) if not isinstance(it.x0, t.Tensor) else it.x0.numel()
    key = (n0, str(it.x0.dtype).replace("torch.", ""), e.ctx.index, fast, bool(it.adaptive), lookahead)
    cache = getattr(alg, "_workspace", None) if host_in else None
    if cache is not None and cache[0] == key:
        x, z, scratch, z_prev, x_next, grad_z, spare_x, spare_z, own_grad = cache[1]
        if isinstance(it.x0, t.Tensor):
            x.copy_(it.x0.detach().contiguous().view(-1), non_blocking=True)
        else:
            x.copy_(t.as_tensor(np.ascontiguousarray(it.x0).reshape(-1)))
    else:
        x = _to_device_copy(it.x0, e.ctx)
        z, scratch = t.empty_like(x), t.empty_like(x)
        z_prev = t.empty_like(x) if fast else None
        x_next = t.empty_like(x) if fast and not it.adaptive else None
        grad_z = t.empty_like(x) if (not fast and it.adaptive) else None
        # fixed-stepsize FFB through the device exchange: two spare vectors let pb_solve run ahead of the stop decision (one launch per
        # iteration: pb_solve_opts.spare_*; one persistent kernel: the third buffers of its x / z rings); `scratch` doubles as the
        # spare gradient buffer
        spare_x = t.empty_like(x) if lookahead else None
        spare_z = t.empty_like(x) if lookahead else None
        own_grad = None if hasattr(f, "gradient_buffer") else t.empty_like(x)
        if host_in:
            alg._workspace = (key, (x, z, scratch, z_prev, x_next, grad_z, spare_x, spare_z, own_grad))
    n = x.numel()
    grad = f.gradient_buffer() if hasattr(f, "gradient_buffer") else own_grad
    check_vec(grad, n, x.dtype)
    bufs = {b.data_ptr(): b for b in (x, grad, z, scratch, z_prev, x_next, grad_z, spare_x, spare_z) if b is not None}
    n_glob = it.n_global if it.n_global is not None else n * e.comm.size
    opts = L.pb_solve_opts(L.PB_ALG_FFB if fast else L.PB_ALG_FB, 1 if it.adaptive else 0, seq[0], 1 if getattr(alg, "profile", False) else 0, alg.maxit, int(n_glob), float(tol),
                           0.0 if it.gamma is None else float(R(it.gamma)), float(getattr(it, "mf", 0.0)), seq[1],
                           float(it.minimum_gamma), float(it.reduce_gamma), float(it.increase_gamma),
                           spare_x.data_ptr() if lookahead else None, spare_z.data_ptr() if lookahead else None,
                           scratch.data_ptr() if lookahead else None)
    gdesc = g.ball_descriptor(R) if ball else g.descriptor(R)
    res = L.pb_solve_result()
    t1 = time.perf_counter()
    L.check(e.lib.pb_solve(e.ctx.h, pb_dtype(R), n, C.byref(fdesc), C.byref(gdesc), C.byref(opts), ptr(x), ptr(grad), ptr(z), ptr(z_prev),
                           ptr(x_next), ptr(grad_z), ptr(scratch), C.byref(res)))
    t2 = time.perf_counter()
    if res.warned_small_gamma:
        warnings.warn(f"stepsize `gamma` became too small ({R(res.gamma)})")
    st = FastForwardBackwardState() if fast else ForwardBackwardState()
    st._engine, st._R = e, R
    st.x, st.grad_f_x, st.z = bufs[res.x], bufs[res.grad], bufs[res.z]
    if fast:
        st.z_prev = bufs[res.z_prev]
    st.gamma, st.f_x, st.g_z = R(res.gamma), R(res.f_x), R(res.g_z)

    class _Sc:
        res_inf = res.res_inf

    st._sc = _Sc
    it.backtracks = int(res.backtracks)
    alg.last_persistent_ctas = int(res.persistent_ctas)     # > 0: the whole solve was ONE persistent kernel (csrc/persist.cu)
    alg.last_multi_iter_kernel = bool(res.multi_iter_kernel)   # the iterations ran inside one persistent step kernel (csrc/step_multi.cu)
    alg.last_parity = {"res_inf": float(res.res_inf), "res_sq": float(res.res_sq), "gdr": float(res.gdr), "gsum": float(res.gsum)}
    alg.last_iteration, alg.last_state = it, st
    sol = _like_input(it.x0, st.z)
    # host-side wall clock of the phases (pb_solve returns after its last scalar read, i.e. with the stream drained)
    alg.last_timing = {"upload_and_alloc_s": t1 - t0, "pb_solve_s": t2 - t1, "download_s": time.perf_counter() - t2}
    if getattr(alg, "profile", False):
        alg.last_timing.update(loop_ms=res.loop_ms, step_kernel_ms=res.step_kernel_ms, step_kernel_launches=int(res.step_kernel_launches))
    return sol, int(res.iterations)


def ForwardBackward(maxit=10_000, tol=1e-8, stop=None, solution=default_solution, verbose=False, freq=100,
                    display=default_display, **kwargs):
    """forward_backward.jl:161-179."""
    if stop is None:
        def stop(it, state, _tol=tol):
            return default_stopping_criterion(_tol, it, state)
        stop._default_tol = tol
    return IterativeAlgorithm(ForwardBackwardIteration, maxit, stop, solution, verbose, freq, display, **kwargs)


def FastForwardBackward(maxit=10_000, tol=1e-8, stop=None, solution=default_solution, verbose=False, freq=100,
                        display=default_display, **kwargs):
    """fast_forward_backward.jl:186-204."""
    if stop is None:
        def stop(it, state, _tol=tol):
            return default_stopping_criterion(_tol, it, state)
        stop._default_tol = tol
    return IterativeAlgorithm(FastForwardBackwardIteration, maxit, stop, solution, verbose, freq, display, **kwargs)


# aliases (forward_backward.jl:183-184, fast_forward_backward.jl:208-209)
ProximalGradientIteration = ForwardBackwardIteration
ProximalGradient = ForwardBackward
FastProximalGradientIteration = FastForwardBackwardIteration
FastProximalGradient = FastForwardBackward
