"""AFBA (asymmetric forward-backward-adjoint), Vu-Condat and Chambolle-Pock on the device.

Reference: src/algorithms/primal_dual.jl:57-111 (parameters, defaults, errors), :113-159 (VuCondat / ChambollePock = AFBA with
theta = 2), :161-172 (state), :174-211 (iterate), :213-218 (stop rule, solution, display), :334-427 (default stepsizes).  `l` is
the default IndZero (its conjugate has zero gradient); other `l` are not supported.

Mapping of one iteration onto library kernels (every vector on the device):
  * :181-185  temp_x = x - gamma1*(L'y + grad f), xbar = prox_{gamma1 g}(temp_x), FPR_x = xbar - x and norm(FPR_x, Inf):
              ONE fused pass (K1 `pb_fb_step` with `grad = L'y + grad f`; its residual is -FPR_x)
  * :190-195  temp_y = y + gamma2*L(theta*xbar + (1-theta)*x); ybar = prox of the conjugate of h (Moreau, `pb_conj_prox`)
  * :198-209  the two corrections are skipped when their coefficient mu(2-theta)gamma1 / (1-mu)(2-theta)gamma2 is exactly 0
              (adding an exact zero vector changes nothing), which removes two of the four products with L for theta = 2.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib as L_
from .algorithms import IterativeAlgorithm, _Engine, _to_device_copy
from .functions import MatrixOp, Zero
from .host import pb_dtype, ptr, real_type, torch


def _approx(a, b, R):
    return abs(float(a) - float(b)) <= float(np.sqrt(np.finfo(R).eps)) * max(abs(float(a)), abs(float(b)))


def AFBA_default_stepsizes(nmL, h_is_zero, theta, mu, beta_f, beta_l, R):
    """primal_dual.jl:334-427, all arithmetic in R."""
    theta, mu, beta_f, beta_l = R(theta), R(mu), R(beta_f), R(beta_l)
    if h_is_zero:
        with np.errstate(divide="ignore"):
            return R(R(1.99) / beta_f), R(1)
    par, par2, alpha = R(5), R(100), R(1)
    nmL = R(nmL)

    def pick():
        if nmL > par * max(beta_l, beta_f):
            return R(1)
        if beta_f > par * beta_l:
            return R(par2 * nmL / beta_f)
        if beta_l > par * beta_f:
            return R(beta_l / (par2 * nmL))
        return R(1)

    if _approx(theta, 2, R):
        alpha = pick()
        g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
        g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
    elif _approx(theta, 1, R) and _approx(mu, 1, R):
        if nmL > par2 * beta_l:
            alpha = R(1)
        elif beta_l > par * beta_f:
            alpha = R(beta_l / (par2 * nmL))
        g1 = R(R(1.99) / beta_f) if beta_f > 0 else R(R(1) / R(nmL / alpha))
        g2 = R(R(0.99) / R(beta_l / R(2) + g1 * R(nmL * nmL)))
    elif _approx(theta, 0, R) and _approx(mu, 1, R):
        if _approx(beta_f, 0, R):
            nmL = R(nmL * np.sqrt(R(3)))
            alpha = R(1) if nmL > par * beta_l else R(beta_l / (par2 * nmL))
            g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
            g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
        else:
            alpha = pick()
            xi = R(R(1) + R(2) * nmL / R(nmL + alpha * beta_f / R(2)))
            g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
            g2 = R(R(0.99) / R(beta_l / R(2) + xi * nmL * alpha))
    elif _approx(mu, 0, R):
        temp = R(theta * theta - R(3) * theta + R(3))
        if _approx(beta_l, 0, R):
            nmL = R(nmL * np.sqrt(temp))
            with np.errstate(divide="ignore"):
                alpha = R(1) if nmL > par * beta_f else R(par2 * nmL / beta_f)
            g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
            g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
        else:
            alpha = pick()
            eta = R(R(1) + (temp - R(1)) * alpha * nmL / R(alpha * nmL + beta_l / R(2)))
            g1 = R(R(1) / R(beta_f / R(2) + eta * nmL / alpha))
            g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
    elif _approx(theta, 0, R) and _approx(mu, 0.5, R):
        if _approx(beta_l, 0, R) or _approx(beta_f, 0, R):
            alpha = pick()
        else:
            alpha = R(np.sqrt(R(beta_l / beta_f)) / R(2))
        g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
        g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
    else:
        raise ValueError("this choice of theta and mu is not supported!")
    return g1, g2


class AFBAState:
    """primal_dual.jl:161-172: x, y, xbar, ybar, gradf, FPR_x, FPR_y (+ scratch)."""

    @property
    def FPR_x(self):
        """xbar - x of the last iteration (the fused step keeps x - xbar; the sign flip is exact)."""
        e = self._engine
        L_.check(e.lib.pb_scale(e.ctx.h, pb_dtype(self._R), self._negFPR_x.numel(), -1.0, ptr(self._negFPR_x), ptr(self._fprx)))
        return self._fprx


class AFBAIteration:
    """primal_dual.jl:57-111.  `L=None` is the identity (0*I when h is Zero, :62-66); a matrix is wrapped into `MatrixOp`."""

    def __init__(self, x0, y0, f=None, g=None, h=None, l=None, L=None, beta_f=None, beta_l=None, theta=1, mu=1, lambda_=1,
                 gamma=None, comm=None, **kw):
        if "lambda" in kw:
            lambda_ = kw.pop("lambda")
        if kw:
            raise TypeError(f"unexpected keyword arguments {sorted(kw)}")
        if l is not None:
            raise NotImplementedError("only the default l = IndZero() is supported")
        R = real_type(x0.dtype)
        self.R, self.x0, self.y0 = R, x0, y0
        self.f = f if f is not None else Zero()
        self.g = g if g is not None else Zero()
        self.h = h if h is not None else Zero()
        self.h_zero = isinstance(self.h, Zero)
        for term, nm in ((self.g, "g"), (self.h, "h")):
            if not getattr(term, "fused", False) or term.kind not in (L_.PB_PROX_ZERO, L_.PB_PROX_L1, L_.PB_PROX_BOX, L_.PB_PROX_SQRL2):
                raise NotImplementedError(f"AFBA on the device supports element-wise {nm} (Zero, NormL1, IndBox, SqrNormL2)")
        nmL = None
        if L is not None and not hasattr(L, "mul_into"):
            nmL = float(np.linalg.norm(np.asarray(L.detach().cpu().numpy() if hasattr(L, "detach") else L, np.float64), 2))
            L = MatrixOp(L, device=x0.device if hasattr(x0, "is_cuda") and x0.is_cuda else None)
        self.L = L
        if beta_f is None:
            if not isinstance(self.f, Zero):
                raise ValueError("argument beta_f must be specified together with f")          # :70-74
            beta_f = 0
        self.beta_f = beta_f
        self.theta, self.mu, self.lambda_ = R(theta), R(mu), R(lambda_)
        if gamma is None:
            if self.lambda_ != 1:
                raise ValueError("if lambda != 1, then you need to provide stepsizes manually")   # :105-106
            if L is None:
                nmL = 0.0 if self.h_zero else 1.0
            elif nmL is None:
                nmL = L.opnorm() if hasattr(L, "opnorm") else None
                if nmL is None:
                    raise ValueError("pass `gamma=(gamma1, gamma2)` or a matrix L (the default stepsizes need opnorm(L))")
            gamma = AFBA_default_stepsizes(nmL, self.h_zero, theta, mu, beta_f, 0, R)
        self.gamma = (R(gamma[0]), R(gamma[1]))
        self.comm = comm

    # L and L' (the reference's `L = 0*I` when h is Zero gives zero vectors)
    def _Lmul(self, out, x):
        if self.L is None:
            if self.h_zero:
                out.zero_()
            else:
                out.copy_(x)
            return out
        return self.L.mul_into(out, x)

    def _Ltmul(self, out, y):
        if self.L is None:
            if self.h_zero:
                out.zero_()
            else:
                out.copy_(y)
            return out
        return self.L.mul_t_into(out, y)

    def step(self, st=None):                                                                # :174-211
        R, t = self.R, torch()
        g1, g2 = self.gamma
        if st is None:
            st = AFBAState()
            e = _Engine(self, self.x0)
            st._engine, st._R = e, R
            st.x = _to_device_copy(self.x0, e.ctx)                                          # :176
            st.y = _to_device_copy(self.y0, e.ctx)
            st.xbar, st.gradf, st.temp_x, st._negFPR_x, st._fprx = (t.empty_like(st.x) for _ in range(5))
            st.ybar, st.FPR_y, st.temp_y = (t.empty_like(st.y) for _ in range(3))
            self._gd, self._hd = self.g.descriptor(R), self.h.descriptor(R)
        e = st._engine
        lib, h_, dt = e.lib, e.ctx.h, pb_dtype(R)
        n, m = st.x.numel(), st.y.numel()
        e.eval_f(self.f, st.x, st.gradf)                                                    # :179-180
        self._Ltmul(st.temp_x, st.y)                                                        # :181
        L_.check(lib.pb_lincomb2(h_, dt, n, 1.0, ptr(st.temp_x), 1.0, ptr(st.gradf), ptr(st.temp_x)))   # :182
        # :183-185 and :198: x - gamma1*temp_x, prox, x - xbar (= -FPR_x), norm(FPR_x, Inf)
        L_.check(lib.pb_fb_step(h_, dt, n, ptr(st.x), ptr(st.temp_x), float(g1), C.byref(self._gd), None, ptr(st.xbar), ptr(st._negFPR_x)))
        _, sc_x = e.read()
        # :190-195
        L_.check(lib.pb_lincomb2(h_, dt, n, float(self.theta), ptr(st.xbar), float(R(R(1) - self.theta)), ptr(st.x), ptr(st.temp_x)))
        self._Lmul(st.temp_y, st.temp_x)
        L_.check(lib.pb_forward(h_, dt, m, ptr(st.y), ptr(st.temp_y), float(-g2), ptr(st.temp_y)))      # y + gamma2*temp_y
        L_.check(lib.pb_conj_prox(h_, dt, m, ptr(st.temp_y), float(g2), C.byref(self._hd), ptr(st.ybar)))
        L_.check(lib.pb_sub(h_, dt, m, ptr(st.ybar), ptr(st.y), ptr(st.FPR_y)))             # :199
        L_.check(lib.pb_nrm2sq(h_, dt, m, ptr(st.FPR_y)))                                   # norm(FPR_y, Inf) -> AUXINF
        _, sc_y = e.read()
        st._fpr_norm = R(R(sc_x.res_inf) + R(sc_y.aux_inf))                                 # :213-215
        # :202-204  x += lambda*(FPR_x - L'(c1*FPR_y))
        c1 = R(R(self.mu * R(R(2) - self.theta)) * g1)
        if c1 != 0:
            L_.check(lib.pb_scale(h_, dt, m, float(c1), ptr(st.FPR_y), ptr(st.temp_y)))
            self._Ltmul(st.temp_x, st.temp_y)
            L_.check(lib.pb_lincomb2(h_, dt, n, -1.0, ptr(st._negFPR_x), -1.0, ptr(st.temp_x), ptr(st.temp_x)))
        else:
            L_.check(lib.pb_scale(h_, dt, n, -1.0, ptr(st._negFPR_x), ptr(st.temp_x)))
        c2 = R(R(R(R(1) - self.mu) * R(R(2) - self.theta)) * g2)                            # :207 (uses FPR_x before x moves)
        if c2 != 0:
            fx = st.gradf                                                                   # gradf is free from here on: scratch
            L_.check(lib.pb_scale(h_, dt, n, float(-c2), ptr(st._negFPR_x), ptr(fx)))       # c2*FPR_x
        L_.check(lib.pb_lincomb2(h_, dt, n, 1.0, ptr(st.x), float(self.lambda_), ptr(st.temp_x), ptr(st.x)))
        # :207-209  y += lambda*(FPR_y + L(c2*FPR_x))
        if c2 != 0:
            self._Lmul(st.temp_y, fx)
            L_.check(lib.pb_lincomb2(h_, dt, m, 1.0, ptr(st.FPR_y), 1.0, ptr(st.temp_y), ptr(st.temp_y)))
            L_.check(lib.pb_lincomb2(h_, dt, m, 1.0, ptr(st.y), float(self.lambda_), ptr(st.temp_y), ptr(st.y)))
        else:
            L_.check(lib.pb_lincomb2(h_, dt, m, 1.0, ptr(st.y), float(self.lambda_), ptr(st.FPR_y), ptr(st.y)))
        return st

    init = step

    def __iter__(self):
        st = self.step(None)
        while True:
            yield st
            st = self.step(st)


def VuCondatIteration(**kwargs):
    """primal_dual.jl:139."""
    return AFBAIteration(**{**kwargs, "theta": 2})


def ChambollePockIteration(**kwargs):
    """primal_dual.jl:158-159."""
    return AFBAIteration(**{**kwargs, "theta": 2, "f": None, "l": None})


def default_stopping_criterion(tol, it, state):
    return float(state._fpr_norm) <= float(tol)                                             # :213-215


def default_solution(it, state):
    return (state.xbar, state.ybar)                                                         # :216


def default_display(k, it, state):
    print("%6d | %7.4e" % (k, float(state._fpr_norm)))                                      # :217-218


class _PairAlgorithm(IterativeAlgorithm):
    """The solution is the pair (xbar, ybar): return each in the container type of x0 / y0."""

    def __call__(self, **kwargs):
        from .algorithms import _like_input

        it = self.iterator_type(**{**self.kwargs, **kwargs})
        for k, state in enumerate(it, start=1):
            if k >= self.maxit or self.stop(it, state):
                if self.verbose:
                    self.display(k, it, state)
                self.last_iteration, self.last_state = it, state
                xb, yb = self.solution(it, state)
                return (_like_input(it.x0, xb), _like_input(it.y0, yb)), k
            if self.verbose and k % self.freq == 0:
                self.display(k, it, state)


def AFBA(maxit=10_000, tol=1e-5, stop=None, solution=default_solution, verbose=False, freq=100, display=default_display, **kwargs):
    """primal_dual.jl:245-263."""
    if stop is None:
        def stop(it, state, _tol=tol):
            return default_stopping_criterion(_tol, it, state)
    return _PairAlgorithm(AFBAIteration, maxit, stop, solution, verbose, freq, display, driver="python", **kwargs)


def VuCondat(**kwargs):
    """primal_dual.jl:296."""
    return AFBA(**{**kwargs, "theta": 2})


def ChambollePock(**kwargs):
    """primal_dual.jl:331."""
    return AFBA(**{**kwargs, "f": None, "l": None, "theta": 2})
