"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

numpy restatement of row f4 of SURVEY.md section 8f: the asymmetric forward-backward-adjoint iteration (AFBA) and its special
cases Vu-Condat and Chambolle-Pock -- src/algorithms/primal_dual.jl:57-111 (parameters and defaults), :113-159 (VuCondat /
ChambollePock = AFBA with theta = 2), :161-172 (state), :174-211 (the one `iterate` method), :213-218 (stop rule, solution,
display), :334-427 (default stepsizes).  `l` is restricted to the default IndZero (its conjugate is Zero: gradient 0).

Prox of a convex conjugate follows ProximalCore's Moreau identity (third party, restated):
    prox_{gamma h*}(v) = v - gamma * prox_{h/gamma}(v/gamma)
Pinned (tests/test_oracle_afba.py) against x_star and the iteration bounds of test/problems/test_lasso_small.jl:233-275 (three
formulations), test/problems/test_elasticnet.jl:56-113 (five (theta, mu) pairs) and the linear program of
test/problems/test_linear_programs.jl:44-151 (AFBA and VuCondat to 1000 * 100 eps in all four optimality measures).
"""
from __future__ import annotations

import numpy as np

from .fb_oracle import State, ZeroFn, _R, norm_inf


def _approx(a, b, R):
    """Julia `a ≈ b` for reals: |a - b| <= sqrt(eps(R)) * max(|a|, |b|)."""
    return abs(float(a) - float(b)) <= float(np.sqrt(np.finfo(R).eps)) * max(abs(float(a)), abs(float(b)))


def default_stepsizes(nmL, h_is_zero, theta, mu, beta_f, beta_l, R):
    """primal_dual.jl:334-427 (all arithmetic in R)."""
    theta, mu, beta_f, beta_l = R(theta), R(mu), R(beta_f), R(beta_l)
    if h_is_zero:                                                             # :334-336
        with np.errstate(divide="ignore"):
            return R(R(1.99) / beta_f), R(1)
    par, par2, alpha = R(5), R(100), R(1)
    nmL = R(nmL)
    pick = lambda: (R(1) if nmL > par * max(beta_l, beta_f) else                # noqa: E731
                    R(par2 * nmL / beta_f) if beta_f > par * beta_l else
                    R(beta_l / (par2 * nmL)) if beta_l > par * beta_f else R(1))
    if _approx(theta, 2, R):                                                   # Vu-Condat :344-353
        alpha = pick()
        g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
        g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
    elif _approx(theta, 1, R) and _approx(mu, 1, R):                           # SPCA :354-361
        if nmL > par2 * beta_l:
            alpha = R(1)
        elif beta_l > par * beta_f:
            alpha = R(beta_l / (par2 * nmL))
        g1 = R(R(1.99) / beta_f) if beta_f > 0 else R(R(1) / R(nmL / alpha))
        g2 = R(R(0.99) / R(beta_l / R(2) + g1 * R(nmL * nmL)))
    elif _approx(theta, 0, R) and _approx(mu, 1, R):                           # PPCA :362-385
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
    elif _approx(mu, 0, R):                                                    # SDCA & PDCA :386-408
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
    elif _approx(theta, 0, R) and _approx(mu, 0.5, R):                         # PPDCA :409-422
        if _approx(beta_l, 0, R) or _approx(beta_f, 0, R):
            alpha = pick()
        else:
            alpha = R(np.sqrt(R(beta_l / beta_f)) / R(2))
        g1 = R(R(1) / R(beta_f / R(2) + nmL / alpha))
        g2 = R(R(0.99) / R(beta_l / R(2) + nmL * alpha))
    else:
        raise ValueError("this choice of theta and mu is not supported!")      # :424
    return g1, g2


def conj_prox(h, v, gamma):
    """prox!(y, convex_conjugate(h), v, gamma) by the Moreau identity (value not needed by AFBA).  Zero* = IndZero -> 0."""
    R = _R(v)
    if isinstance(h, ZeroFn):
        return np.zeros_like(v)
    p, _ = h.prox((v / R(gamma)).astype(v.dtype), R(R(1) / R(gamma)))
    return (v - R(gamma) * p).astype(v.dtype)


class SqrNormL2Smooth:
    """AutoDifferentiable(SqrNormL2(lam)) of test_elasticnet.jl:69: f = lam/2*||x||^2, gradient lam*x."""

    def __init__(self, lam=1.0):
        self.lam = lam

    def value_and_gradient(self, x):
        R = _R(x)
        return R(R(self.lam) / R(2) * np.sum(x * x, dtype=R)), (R(self.lam) * x).astype(x.dtype)


class LinearSmooth:
    """f(x) = <c, x> (test/problems/test_linear_programs.jl:104: AutoDifferentiable(x -> dot(c, x)))."""

    def __init__(self, c):
        self.c = c

    def value_and_gradient(self, x):
        return _R(x)(self.c @ x), self.c.copy()


class AFBAIteration:
    """primal_dual.jl:57-111.  L: None = I (or 0*I when h is Zero, :62-66), else a matrix."""

    def __init__(self, x0, y0, f=None, g=None, h=None, L=None, beta_f=None, theta=1, mu=1, lambda_=1, gamma=None):
        R = _R(x0)
        self.R, self.x0, self.y0 = R, x0, y0
        self.f = f if f is not None else ZeroFn()
        self.g = g if g is not None else ZeroFn()
        self.h = h if h is not None else ZeroFn()
        self.h_zero = isinstance(self.h, ZeroFn)
        self.L = L
        if beta_f is None:
            if not isinstance(self.f, ZeroFn):
                raise ValueError("argument beta_f must be specified together with f")   # :70-74
            beta_f = 0
        self.theta, self.mu, self.lambda_ = R(theta), R(mu), R(lambda_)
        if gamma is None:
            if self.lambda_ != 1:
                raise ValueError("if lambda != 1, then you need to provide stepsizes manually")   # :105-106
            nmL = (0.0 if self.h_zero else 1.0) if L is None else np.linalg.norm(np.asarray(L, np.float64), 2)
            gamma = default_stepsizes(nmL, self.h_zero, theta, mu, beta_f, 0, R)
        self.gamma = (R(gamma[0]), R(gamma[1]))

    def _L(self, x):
        if self.L is None:
            return np.zeros_like(self.y0) if self.h_zero else x.copy()
        return (np.asfortranarray(self.L) @ x).astype(x.dtype)

    def _Lt(self, y):
        if self.L is None:
            return np.zeros_like(self.x0) if self.h_zero else y.copy()
        return (np.asfortranarray(self.L).T @ y).astype(y.dtype)

    def step(self, st=None):                                                   # :174-211
        R = self.R
        g1, g2 = self.gamma
        if st is None:
            st = State(x=self.x0.copy(), y=self.y0.copy())
        _, gradf = self.f.value_and_gradient(st.x)                             # :179-180
        t = self._Lt(st.y)                                                     # :181
        t = (t + gradf).astype(t.dtype)                                        # :182
        t = (t * R(-g1)).astype(t.dtype)                                       # :183
        t = (t + st.x).astype(t.dtype)                                         # :184
        st.xbar, _ = self.g.prox(t, g1)                                        # :185
        tx = (self.theta * st.xbar + R(R(1) - self.theta) * st.x).astype(st.x.dtype)   # :190
        ty = self._L(tx)                                                       # :191  (gradl = 0 for l = IndZero, :188-189,192)
        ty = (ty * g2).astype(ty.dtype)                                        # :193
        ty = (ty + st.y).astype(ty.dtype)                                      # :194
        st.ybar = conj_prox(self.h, ty, g2)                                    # :195
        st.FPR_x = (st.xbar - st.x).astype(st.x.dtype)                         # :198-199
        st.FPR_y = (st.ybar - st.y).astype(st.y.dtype)
        c1 = R(R(self.mu * R(R(2) - self.theta)) * g1)                         # :202
        tx = self._Lt((c1 * st.FPR_y).astype(st.y.dtype))                      # :203
        st.x = (st.x + self.lambda_ * (st.FPR_x - tx).astype(st.x.dtype)).astype(st.x.dtype)   # :204
        c2 = R(R(R(R(1) - self.mu) * R(R(2) - self.theta)) * g2)               # :207
        ty = self._L((c2 * st.FPR_x).astype(st.x.dtype))                       # :208
        st.y = (st.y + self.lambda_ * (st.FPR_y + ty).astype(st.y.dtype)).astype(st.y.dtype)   # :209
        return st

    def __iter__(self):
        st = self.step(None)
        while True:
            yield st
            st = self.step(st)


def afba(x0, y0, maxit=10_000, tol=1e-5, **kw):
    """AFBA(...)(...) -> ((xbar, ybar), k); stop: norm(FPR_x, Inf) + norm(FPR_y, Inf) <= tol (:213-215)."""
    it = AFBAIteration(x0, y0, **kw)
    R = it.R
    for k, st in enumerate(it, start=1):
        if k >= maxit or float(R(norm_inf(st.FPR_x) + norm_inf(st.FPR_y))) <= float(tol):
            return (st.xbar, st.ybar), k


def vu_condat(x0, y0, **kw):
    return afba(x0, y0, theta=2, **kw)                                         # :139


def chambolle_pock(x0, y0, **kw):
    return afba(x0, y0, theta=2, f=None, **kw)                                 # :158-159, :331
