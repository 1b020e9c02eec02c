"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

numpy/BLAS restatement of the two "next" rows of SURVEY.md section 8f that reuse the fused forward-backward step:

  * PANOC (src/algorithms/panoc.jl:41-53 parameters, :57-83 state, :88-112 init, :114-135 direction hooks, :138-255 step,
    :257-259 stop/solution) with its directions L-BFGS (src/accel/lbfgs.jl:5-105) and NoAcceleration
    (src/accel/noaccel.jl:1-5), and the general-`A` form of the step-size tools (src/utilities/fb_tools.jl:3-63);
  * DouglasRachford (src/algorithms/douglas_rachford.jl:31-71).

Only `tests/`, `__graft_entry__.smoke()` and the CPU legs of `bench.py` may import this module.  Every function cites the
reference lines it follows (paths relative to /root/reference).  The Julia reference cannot be executed in this image.

Pinning (tests/test_oracle_panoc.py):
  * L-BFGS: the five literal directions of test/accel/test_lbfgs.jl:6-101 (memory 3, 10-dim quadratic), `H*x == x` after reset.
  * PANOC: x_star and the iteration bounds of test/problems/test_lasso_small.jl:159-181 (fixed / adaptive, A given as a
    matrix, f not quadratic), test_lasso_small_strongly_convex.jl:155-162, test_sparse_logistic_small.jl:101-109, the
    fixed-point test of test_nonconvex_qp.jl:28-36,95-103, and the FB/PANOC state equivalence of
    test/problems/test_equivalence.jl:51-83.
  * DouglasRachford: x_star / `it < 30` of test/problems/test_lasso_small.jl:205-214.
  * `LeastSquares.prox` (Cholesky of A'A + I/gamma or, for a wide A, of AA' + I/gamma with the matrix-inversion lemma) and
    `SqrNormL2` / `Translate` restate ProximalOperators.jl 0.15 from its published algorithm: third-party arithmetic that is
    not under /root/reference; pinned only through the DouglasRachford test above.

Julia typing idioms kept on purpose: `0.5 / state.gamma` (panoc.jl:193) is a Float64 literal over an R value, so for
R = Float32 `sigma` and `threshold` are Float64 and the acceptance test is made in Float64; everything else stays in R.
"""
from __future__ import annotations

import math

import numpy as np
from scipy.linalg import cho_factor, cho_solve

from .fb_oracle import State, ZeroFn, _R, dot, f_model, norm2, norm_inf

# --------------------------------------------------------------------------------------------------
# linear maps: `A` of f(Ax) + g(x).  None = the identity (reference default `A = I`, panoc.jl:43)
# --------------------------------------------------------------------------------------------------


def _mul(A, x):
    """A * x (BLAS gemv for a matrix; a copy for I)."""
    return x.copy() if A is None else np.asfortranarray(A) @ x


def _mul_t(A, v):
    """A' * v."""
    return v.copy() if A is None else np.asfortranarray(A).T @ v


# --------------------------------------------------------------------------------------------------
# step-size tools with a general A  (src/utilities/fb_tools.jl)
# --------------------------------------------------------------------------------------------------


def lower_bound_smoothness_constant(f, A, x, grad_f_Ax):
    """fb_tools.jl:7-12."""
    R = _R(x)
    xeps = x + R(1)
    _, grad_eps = f.value_and_gradient(_mul(A, xeps))
    return R(norm2(_mul_t(A, grad_eps - grad_f_Ax)) / R(math.sqrt(x.shape[0])))


def backtrack_stepsize(gamma, f, A, g, x, f_Ax, At_grad_f_Ax, y, z, g_z, res, Az, grad_f_Az, alpha, minimum_gamma,
                       reduce_gamma=0.5):
    """fb_tools.jl:24-63 as PANOC calls it (panoc.jl:143-159): mutates y, z, res, Az, grad_f_Az in place.
    Returns (gamma, g_z, f_Az, f_Az_upp, trials, warned)."""
    R = _R(x)
    eps = R(np.finfo(R).eps)
    reduce_gamma = R(reduce_gamma)
    trials = 0
    f_upp = f_model(f_Ax, At_grad_f_Ax, res, R(alpha / gamma))                 # :42
    Az[...] = _mul(A, z)                                                       # :43
    f_Az, grad_tmp = f.value_and_gradient(Az)                                  # :44
    tol = R(R(10) * eps * R(R(1) + abs(f_Az)))                                 # :45
    while f_Az > R(f_upp + tol) and gamma >= minimum_gamma:                    # :46
        gamma = R(gamma * reduce_gamma)                                        # :47
        y[...] = x - gamma * At_grad_f_Ax                                      # :48
        z_new, g_z = g.prox(y, gamma)                                          # :49
        z[...] = z_new
        res[...] = x - z                                                       # :50
        f_upp = f_model(f_Ax, At_grad_f_Ax, res, R(alpha / gamma))             # :51
        Az[...] = _mul(A, z)                                                   # :52
        f_Az, grad_tmp = f.value_and_gradient(Az)                              # :53
        tol = R(R(10) * eps * R(R(1) + abs(f_Az)))                             # :54
        trials += 1
    grad_f_Az[...] = grad_tmp                                                  # :56-58
    return gamma, g_z, R(f_Az), f_upp, trials, bool(gamma < minimum_gamma)


# --------------------------------------------------------------------------------------------------
# directions  (src/accel/lbfgs.jl, src/accel/noaccel.jl)
# --------------------------------------------------------------------------------------------------


class LBFGSOperator:
    """lbfgs.jl:5-28 (storage), :30-51 (update!), :53-56 (reset!), :66-95 (two-loop product).  Ring indices are 1-based
    like the reference so that the visiting order of the pairs is the same."""

    def __init__(self, M, x):
        R = _R(x)
        self.M, self.R = int(M), R
        self.currmem, self.curridx = 0, 0
        self.s_M = [np.zeros_like(x) for _ in range(self.M)]
        self.y_M = [np.zeros_like(x) for _ in range(self.M)]
        self.ys_M = np.zeros(self.M, dtype=R)
        self.alphas = np.zeros(self.M, dtype=R)
        self.H = R(1)

    def update(self, s, y):
        ys = dot(s, y)                                                         # :33
        if ys > 0:                                                             # :34
            self.curridx += 1                                                  # :35-38
            if self.curridx > self.M:
                self.curridx = 1
            self.currmem = min(self.currmem + 1, self.M)                       # :39-42
            self.ys_M[self.curridx - 1] = ys                                   # :43
            self.s_M[self.curridx - 1][...] = s                                # :44-45
            self.y_M[self.curridx - 1][...] = y
            yty = dot(y, y)                                                    # :46
            self.H = self.R(ys / yty)                                          # :47
        return self

    def reset(self):
        self.currmem, self.curridx = 0, 0                                      # :54
        self.H = self.R(1)                                                     # :55

    def mul(self, v):
        """d = H * v by the two-loop recursion (:66-95)."""
        R = self.R
        d = v.copy()                                                           # :67
        idx = self.curridx
        for _ in range(self.currmem):                                          # loop1 :74-84
            self.alphas[idx - 1] = R(dot(self.s_M[idx - 1], d) / self.ys_M[idx - 1])
            d -= self.alphas[idx - 1] * self.y_M[idx - 1]
            idx -= 1
            if idx == 0:
                idx = self.M
        d *= self.H                                                            # :69
        for _ in range(self.currmem):                                          # loop2 :86-95
            idx += 1
            if idx > self.M:
                idx = 1
            beta = R(dot(self.y_M[idx - 1], d) / self.ys_M[idx - 1])
            d += R(self.alphas[idx - 1] - beta) * self.s_M[idx - 1]
        return d


class LBFGS:
    """lbfgs.jl:97-105."""

    def __init__(self, M=5):
        self.M = int(M)

    def initialize(self, x):
        return LBFGSOperator(self.M, x)


class NoAcceleration:
    """noaccel.jl:1-5: direction = -res, no state."""

    def initialize(self, x):
        return None


# --------------------------------------------------------------------------------------------------
# PANOC  (src/algorithms/panoc.jl)
# --------------------------------------------------------------------------------------------------


def _is_quadratic(f):
    """ProximalCore.is_generalized_quadratic(Tf) (panoc.jl:217): a trait of the TYPE of f; False unless declared."""
    return bool(getattr(f, "is_generalized_quadratic", False))


class PANOCIteration:
    """panoc.jl:41-53.  `A=None` is the identity."""

    def __init__(self, x0, f=None, A=None, g=None, alpha=0.95, beta=0.5, Lf=None, gamma=None, adaptive=None,
                 minimum_gamma=1e-7, max_backtracks=20, directions=None):
        R = _R(x0)
        self.R, self.x0 = R, x0
        self.f = f if f is not None else ZeroFn()
        self.A = A
        self.g = g if g is not None else ZeroFn()
        self.alpha, self.beta = R(alpha), R(beta)
        self.Lf = Lf
        if gamma is None and Lf is not None:
            gamma = R(self.alpha / R(Lf))                                      # :49
        self.gamma = gamma
        self.adaptive = (gamma is None) if adaptive is None else bool(adaptive)   # :50
        self.minimum_gamma = R(minimum_gamma)
        self.max_backtracks = int(max_backtracks)
        self.directions = directions if directions is not None else LBFGS(5)
        self.backtracks = 0          # gamma halvings (fb_tools.jl:46-55)
        self.tau_backtracks = 0      # line-search halvings (panoc.jl:203-250)

    def f_model(self, st):
        """panoc.jl:85-86."""
        return f_model(st.f_Ax, st.At_grad_f_Ax, st.res, self.R(self.alpha / st.gamma))

    def init(self):                                                            # :88-112
        R = self.R
        x = self.x0.copy()
        Ax = _mul(self.A, x)
        f_Ax, grad_f_Ax = self.f.value_and_gradient(Ax)
        if self.gamma is None:
            gamma = R(self.alpha / lower_bound_smoothness_constant(self.f, self.A, x, grad_f_Ax))
        else:
            gamma = R(self.gamma)
        At_grad = _mul_t(self.A, grad_f_Ax)
        y = x - gamma * At_grad
        z, g_z = self.g.prox(y, gamma)
        like_x = lambda: np.empty_like(x)      # noqa: E731
        like_Ax = lambda: np.empty_like(Ax)    # noqa: E731
        return State(x=x, Ax=Ax, f_Ax=R(f_Ax), grad_f_Ax=grad_f_Ax.copy(), At_grad_f_Ax=At_grad, gamma=gamma, y=y, z=z,
                     g_z=g_z, res=x - z, H=self.directions.initialize(x), tau=R(0), x_prev=like_x(), res_prev=like_x(),
                     d=like_x(), Ad=like_Ax(), x_d=like_x(), Ax_d=like_Ax(), f_Ax_d=R(0), grad_f_Ax_d=like_Ax(),
                     At_grad_f_Ax_d=like_x(), z_curr=like_x(), Az=like_Ax(), grad_f_Az=like_Ax(), At_grad_f_Az=like_x())

    def _fb(self, st):
        """y, z, res, g_z from x and At_grad_f_Ax (:199-201, :246-248); returns FBE = f_model + g_z."""
        st.y[...] = st.x - st.gamma * st.At_grad_f_Ax
        z_new, st.g_z = self.g.prox(st.y, st.gamma)
        st.z[...] = z_new
        st.res[...] = st.x - st.z
        return self.R(self.f_model(st) + st.g_z)

    def step(self, st):                                                        # :138-255
        R = self.R
        inf = R(np.inf)
        f_Az, a, b, c = inf, inf, inf, inf
        if self.adaptive:                                                      # :141-164
            gamma_prev = st.gamma
            st.gamma, st.g_z, f_Az, f_Az_upp, trials, _ = backtrack_stepsize(
                st.gamma, self.f, self.A, self.g, st.x, st.f_Ax, st.At_grad_f_Ax, st.y, st.z, st.g_z, st.res, st.Az,
                st.grad_f_Az, self.alpha, self.minimum_gamma)
            self.backtracks += trials
            if st.gamma != gamma_prev and st.H is not None:
                st.H.reset()
        else:
            f_Az_upp = self.f_model(st)                                        # :166
        FBE_x = R(f_Az_upp + st.g_z)                                           # :170
        if st.H is not None:                                                   # :173, :114-121
            st.d[...] = st.H.mul(st.res)
            st.d *= R(-1)
        else:
            st.d[...] = -st.res
        st.x_prev[...] = st.x                                                  # :176-177
        st.res_prev[...] = st.res
        st.tau = R(1)                                                          # :180
        st.Ad[...] = _mul(self.A, st.d)                                        # :181
        st.x_d[...] = st.x + st.d                                              # :183-184
        st.Ax_d[...] = st.Ax + st.Ad
        st.f_Ax_d, grad = self.f.value_and_gradient(st.Ax_d)                   # :185-187
        st.f_Ax_d = R(st.f_Ax_d)
        st.grad_f_Ax_d[...] = grad
        st.At_grad_f_Ax_d[...] = _mul_t(self.A, st.grad_f_Ax_d)
        st.x[...] = st.x_d                                                     # :189-194
        st.Ax[...] = st.Ax_d
        st.grad_f_Ax[...] = st.grad_f_Ax_d
        st.At_grad_f_Ax[...] = st.At_grad_f_Ax_d
        st.z_curr[...] = st.z
        st.f_Ax = st.f_Ax_d
        # :196-198 -- `0.5 / gamma` promotes to Float64 for R = Float32 (Julia literal), so do sigma and threshold
        sigma = np.float64(self.beta) * (np.float64(0.5) / np.float64(st.gamma)) * np.float64(R(R(1) - self.alpha))
        tol = R(R(10) * R(np.finfo(R).eps) * R(R(1) + abs(FBE_x)))
        nr = norm2(st.res)
        threshold = np.float64(FBE_x) - sigma * np.float64(R(nr * nr)) + np.float64(tol)
        FBE_new = self._fb(st)                                                 # :199-202
        st.line_search_trace = (float(FBE_x), float(threshold), float(FBE_new))   # diagnostics: the first test of :205
        for k in range(1, self.max_backtracks + 1):                            # :204-250
            if np.float64(FBE_new) <= threshold:
                break
            if np.isinf(f_Az):                                                 # :209-211
                st.Az[...] = _mul(self.A, st.z_curr)
            st.tau = R(0) if k >= self.max_backtracks else R(st.tau / R(2))    # :213
            one_m = R(R(1) - st.tau)
            st.x[...] = st.tau * st.x_d + one_m * st.z_curr                    # :214-215
            st.Ax[...] = st.tau * st.Ax_d + one_m * st.Az
            if _is_quadratic(self.f):                                          # :217-237
                if np.isinf(f_Az):
                    f_Az, grad = self.f.value_and_gradient(st.Az)
                    f_Az = R(f_Az)
                    st.grad_f_Az[...] = grad
                if np.isinf(c):
                    st.At_grad_f_Az[...] = _mul_t(self.A, st.grad_f_Az)
                    c = f_Az
                    b = R(dot(st.Ax_d, st.grad_f_Az) - dot(st.Az, st.grad_f_Az))
                    a = R(R(st.f_Ax_d - b) - c)
                st.f_Ax = R(R(R(a * R(st.tau * st.tau)) + R(b * st.tau)) + c)
                st.grad_f_Ax[...] = st.tau * st.grad_f_Ax_d + one_m * st.grad_f_Az
                st.At_grad_f_Ax[...] = st.tau * st.At_grad_f_Ax_d + one_m * st.At_grad_f_Az
            else:                                                              # :238-244
                f_Ax, grad = self.f.value_and_gradient(st.Ax)
                st.f_Ax = R(f_Ax)
                st.grad_f_Ax[...] = grad
                st.At_grad_f_Ax[...] = _mul_t(self.A, st.grad_f_Ax)
            FBE_new = self._fb(st)                                             # :246-249
            self.tau_backtracks += 1
        if st.H is not None:                                                   # :252, :123-128
            st.x_prev[...] = st.x - st.x_prev
            st.res_prev[...] = st.res - st.res_prev
            st.H.update(st.x_prev, st.res_prev)
        return st

    def __iter__(self):
        st = self.init()
        while True:
            yield st
            st = self.step(st)


def default_stop(tol, st):
    """panoc.jl:257-258."""
    return float(norm_inf(st.res) / st.gamma) <= float(tol)


def panoc(x0, f=None, A=None, g=None, maxit=1_000, tol=1e-8, trace=None, **kw):
    """PANOC(...)(x0=..., ...) through the driver loop of src/ProximalAlgorithms.jl:114-123.  Returns (state.z, k)."""
    it = PANOCIteration(x0, f, A, g, **kw)
    for k, st in enumerate(it, start=1):
        if trace is not None:
            trace(k, st)
        if k >= maxit or default_stop(tol, st):
            return st.z, k


# --------------------------------------------------------------------------------------------------
# proximable smooth terms used by DouglasRachford (ProximalOperators.jl 0.15, restated; see the module header)
# --------------------------------------------------------------------------------------------------


class LeastSquaresProx:
    """LeastSquares(A, b) as a PROXIMABLE term (test/problems/test_lasso_small.jl:39 `fA_prox`):
    prox_{gamma f}(x) = argmin 0.5*||A w - b||^2 + ||w - x||^2/(2 gamma) = (A'A + I/gamma)^-1 (A'b + x/gamma).
    Tall A: Cholesky of A'A + I/gamma.  Wide A: q = A'b + x/gamma; y = gamma*(q - A'((AA' + I/gamma)^-1 (A q)))."""

    is_generalized_quadratic = True

    def __init__(self, A, b):
        self.A = np.asfortranarray(A)
        self.b = np.ascontiguousarray(b)
        self.Atb = self.A.T @ self.b
        m, n = self.A.shape
        self.tall = m >= n
        self.S = (self.A.T @ self.A) if self.tall else (self.A @ self.A.T)
        self.gamma = None
        self.fact = None

    def value_and_gradient(self, x):
        R = _R(x)
        res = self.A @ x - self.b
        nr = norm2(res)
        return R(nr * nr / R(2)), self.A.T @ res

    def prox(self, x, gamma):
        R = _R(x)
        gamma = R(gamma)
        if self.gamma is None or gamma != self.gamma:
            k = self.S.shape[0]
            self.fact = cho_factor(self.S + np.eye(k, dtype=x.dtype) / gamma, lower=True)
            self.gamma = gamma
        q = self.Atb + x / gamma
        if self.tall:
            y = cho_solve(self.fact, q)
        else:
            y = gamma * (q - self.A.T @ cho_solve(self.fact, self.A @ q))
        y = y.astype(x.dtype, copy=False)
        res = self.A @ y - self.b
        nr = norm2(res)
        return y, R(nr * nr / R(2))


class SqrNormL2Translated:
    """Translate(SqrNormL2(lam), -b) (test/problems/test_lasso_small.jl:38 `f_prox`): f(x) = (lam/2)*||x - b||^2.
    prox: w = x - b; y = w / (1 + gamma*lam); y + b.  value (lam/2)*||w/(1+gamma*lam)||^2.  PARITY UNPINNED."""

    is_generalized_quadratic = True

    def __init__(self, b, lam=1.0):
        self.b, self.lam = b, lam

    def value_and_gradient(self, x):
        R = _R(x)
        d = x - self.b
        nr = norm2(d)
        return R(R(self.lam) / R(2) * R(nr * nr)), (R(self.lam) * d).astype(x.dtype)

    def prox(self, x, gamma):
        R = _R(x)
        gl = R(R(gamma) * R(self.lam))
        w = x - self.b
        y = (w / R(R(1) + gl)).astype(x.dtype)
        val = R(R(self.lam) / R(2) * np.sum(y * y, dtype=R))
        return (y + self.b).astype(x.dtype), val


class LogisticLoss:
    """test/problems/test_sparse_logistic_small.jl:20-26: f(u) = sum(log(1 + exp(-(u - b)))) (labels all one), a smooth
    NON-quadratic term (AutoDifferentiable in the reference: trait false -> general branch of panoc.jl:238-244)."""

    def __init__(self, b):
        self.b = b

    def value_and_gradient(self, u):
        R = _R(u)
        e = np.exp(-(u - self.b))
        return R(np.sum(np.log(R(1) + e), dtype=R)), (-e / (R(1) + e)).astype(u.dtype)


class QuadraticForm:
    """test/problems/test_nonconvex_qp.jl:15-18: f(x) = dot(Q*x, x)/2 + dot(q, x) through AutoDifferentiable (trait false)."""

    def __init__(self, Q, q):
        self.Q, self.q = np.asfortranarray(Q), q

    def value_and_gradient(self, x):
        R = _R(x)
        Qx = self.Q @ x
        return R(dot(Qx, x) / R(2) + dot(self.q, x)), (Qx + self.q).astype(x.dtype)


# --------------------------------------------------------------------------------------------------
# DouglasRachford  (src/algorithms/douglas_rachford.jl)
# --------------------------------------------------------------------------------------------------


class DouglasRachfordIteration:
    """douglas_rachford.jl:31-42 (parameters), :46-52 (state), :54-63 (the one `iterate` method: init and step are the same)."""

    def __init__(self, x0, f=None, g=None, gamma=None):
        if gamma is None:
            raise TypeError("DouglasRachfordIteration: keyword argument `gamma` not assigned")   # :41 has no default
        self.R = _R(x0)
        self.x0 = x0
        self.f = f if f is not None else ZeroFn()
        self.g = g if g is not None else ZeroFn()
        self.gamma = self.R(gamma)

    def step(self, st=None):
        if st is None:
            st = State(x=self.x0.copy())                                       # :56
        st.y, _ = self.f.prox(st.x, self.gamma)                                # :58
        st.r = 2 * st.y - st.x                                                 # :59
        st.z, _ = self.g.prox(st.r, self.gamma)                                # :60
        st.res = st.y - st.z                                                   # :61
        st.x = st.x - st.res                                                   # :62
        return st

    def __iter__(self):
        st = self.step(None)
        while True:
            yield st
            st = self.step(st)


def douglas_rachford(x0, f=None, g=None, gamma=None, maxit=1_000, tol=1e-8):
    """DouglasRachford(...)(...) -> (state.y, k); stop rule norm(res, Inf)/gamma <= tol (:65-70)."""
    it = DouglasRachfordIteration(x0, f, g, gamma)
    for k, st in enumerate(it, start=1):
        if k >= maxit or float(norm_inf(st.res) / it.gamma) <= float(tol):
            return st.y, k
