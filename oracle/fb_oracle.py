"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY.  Not part of the product path.

A numpy/BLAS restatement of the ONE hot path of ProximalAlgorithms.jl v0.7.0 that this repository
accelerates: the ForwardBackward / FastForwardBackward (FISTA) iteration with its backtracking line
search, Nesterov sequences and the driver loop.  Only `tests/`, `__graft_entry__.smoke()` and the
`cpu_baseline` / `--impl reference` legs of `bench.py` may import this module.

Every function cites the reference lines it follows (paths relative to /root/reference).  The reference
is pure Julia and Julia is not installed in this image, so the reference itself cannot be executed here.
Pinning (see tests/test_oracle_golden.py):
  * NormL1 + LeastSquares + FB/FFB: pinned against the reference's own known answers -- the literal `x_star`
    and every iteration bound of test/problems/test_lasso_small.jl:42,46-135 and
    test_lasso_small_strongly_convex.jl:14-20,65-144, the `xstar` vectors carried by
    benchmark/data/lasso_{tiny,small,medium}.jld2, and the Nesterov identities of test/accel/test_nesterov.jl:63-81.
  * IndBox: pinned only by the inline clamp of test/problems/test_nonconvex_qp.jl:33.
  * IndBallL2, NormL21: **parity unpinned** -- they appear in no reference test.  Their arithmetic lives in the
    third-party package ProximalOperators.jl (compat "0.15", benchmark/Project.toml:10-11; no Manifest is
    committed, so the exact version is unpinned) and is restated here from that package's published algorithm.

Rounding idioms of Julia that are mimicked on purpose:
  * `norm(v)^2` is sqrt-then-square (fb_tools.jl:4, benchmark/benchmarks.jl:16); `norm` of a dense BLAS-float
    vector is BLAS nrm2 for length >= 32 and a Float64 sequential sum of squares below that (LinearAlgebra
    generic_norm2), `dot` is BLAS dot.
  * all scalars live in R = real(eltype(x0)); `R(sqrt(length(x)))` takes the sqrt in Float64 first (fb_tools.jl:11).
  * `A*x` / `A'*r` are BLAS gemv on a column-major matrix (numpy dispatches F-ordered operands to the same gemv).
"""
from __future__ import annotations

import itertools
import math

import numpy as np
from scipy.linalg import blas as _blas

# --------------------------------------------------------------------------------------------------
# L-1  array substrate idioms (Julia LinearAlgebra behaviour, not reference code)
# --------------------------------------------------------------------------------------------------


def _R(x):
    """real(eltype(x)) as a numpy scalar type."""
    return np.float32 if x.dtype == np.float32 else np.float64


def norm2(v):
    """Julia `norm(v)`: BLAS nrm2 when length >= 32 else Float64 sequential sum of squares (generic_norm2)."""
    R = _R(v)
    n = v.shape[0]
    if n == 0:
        return R(0)
    if n >= 32:
        f = _blas.snrm2 if R is np.float32 else _blas.dnrm2
        return R(f(v))
    s = 0.0
    for e in v.tolist():  # python floats are Float64; v.tolist() widens float32 exactly
        s += e * e
    return R(math.sqrt(s))


def norm_inf(v):
    R = _R(v)
    return R(np.max(np.abs(v))) if v.size else R(0)


def dot(a, b):
    """Julia `dot` on BLAS floats -> BLAS dot."""
    R = _R(a)
    f = _blas.sdot if R is np.float32 else _blas.ddot
    return R(f(a, b))


# --------------------------------------------------------------------------------------------------
# L0  first-order oracles: smooth terms
# --------------------------------------------------------------------------------------------------


class LeastSquares:
    """f(x) = 0.5*||A x - b||^2 with the benchmark's explicit gradient (benchmark/benchmarks.jl:11-17):
    `res = A*x - b; (norm(res)^2/2, A'*res)` -- the gradient is a fresh array every call."""

    is_generalized_quadratic = True   # ProximalOperators' trait for LeastSquares (read by panoc.jl:217)

    def __init__(self, A, b):
        self.A = np.asfortranarray(A)
        self.b = np.ascontiguousarray(b)
        self.calls = 0

    def value_and_gradient(self, x):
        self.calls += 1
        R = _R(x)
        res = self.A @ x - self.b
        nr = norm2(res)
        return R(nr * nr / R(2)), self.A.T @ res

    def __call__(self, x):
        return self.value_and_gradient(x)[0]


class BlockDiagLeastSquares:
    """f(x) = 0.5*||A x - b||^2 with A = blockdiag(A_1..A_B), A_i in R^{mb x nb} ("implicit A via batched GEMV",
    BASELINE.json configs[1]; structure fixed in SURVEY.md section 7 hard part 6).  `blocks` has shape (B, mb, nb)."""

    def __init__(self, blocks, b):
        self.blocks = blocks
        self.b = b
        self.nblk, self.mb, self.nb = blocks.shape

    def value_and_gradient(self, x):
        R = _R(x)
        xb = x.reshape(self.nblk, self.nb)
        res = np.einsum("bij,bj->bi", self.blocks, xb).reshape(-1) - self.b
        nr = norm2(res)
        grad = np.einsum("bij,bi->bj", self.blocks, res.reshape(self.nblk, self.mb)).reshape(-1)
        return R(nr * nr / R(2)), grad.astype(x.dtype, copy=False)


class SquaredDistance:
    """benchmark/benchmarks.jl:19-28: f(x) = norm(x-b)^2/2, gradient x-b."""

    def __init__(self, b):
        self.b = b

    def value_and_gradient(self, x):
        R = _R(x)
        d = x - self.b
        nr = norm2(d)
        return R(nr * nr / R(2)), d


class Quadratic:
    """test/runtests.jl:6-16 fixture: f(x) = 0.5 x'Qx + q'x; gradient Qx + q."""

    def __init__(self, Q, q):
        self.Q, self.q = Q, q

    def value_and_gradient(self, x):
        R = _R(x)
        g = self.Q @ x
        return R(dot(x, g) / R(2) + dot(x, self.q)), g + self.q


class ZeroFn:
    """ProximalCore.Zero: value 0, gradient zero(x) (src/ProximalAlgorithms.jl:38-40); prox is the identity."""

    def value_and_gradient(self, x):
        return _R(x)(0), np.zeros_like(x)

    def prox(self, y, gamma):
        return y.copy(), _R(y)(0)


# --------------------------------------------------------------------------------------------------
# L0  first-order oracles: proximable terms (ProximalOperators.jl 0.15 published algorithms, restated)
#     call sites in the reference: fast_forward_backward.jl:80,141; forward_backward.jl:72,118; fb_tools.jl:49
# --------------------------------------------------------------------------------------------------


class NormL1:
    """g(x) = lam*||x||_1.  prox: z_i = y_i + (y_i <= -gl ? gl : (y_i >= gl ? -gl : -y_i)), gl = gamma*lam;
    returns lam*sum|z_i|.  Pinned by x_star of test/problems/test_lasso_small.jl:42."""

    def __init__(self, lam=1.0):
        self.lam = lam

    def prox(self, y, gamma):
        R = _R(y)
        gl = R(gamma) * R(self.lam)
        z = y + np.where(y <= -gl, gl, np.where(y >= gl, -gl, -y)).astype(y.dtype)
        return z, R(R(self.lam) * np.sum(np.abs(z), dtype=R))

    def __call__(self, x):
        R = _R(x)
        return R(R(self.lam) * np.sum(np.abs(x), dtype=R))


class IndBox:
    """Indicator of [lo, hi]^n (scalar or per-element bounds).  prox = clamp, value 0.
    Pinned by `min.(upp, max.(low, .))` at test/problems/test_nonconvex_qp.jl:33."""

    def __init__(self, lo, hi):
        self.lo, self.hi = lo, hi

    def prox(self, y, gamma):
        R = _R(y)
        lo = np.asarray(self.lo, dtype=y.dtype)
        hi = np.asarray(self.hi, dtype=y.dtype)
        z = np.where(y < lo, lo, np.where(y > hi, hi, y)).astype(y.dtype)
        return z, R(0)


class IndBallL2:
    """Indicator of {||x||_2 <= r}.  prox: scal = r/norm(y); z = y if scal > 1 else scal*y.  value 0.  PARITY UNPINNED."""

    def __init__(self, r=1.0):
        self.r = r

    def prox(self, y, gamma):
        R = _R(y)
        ny = norm2(y)
        with np.errstate(divide="ignore"):
            scal = R(self.r) / ny
        if scal > 1:
            return y.copy(), R(0)
        return (scal * y).astype(y.dtype), R(0)


class NormL21:
    """g(X) = lam * sum_j ||X[:,j]||_2 over contiguous groups of `group` elements (Julia dim=1 on a group x ngroups
    column-major matrix).  prox per group: ns = sqrt(sum y^2); scal = max(0, 1 - gl/ns); z = scal*y;
    value lam*sum(scal*ns).  PARITY UNPINNED."""

    def __init__(self, lam, group):
        self.lam, self.group = lam, int(group)

    def prox(self, y, gamma):
        R = _R(y)
        gl = R(gamma) * R(self.lam)
        yg = y.reshape(-1, self.group)
        ns = np.sqrt(np.sum(yg * yg, axis=1, dtype=R)).astype(R)
        with np.errstate(divide="ignore", invalid="ignore"):
            scal = (R(1) - gl / ns).astype(R)
        scal = np.where(scal <= 0, R(0), scal).astype(R)  # NaN (ns==0 and gl==0) stays NaN as in the package
        z = (scal[:, None] * yg).astype(y.dtype).reshape(-1)
        return z, R(R(self.lam) * np.sum(scal * ns, dtype=R))


def prox(g, y, gamma):
    """ProximalCore.prox(g, y, gamma) -> (z, g(z)) (allocating form)."""
    return g.prox(y, gamma)


# --------------------------------------------------------------------------------------------------
# L1  step utilities  (src/utilities/fb_tools.jl)
# --------------------------------------------------------------------------------------------------


def f_model(f_x, grad_f_x, res, L):
    """fb_tools.jl:3-5   f_x - real(dot(grad, res)) + (L/2)*norm(res)^2, all in R."""
    R = _R(res)
    nr = norm2(res)
    return R(R(f_x) - dot(grad_f_x, res) + R(R(L) / R(2)) * R(nr * nr))


def lower_bound_smoothness_constant(f, x, grad_f_x):
    """fb_tools.jl:7-12 with A = I (as called from forward_backward.jl:68-70, fast_forward_backward.jl:76-78)."""
    R = _R(x)
    xeps = x + R(1)
    _, grad_eps = f.value_and_gradient(xeps)
    return R(norm2(grad_eps - grad_f_x) / R(math.sqrt(x.shape[0])))


class BacktrackResult:
    __slots__ = ("gamma", "g_z", "f_z", "f_z_upp", "grad_f_z", "trials", "warned")


def backtrack_stepsize(gamma, f, g, x, f_x, grad_f_x, y, z, g_z, res, want_grad,
                       minimum_gamma, reduce_gamma):
    """fb_tools.jl:24-63 with A = nothing, Az aliased to z (the only way FB/FFB call it).  Mutates y, z, res in place."""
    R = _R(x)
    alpha = R(1)
    eps = R(np.finfo(R).eps)
    out = BacktrackResult()
    out.trials = 0
    f_upp = f_model(f_x, grad_f_x, res, R(alpha / gamma))                      # :42
    f_z, grad_tmp = f.value_and_gradient(z)                                    # :43-44
    tol = R(R(10) * eps * R(R(1) + abs(f_z)))                                  # :45
    while f_z > R(f_upp + tol) and gamma >= minimum_gamma:                     # :46
        gamma = R(gamma * reduce_gamma)                                        # :47
        y[...] = x - gamma * grad_f_x                                          # :48
        z_new, g_z = g.prox(y, gamma)                                          # :49
        z[...] = z_new
        res[...] = x - z                                                       # :50
        f_upp = f_model(f_x, grad_f_x, res, R(alpha / gamma))                  # :51
        f_z, grad_tmp = f.value_and_gradient(z)                                # :52-53
        tol = R(R(10) * eps * R(R(1) + abs(f_z)))                              # :54
        out.trials += 1
    out.warned = bool(gamma < minimum_gamma)                                   # :59-61
    out.gamma, out.g_z, out.f_z, out.f_z_upp = gamma, g_z, f_z, f_upp
    out.grad_f_z = grad_tmp if want_grad else None                             # :56-58
    return out


# --------------------------------------------------------------------------------------------------
# L1  extrapolation sequences  (src/accel/nesterov.jl)
# --------------------------------------------------------------------------------------------------


def fixed_nesterov_sequence(R):
    """nesterov.jl:14-17   t+ = (1+sqrt(1+4t^2))/2 ; beta = (t-1)/t+ ; t0 = 1."""
    t = R(1)
    while True:
        t_next = R((R(1) + np.sqrt(R(R(1) + R(4) * R(t * t)))) / R(2))
        yield R((t - R(1)) / t_next)
        t = t_next


def simple_nesterov_sequence(R):
    """nesterov.jl:36   R(k-1)/(k+2), k = 1,2,..."""
    for k in itertools.count(1):
        yield R(R(k - 1) / R(k + 2))


def constant_nesterov_sequence(m, stepsize):
    """nesterov.jl:51-54   repeated((1-sqrt(m*gamma))/(1+sqrt(m*gamma)))."""
    R = type(m)
    k_inv = R(m * stepsize)
    val = R((R(1) - np.sqrt(k_inv)) / (R(1) + np.sqrt(k_inv)))
    return itertools.repeat(val)


class AdaptiveNesterovSequence:
    """nesterov.jl:56-60,80,89-103."""

    def __init__(self, m):
        self.R = type(m)
        self.m = m
        self.stepsize = self.R(-1)
        self.theta = self.R(-1)

    def next(self, stepsize):
        R = self.R
        stepsize = R(stepsize)
        if self.stepsize < 0:                                                  # :90-93
            self.stepsize = stepsize
            self.theta = R(np.sqrt(R(self.m * stepsize))) if self.m > 0 else R(1)
        th2 = R(self.theta * self.theta)
        b = R(R(th2 / self.stepsize) - self.m)                                 # :94
        delta = R(R(b * b) + R(R(R(4) * th2) / R(self.stepsize * stepsize)))   # :95
        theta = R(R(stepsize * R(-b + np.sqrt(delta))) / R(2))                 # :96
        beta = R(R(R(stepsize * self.theta) * R(R(1) - self.theta))
                 / R(R(self.stepsize * theta) + R(stepsize * th2)))            # :97-99
        self.stepsize = stepsize
        self.theta = theta
        return beta


# --------------------------------------------------------------------------------------------------
# L2  iterators and states
# --------------------------------------------------------------------------------------------------


class State:
    """Field names follow ForwardBackwardState (forward_backward.jl:52-63) / FastForwardBackwardState
    (fast_forward_backward.jl:60-71)."""

    def __init__(self, **kw):
        self.__dict__.update(kw)


class ForwardBackwardIteration:
    """forward_backward.jl:38-48 (parameters), :65-84 (init), :86-123 (step)."""

    def __init__(self, x0, f=None, g=None, Lf=None, gamma=None, adaptive=None,
                 minimum_gamma=1e-7, reduce_gamma=0.5, increase_gamma=1.0):
        R = _R(x0)
        self.R = R
        self.x0 = x0
        self.f = f if f is not None else ZeroFn()
        self.g = g if g is not None else ZeroFn()
        self.Lf = Lf
        if gamma is None and Lf is not None:
            gamma = 1 / Lf                                                     # :43
        self.gamma = gamma
        self.adaptive = (gamma is None) if adaptive is None else adaptive      # :44
        self.minimum_gamma = R(minimum_gamma)
        self.reduce_gamma = R(reduce_gamma)
        self.increase_gamma = R(increase_gamma)
        self.backtracks = 0

    def _init_common(self):
        R = self.R
        x = self.x0.copy()                                                     # :66
        f_x, grad_f_x = self.f.value_and_gradient(x)                           # :67
        if self.gamma is None:                                                 # :68-70
            gamma = R(R(1) / lower_bound_smoothness_constant(self.f, x, grad_f_x))
        else:
            gamma = R(self.gamma)
        y = x - gamma * grad_f_x                                               # :71
        z, g_z = self.g.prox(y, gamma)                                         # :72
        return State(x=x, f_x=R(f_x), grad_f_x=grad_f_x.copy(), gamma=gamma, y=y, z=z, g_z=g_z, res=x - z)

    def init(self):
        st = self._init_common()
        st.grad_f_z = np.empty_like(st.x)                                      # :62
        return st

    def step(self, st):
        R = self.R
        if self.adaptive:                                                      # :90-110
            st.gamma = R(st.gamma * self.increase_gamma)
            bt = backtrack_stepsize(st.gamma, self.f, self.g, st.x, st.f_x, st.grad_f_x, st.y, st.z, st.g_z,
                                    st.res, True, self.minimum_gamma, self.reduce_gamma)
            self.backtracks += bt.trials
            st.gamma, st.g_z, st.f_x = bt.gamma, bt.g_z, bt.f_z
            st.grad_f_z[...] = bt.grad_f_z
            st.x, st.z = st.z, st.x
            st.grad_f_x, st.grad_f_z = st.grad_f_z, st.grad_f_x
        else:                                                                  # :111-115
            st.x, st.z = st.z, st.x
            st.f_x, grad = self.f.value_and_gradient(st.x)
            st.grad_f_x[...] = grad
        st.y[...] = st.x - st.gamma * st.grad_f_x                              # :117
        z_new, st.g_z = self.g.prox(st.y, st.gamma)                            # :118
        st.z[...] = z_new
        st.res[...] = st.x - st.z                                              # :120
        return st

    def __iter__(self):
        st = self.init()
        while True:
            yield st
            st = self.step(st)


class FastForwardBackwardIteration(ForwardBackwardIteration):
    """fast_forward_backward.jl:44-56 (parameters), :73-97 (init), :99-104 (beta dispatch), :106-145 (step)."""

    def __init__(self, x0, f=None, g=None, mf=0, Lf=None, gamma=None, adaptive=None, minimum_gamma=1e-7,
                 reduce_gamma=0.5, increase_gamma=1.0, extrapolation_sequence=None):
        super().__init__(x0, f, g, Lf, gamma, adaptive, minimum_gamma, reduce_gamma, increase_gamma)
        self.mf = self.R(mf)
        self.extrapolation_sequence = extrapolation_sequence

    def init(self):
        st = self._init_common()                                               # :74-89
        st.z_prev = st.x.copy()                                                # :69 default
        if self.extrapolation_sequence is not None:                            # :90-94
            st.extrapolation_sequence = iter(self.extrapolation_sequence)
        else:
            st.extrapolation_sequence = AdaptiveNesterovSequence(self.mf)
        return st

    def step(self, st):
        R = self.R
        if self.adaptive:                                                      # :110-129
            st.gamma = R(st.gamma * self.increase_gamma)
            bt = backtrack_stepsize(st.gamma, self.f, self.g, st.x, st.f_x, st.grad_f_x, st.y, st.z, st.g_z,
                                    st.res, False, self.minimum_gamma, self.reduce_gamma)
            self.backtracks += bt.trials
            st.gamma, st.g_z = bt.gamma, bt.g_z
        else:
            st.gamma = R(self.gamma)                                           # :130-132
        seq = st.extrapolation_sequence                                        # :99-104, :134
        beta = seq.next(st.gamma) if isinstance(seq, AdaptiveNesterovSequence) else R(next(seq))
        st.beta = beta
        st.x[...] = st.z + beta * (st.z - st.z_prev)                           # :135
        st.z_prev, st.z = st.z, st.z_prev                                      # :136
        st.f_x, grad = self.f.value_and_gradient(st.x)                         # :138
        st.grad_f_x[...] = grad                                                # :139
        st.y[...] = st.x - st.gamma * st.grad_f_x                              # :140
        z_new, st.g_z = self.g.prox(st.y, st.gamma)                            # :141
        st.z[...] = z_new
        st.res[...] = st.x - st.z                                              # :142
        return st


# --------------------------------------------------------------------------------------------------
# L3  driver loop  (src/ProximalAlgorithms.jl:114-123) with the default stop rule
#     norm(res, Inf)/gamma <= tol  (forward_backward.jl:125-126, fast_forward_backward.jl:147-152)
# --------------------------------------------------------------------------------------------------


def default_stop(tol, st):
    return float(norm_inf(st.res) / st.gamma) <= float(tol)


def run(iteration, maxit=10_000, tol=1e-8, stop=None, trace=None):
    """Returns (solution = state.z, k).  k counts the init state as iteration 1."""
    for k, st in enumerate(iteration, start=1):
        if trace is not None:
            trace(k, st)
        if k >= maxit or (stop(iteration, st) if stop else default_stop(tol, st)):
            return st.z, k


def forward_backward(x0, f, g, maxit=10_000, tol=1e-8, **kw):
    return run(ForwardBackwardIteration(x0, f, g, **kw), maxit, tol)


def fast_forward_backward(x0, f, g, maxit=10_000, tol=1e-8, **kw):
    return run(FastForwardBackwardIteration(x0, f, g, **kw), maxit, tol)


# --------------------------------------------------------------------------------------------------
# single fused-step restatements used as elementwise checkers by the GPU parity tests
# --------------------------------------------------------------------------------------------------


def fb_step_unfused(x, grad, gamma, g):
    """y = x - gamma*grad; z = prox(y); res = x - z   (forward_backward.jl:117-120), separate mul/add roundings."""
    R = _R(x)
    y = x - R(gamma) * grad
    z, g_z = g.prox(y, R(gamma))
    return y, z, x - z, g_z


def ffb_step_unfused(x, grad, z_prev, gamma, beta, g):
    """fb_step followed by the NEXT iteration's extrapolation x+ = z + beta*(z - z_prev)
    (fast_forward_backward.jl:135), which the fused kernel performs in the same pass."""
    R = _R(x)
    y, z, res, g_z = fb_step_unfused(x, grad, gamma, g)
    x_next = z + R(beta) * (z - z_prev)
    return y, z, res, g_z, x_next
