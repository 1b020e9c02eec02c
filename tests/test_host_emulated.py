"""CPU tests of the solvers' HOST logic (buffer renames, scalar arithmetic in R, line-search control flow) against the
oracle, with the kernels replaced by the numpy emulation of the C ABI in tests/emu_lib.py.  The emulation is test
infrastructure: the product has no CPU path and these tests do not claim kernel parity -- the `-m gpu` tests do that
through the real library."""
import numpy as np
import pytest
import torch

import proxb200 as pa
from oracle import fb_oracle as o
from oracle import panoc_oracle as po
from oracle import afba_oracle as ao
from oracle import tv_oracle as tvo
from proxb200 import algorithms, functions, host

from emu_lib import EmuContext

TYPES = [np.float64, np.float32]


@pytest.fixture()
def emu(monkeypatch):
    ctx = EmuContext()
    monkeypatch.setattr(host.Context, "get", classmethod(lambda cls, device=None: ctx))

    def check_vec(t_, n=None, dtype=None):
        assert t_.is_contiguous() and t_.dim() == 1
        assert n is None or t_.numel() == n
        assert dtype is None or t_.dtype == dtype

    for mod in (host, functions, algorithms):
        monkeypatch.setattr(mod, "check_vec", check_vec)
    from proxb200 import accel

    monkeypatch.setattr(accel, "check_vec", check_vec)
    return ctx


def _lasso_4x5(golden, T):
    d = golden("unit_lasso_4x5")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    lam = T(T(0.1) * np.max(np.abs(A.T @ b)))
    return A, b, lam, d["xstar"].astype(T)


@pytest.mark.parametrize("T", TYPES)
def test_emulated_ffb_matches_oracle(emu, golden, T):
    # sanity of the emulation itself on the already GPU-verified FB/FFB host code
    A, b, lam, xstar = _lasso_4x5(golden, T)
    for adaptive in (False, True):
        kw = {} if adaptive else {"Lf": T(np.linalg.norm(A, 2) ** 2)}
        z_o, it_o = o.fast_forward_backward(np.zeros(5, T), o.LeastSquares(A, b), o.NormL1(lam), tol=T(1e-4), **kw)
        z, it = pa.FastForwardBackward(tol=T(1e-4), driver="python")(x0=np.zeros(5, T), f=pa.LeastSquares(A, b), g=pa.NormL1(lam), **kw)
        assert abs(it - it_o) <= 1 and np.max(np.abs(z - z_o)) <= (1e-9 if T is np.float64 else 1e-4)


@pytest.mark.parametrize("T", TYPES)
def test_lbfgs_operator_known_directions(emu, golden, T):
    # test/accel/test_lbfgs.jl:103-131 through the product's LBFGSOperator wrapper
    d = golden("lbfgs_known_answers")
    Q, q, xs, dirs = (d[k].astype(T) for k in ("Q", "q", "xs", "dirs_ref"))
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a))   # noqa: E731
    H = pa.LBFGS(3).initialize(dev(np.zeros(10, T)))
    x = xs[0]
    grad = Q @ x + q
    rtol = float(np.sqrt(np.finfo(T).eps))
    assert np.allclose(-(H * dev(grad)).numpy(), dirs[0], rtol=rtol)
    for i in range(1, 5):
        x_prev, grad_prev = x, grad
        x = xs[i]
        grad = Q @ x + q
        H.update(dev(x - x_prev), dev(grad - grad_prev))
        out = H.mul(dev(-grad)).numpy()
        assert np.linalg.norm(out - dirs[i]) <= rtol * np.linalg.norm(dirs[i])
    assert H.currmem == 3 and H.curridx == 1
    H.reset()
    assert np.array_equal(H.mul(dev(x)).numpy(), x) and H.currmem == 0


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("form", ["ident_quadratic", "ident_general", "matrix_A"])
@pytest.mark.parametrize("adaptive", [False, True])
def test_panoc_matches_oracle_statewise(emu, golden, T, form, adaptive):
    """Same problem on the oracle and on the product's host logic: identical control flow (backtrack counters) and
    iterates equal up to the rounding of the reductions, iteration by iteration."""
    d = golden("lasso_small")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    n = A.shape[1]
    kw = {} if adaptive else {"Lf": T(np.linalg.norm(A, 2) ** 2)}
    if form == "matrix_A":
        it_o = po.PANOCIteration(np.zeros(n, T), f=o.SquaredDistance(b), A=A, g=o.NormL1(T(1)), **kw)
        it_p = pa.PANOCIteration(np.zeros(n, T), f=pa.SquaredDistance(b), A=A, g=pa.NormL1(T(1)), **kw)
    else:
        fo, fp = o.LeastSquares(A, b), pa.LeastSquares(A, b)
        if form == "ident_general":
            fo.is_generalized_quadratic = False
            fp.is_generalized_quadratic = False
        it_o = po.PANOCIteration(np.zeros(n, T), f=fo, g=o.NormL1(T(1)), **kw)
        it_p = pa.PANOCIteration(np.zeros(n, T), f=fp, g=pa.NormL1(T(1)), **kw)
    steps = 25
    tol = 1e-9 if T is np.float64 else 5e-3
    for k, (so, sp) in enumerate(zip(it_o, it_p)):
        assert float(sp.gamma) == pytest.approx(float(so.gamma), rel=1e-6)
        assert np.max(np.abs(sp.z.numpy() - so.z)) <= tol * max(1.0, np.max(np.abs(so.z))), (k, form)
        assert np.max(np.abs(sp.x.numpy() - so.x)) <= tol * max(1.0, np.max(np.abs(so.x)))
        assert np.max(np.abs(sp.res.numpy() - so.res)) <= tol
        assert float(sp.tau) == float(so.tau), k
        if k == steps:
            break
    assert it_p.tau_backtracks == it_o.tau_backtracks and it_p.backtracks == it_o.backtracks
    assert it_o.tau_backtracks > 0                      # the line search was exercised
    y = sp.y.numpy()
    assert np.allclose(y, sp.x.numpy() - sp.gamma * sp.At_grad_f_Ax.numpy(), rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("T", TYPES)
def test_panoc_reference_bounds(emu, golden, T):
    # test/problems/test_lasso_small.jl:159-181 on the product's host logic
    A, b, lam, xstar = _lasso_4x5(golden, T)
    Lf = T(np.linalg.norm(A, 2) ** 2)
    x0 = np.zeros(5, T)
    x, it = pa.PANOC(tol=T(1e-4))(x0=x0, f=pa.SquaredDistance(b), A=A, g=pa.NormL1(lam), Lf=Lf)
    assert x.dtype == T and np.max(np.abs(x - xstar)) <= 1e-4 and it < 20 and not x0.any()
    x, it = pa.PANOC(adaptive=True, tol=T(1e-4))(x0=x0, f=pa.SquaredDistance(b), A=A, g=pa.NormL1(lam))
    assert np.max(np.abs(x - xstar)) <= 1e-4 and it < 20
    # FB/PANOC equivalence (test_equivalence.jl:51-83)
    gamma = T(T(0.95) / Lf)
    f = pa.LeastSquares(A, b)
    f.is_generalized_quadratic = False
    fb = iter(pa.ForwardBackwardIteration(x0, f=f, g=pa.NormL1(lam), gamma=gamma))
    pn = iter(pa.PANOCIteration(x0, f=f, g=pa.NormL1(lam), gamma=gamma, max_backtracks=1, directions=pa.NoAcceleration()))
    for _ in range(10):
        s1, s2 = next(fb), next(pn)
        assert np.allclose(s1.z.numpy(), s2.z.numpy(), rtol=float(np.sqrt(np.finfo(T).eps)), atol=0)


@pytest.mark.parametrize("T", TYPES)
def test_panoc_user_callbacks_and_box(emu, T):
    # test_nonconvex_qp.jl:9-36 with a user-defined smooth term (value_and_gradient on tensors) and IndBox
    Q, q = np.diag([-0.5, 1.0]).astype(T), np.array([0.3, 0.5], T)

    class Quad:
        def value_and_gradient(self, x):
            xv = x.numpy()
            g = Q @ xv + q
            return T(0.5 * xv @ (Q @ xv) + q @ xv), torch.as_tensor(g.astype(T))

    x, it = pa.PANOC(tol=1e-4)(x0=np.zeros(2, T), f=Quad(), g=pa.IndBox(-1.0, 1.0))
    gamma = 0.95
    z = np.minimum(1.0, np.maximum(-1.0, x - gamma * (Q @ x + q)))
    assert np.max(np.abs(x - z)) / gamma <= 1e-4


@pytest.mark.parametrize("T", TYPES)
def test_douglas_rachford_fused_and_unfused_match_oracle(emu, T):
    rng = np.random.default_rng(11)
    n = 257
    b = rng.standard_normal(n).astype(T)
    x0 = rng.standard_normal(n).astype(T)
    gamma = T(0.7)
    it_o = po.DouglasRachfordIteration(x0, f=po.SqrNormL2Translated(b, 1.5), g=o.NormL1(T(0.3)), gamma=gamma)
    it_p = pa.DouglasRachfordIteration(x0, f=pa.SqrNormL2(1.5, b), g=pa.NormL1(0.3), gamma=gamma)

    class UserL1:                                   # same g through the user `prox_` protocol -> unfused sequence
        def prox_(self, z, y, gam):
            zz, v = o.NormL1(T(0.3)).prox(y.numpy(), gam)
            z.copy_(torch.as_tensor(zz))
            return v

    it_u = pa.DouglasRachfordIteration(x0, f=pa.SqrNormL2(1.5, b), g=UserL1(), gamma=gamma)
    for k, (so, sp, su) in enumerate(zip(it_o, it_p, it_u)):
        for name in ("x", "y", "r", "z", "res"):
            assert np.array_equal(getattr(sp, name).numpy(), getattr(so, name)), (k, name)
            assert np.array_equal(getattr(su, name).numpy(), getattr(so, name)), (k, name)
        assert float(sp.res_norm_inf) == float(np.max(np.abs(so.res)))
        if k == 12:
            break
    y_o, k_o = po.douglas_rachford(x0, f=po.SqrNormL2Translated(b, 1.5), g=o.NormL1(T(0.3)), gamma=gamma, tol=T(1e-5))
    y_p, k_p = pa.DouglasRachford(tol=T(1e-5))(x0=x0, f=pa.SqrNormL2(1.5, b), g=pa.NormL1(0.3), gamma=gamma)
    assert k_p == k_o and np.array_equal(y_p, y_o)
    with pytest.raises(TypeError):
        pa.DouglasRachfordIteration(x0)


@pytest.mark.parametrize("T", TYPES)
def test_douglas_rachford_least_squares_like_the_reference(emu, golden, T):
    # test/problems/test_lasso_small.jl:205-214 through the unfused sequence with the factorised least-squares prox
    A, b, lam, xstar = _lasso_4x5(golden, T)
    gamma = T(T(10) / T(np.linalg.norm(A, 2) ** 2))
    x0 = np.zeros(5, T)
    y, it = pa.DouglasRachford(tol=T(1e-4))(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), gamma=gamma)
    y_o, it_o = po.douglas_rachford(x0, f=po.LeastSquaresProx(A, b), g=o.NormL1(lam), gamma=gamma, tol=T(1e-4))
    assert y.dtype == T and np.max(np.abs(y - xstar)) <= 1e-4 and it < 30 and not x0.any()
    assert abs(it - it_o) <= 1 and np.max(np.abs(y - y_o)) <= (1e-9 if T is np.float64 else 1e-4)


@pytest.mark.parametrize("T", TYPES)
def test_tv_douglas_rachford_host_logic(emu, T):
    rng = np.random.default_rng(3)
    H, W = 10, 12
    b = (np.kron(rng.standard_normal((2, 3)), np.ones((5, 4))) + 0.1 * rng.standard_normal((H, W))).astype(T)
    lam, gamma = 0.25, T(1.0)
    f = pa.TVSplit(b, lam)
    x0 = f.initial_point()
    it_p = iter(pa.DouglasRachfordIteration(x0, f=f, g=pa.IndConsensus(5), gamma=gamma))
    it_o = iter(po.DouglasRachfordIteration(np.tile(b.reshape(-1), 5), f=tvo.TVSplit(b, T(lam), (H, W)), g=tvo.Consensus(5), gamma=gamma))
    for k in range(8):
        sp, so = next(it_p), next(it_o)
        assert np.array_equal(sp.x.numpy(), so.x) and np.array_equal(sp.y.numpy(), so.y)
        assert np.array_equal(sp.z.numpy(), so.z[: H * W])
        assert float(sp.res_norm_inf) == float(np.max(np.abs(so.res)))
    y, k = pa.DouglasRachford(tol=T(1e-4), maxit=3000)(x0=x0, f=f, g=pa.IndConsensus(5), gamma=gamma)
    u = f.image(y)
    assert k < 3000 and f.objective(u) < f.objective(b)
    with pytest.raises(ValueError):
        next(iter(pa.DouglasRachfordIteration(x0, f=f, g=pa.IndConsensus(4), gamma=gamma)))


def _afba_cases(golden, T):
    A, b, lam, xstar = _lasso_4x5(golden, T)
    beta_f = T(np.linalg.norm(A, 2) ** 2)
    xs_en = golden("unit_elasticnet")["xstar"].astype(T)
    z5, z4 = np.zeros(5, T), np.zeros(4, T)
    cases = [
        ("lasso_g", dict(x0=z5, y0=z5, f=o.LeastSquares(A, b), g=o.NormL1(lam), beta_f=beta_f, theta=1, mu=1),
         dict(x0=z5, y0=z5, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), beta_f=beta_f, theta=1, mu=1), xstar, 80),
        ("lasso_h", dict(x0=z5, y0=z5, f=o.LeastSquares(A, b), h=o.NormL1(lam), beta_f=beta_f, theta=1, mu=1),
         dict(x0=z5, y0=z5, f=pa.LeastSquares(A, b), h=pa.NormL1(lam), beta_f=beta_f, theta=1, mu=1), xstar, 100),
        ("lasso_L", dict(x0=z5, y0=z4, h=po.SqrNormL2Translated(b, 1.0), L=A, g=o.NormL1(lam), theta=1, mu=1),
         dict(x0=z5, y0=z4, h=pa.SqrNormL2(1.0, b), L=A, g=pa.NormL1(lam), theta=1, mu=1), xstar, 150),
    ]
    for theta, mu, maxit in [(2, 0, 130), (1, 1, 2000), (0, 1, 320), (0, 0, 194), (1, 0, 130)]:
        cases.append((f"enet_{theta}_{mu}",
                      dict(x0=z5, y0=z4, f=ao.SqrNormL2Smooth(1.0), g=o.NormL1(T(1)), h=po.SqrNormL2Translated(b, 1.0), L=A, beta_f=1, theta=theta, mu=mu),
                      dict(x0=z5, y0=z4, f=pa.SqrNormL2(1.0), g=pa.NormL1(1.0), h=pa.SqrNormL2(1.0, b), L=A, beta_f=1, theta=theta, mu=mu), xs_en, maxit))
    return cases


@pytest.mark.parametrize("T", TYPES)
def test_afba_host_logic_matches_oracle(emu, golden, T):
    # test/problems/test_lasso_small.jl:233-275 and test_elasticnet.jl:56-113 on the product's host logic
    for name, kw_o, kw_p, xstar, bound in _afba_cases(golden, T):
        it_o, it_p = ao.AFBAIteration(**kw_o), pa.AFBAIteration(**kw_p)
        assert it_p.gamma == it_o.gamma, name
        for k, (so, sp) in enumerate(zip(it_o, it_p)):
            tol = 1e-12 if T is np.float64 else 1e-5
            assert np.max(np.abs(sp.xbar.numpy() - so.xbar)) <= tol and np.max(np.abs(sp.ybar.numpy() - so.ybar)) <= tol, (name, k)
            assert np.max(np.abs(sp.x.numpy() - so.x)) <= tol and np.max(np.abs(sp.y.numpy() - so.y)) <= tol, (name, k)
            assert np.max(np.abs(sp.FPR_x.numpy() - so.FPR_x)) <= tol
            if k == 10:
                break
        (x, y), it = pa.AFBA(tol=T(1e-6), **{k_: v for k_, v in kw_p.items() if k_ in ("theta", "mu")})(**{k_: v for k_, v in kw_p.items() if k_ not in ("theta", "mu")})
        (xo, yo), ito = ao.afba(tol=T(1e-6), **kw_o)
        assert x.dtype == T and y.dtype == T and np.max(np.abs(x - xstar)) <= 1e-4 and it <= bound, (name, it)
        assert abs(it - ito) <= max(2, ito // 20), (name, it, ito)
    A, b, lam, xstar = _lasso_4x5(golden, T)
    (x, y), it = pa.ChambollePock(tol=T(1e-6))(x0=np.zeros(5, T), y0=np.zeros(4, T), h=pa.SqrNormL2(1.0, b), L=A, g=pa.NormL1(lam))
    assert np.max(np.abs(x - xstar)) <= 1e-4
    with pytest.raises(ValueError):
        pa.AFBAIteration(np.zeros(5, T), np.zeros(5, T), f=pa.LeastSquares(A, b))
    with pytest.raises(ValueError):
        pa.AFBAIteration(np.zeros(5, T), np.zeros(5, T), lambda_=0.5)


@pytest.mark.parametrize("T", TYPES)
def test_two_phase_and_user_prox_keep_the_smooth_value(emu, golden, T):
    """IndBallL2 (two-phase prox) and user `prox_` callbacks run `pb_forward` inside the step, which reuses the scalar slot that
    carries a built-in f's value: the engine must fetch f first (`_Engine.pre_resolve`).  Adaptive stepsizes and PANOC's
    line search consume that value, so a wrong f shows up as a different iteration count."""
    d = golden("lasso_small")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    n = A.shape[1]
    tol = T(1e-6 if T is np.float64 else 1e-4)
    x, it = pa.PANOC(tol=tol)(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=pa.IndBallL2(0.5))
    xo, ito = po.panoc(np.zeros(n, T), f=o.LeastSquares(A, b), g=o.IndBallL2(T(0.5)), tol=tol)
    assert abs(it - ito) <= max(3, ito // 10) and np.max(np.abs(x - xo)) <= (1e-9 if T is np.float64 else 1e-3)
    for mk, mko in ((pa.FastForwardBackward, o.fast_forward_backward), (pa.ForwardBackward, o.forward_backward)):
        z, it = mk(tol=tol, driver="python")(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=pa.IndBallL2(0.5))
        zo, ito = mko(np.zeros(n, T), o.LeastSquares(A, b), o.IndBallL2(T(0.5)), tol=tol)
        assert abs(it - ito) <= max(2, ito // 50) and np.max(np.abs(z - zo)) <= (1e-9 if T is np.float64 else 1e-3), (mk.__name__, it, ito)

    class UserL1:
        def prox_(self, z, y, gam):
            zz, v = o.NormL1(T(1)).prox(y.numpy(), gam)
            z.copy_(torch.as_tensor(zz))
            return v

    z, it = pa.FastForwardBackward(tol=tol)(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=UserL1())
    zo, ito = o.fast_forward_backward(np.zeros(n, T), o.LeastSquares(A, b), o.NormL1(T(1)), tol=tol)
    assert abs(it - ito) <= max(2, ito // 50) and np.max(np.abs(z - zo)) <= (1e-9 if T is np.float64 else 1e-3)


def test_empty_iterates(emu):
    assert pa.DouglasRachford(maxit=3)(x0=np.zeros(0), f=pa.NormL1(1.0), g=pa.IndBox(-1, 1), gamma=1.0)[1] == 1
    assert pa.PANOC(maxit=3)(x0=np.zeros(0), g=pa.NormL1(1.0), gamma=1.0)[1] == 1


def test_chambolle_pock_tv_with_the_stencil_operator(emu):
    """Row f4 as SURVEY.md describes it: TV denoising by Chambolle-Pock with L = forward differences, h = lam*||.||_1.  Host logic
    against the AFBA oracle run on the dense matrix of the same operator; the minimiser agrees with the Douglas-Rachford TV
    splitting (oracle/tv_oracle.py)."""
    from oracle.stencil_oracle import FiniteDifference2D as FDo

    T = np.float64
    rng = np.random.default_rng(4)
    H, W = 6, 7
    img = np.zeros((H, W))
    img[2:5, 1:4] = 1.0
    b = (img + 0.1 * rng.standard_normal((H, W))).reshape(-1)
    lam = 0.2
    Ld = np.stack([FDo(H, W).mul(e) for e in np.eye(H * W)], axis=1)                  # dense 2HW x HW matrix of L
    v = rng.standard_normal(2 * H * W)
    assert np.allclose(Ld.T @ v, FDo(H, W).mul_t(v), atol=1e-14)                     # mul_t is the exact transpose
    gam = (0.3, 0.3)                                                                   # gamma1*gamma2*||L||^2 = 0.72 < 1
    it_o = ao.AFBAIteration(np.zeros(H * W), np.zeros(2 * H * W), g=po.SqrNormL2Translated(b, 1.0), h=o.NormL1(lam), L=Ld, theta=2, gamma=gam)
    it_p = pa.AFBAIteration(np.zeros(H * W), np.zeros(2 * H * W), g=pa.SqrNormL2(1.0, b), h=pa.NormL1(lam), L=pa.FiniteDifference2D(H, W), theta=2, gamma=gam)
    for k, (so, sp) in enumerate(zip(it_o, it_p)):
        assert np.max(np.abs(sp.xbar.numpy() - so.xbar)) <= 1e-12 and np.max(np.abs(sp.ybar.numpy() - so.ybar)) <= 1e-12, k
        if k == 15:
            break
    (x, y), it = pa.ChambollePock(tol=1e-9, maxit=20000)(x0=np.zeros(H * W), y0=np.zeros(2 * H * W), g=pa.SqrNormL2(1.0, b), h=pa.NormL1(lam),
                                                        L=pa.FiniteDifference2D(H, W))
    f = tvo.TVSplit(b.reshape(H, W), lam, (H, W))
    dr = iter(po.DouglasRachfordIteration(np.tile(b, 5), f=f, g=tvo.Consensus(5), gamma=1.0))
    for _ in range(6000):
        st = next(dr)
    u = st.z[: H * W]
    assert it < 20000 and np.max(np.abs(x - u)) <= 1e-5
    assert abs(f.objective(x) - f.objective(u)) <= 1e-8 * f.objective(u)


def test_afba_linear_program_host_logic(emu, golden):
    """test/problems/test_linear_programs.jl:102-125 through the product's host logic (IndNonnegative = IndBox(0, inf), IndPoint(b) =
    IndBox(b, b) with per-element bounds, f = LinearFunction(c)): the four optimality measures of the reference's assert_lp_solution."""
    d = golden("unit_linear_program")
    A, b, c = d["A"], d["b"], d["c"]
    tol = 100 * np.finfo(np.float64).eps
    bt = torch.as_tensor(b.copy())
    (x, y), it = pa.AFBA(tol=tol, maxit=100_000)(x0=np.zeros(10), y0=np.zeros(8), f=pa.LinearFunction(torch.as_tensor(c.copy())),
                                                  g=pa.IndBox(0.0, float("inf")), h=pa.IndBox(bt, bt), L=A, beta_f=0)
    (xo, yo), ito = ao.afba(np.zeros(10), np.zeros(8), f=ao.LinearSmooth(c), g=o.IndBox(0.0, np.inf), h=o.IndBox(b, b), L=A, beta_f=0, tol=tol, maxit=100_000)
    assert abs(it - ito) <= max(5, ito // 20), (it, ito)
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    quality = (-min(0.0, x64.min()), np.linalg.norm(A @ x64 - b), max(0.0, (-A.T @ y64 - c).max()), abs((c + A.T @ y64) @ x64))
    assert all(q <= 1000 * tol for q in quality), quality


@pytest.mark.parametrize("T", TYPES)
def test_panoc_reference_problems_with_user_callbacks(emu, golden, T):
    """The reference's remaining PANOC tests through the product's host logic, with the smooth term supplied the way the reference
    supplies it (a user callback, i.e. NOT a built-in quadratic): test_sparse_logistic_small.jl:101-109 (adaptive, A a matrix),
    test_lasso_small_strongly_convex.jl:155-162, test_nonconvex_qp.jl:68-103 (random 100-dim box-constrained QPs)."""
    A, b, _, _ = _lasso_4x5(golden, T)

    class Logistic:                                     # logistic_loss(u - b), labels all one
        def value_and_gradient(self, u):
            v, g = po.LogisticLoss(b).value_and_gradient(u.numpy())
            return v, torch.as_tensor(g)

    d = golden("unit_sparse_logistic")
    x, it = pa.PANOC(adaptive=True, tol=T(1e-6))(x0=np.zeros(5, T), f=Logistic(), A=A, g=pa.NormL1(float(d["lam"])))
    assert x.dtype == T and np.max(np.abs(x - d["xstar"].astype(T))) <= 1e-4 and it < 50

    sc = golden("unit_lasso_sc_5x5")
    A2, b2, x02 = np.asfortranarray(sc["A"].astype(T)), sc["b"].astype(T), sc["x0"].astype(T)

    class LsqCallback:                                  # fA_autodiff of the reference test
        def value_and_gradient(self, x):
            v, g = o.LeastSquares(A2, b2).value_and_gradient(x.numpy())
            return v, torch.as_tensor(g)

    y, it = pa.PANOC(tol=T(1e-4))(x0=x02, f=LsqCallback(), g=pa.NormL1(T(sc["lam"])), Lf=T(sc["Lf"]))
    assert np.max(np.abs(y - sc["xstar"].astype(T))) <= 1e-4 and it < 45 and np.array_equal(x02, sc["x0"].astype(T))

    if T is np.float64:
        for k in range(1, 4):
            rng = np.random.default_rng(k)
            n = 100
            U, _ = np.linalg.qr(rng.standard_normal((n, n)))
            ev = 2 * rng.random(n) - 1
            Q = U @ np.diag(ev) @ U.T
            Q = 0.5 * (Q + Q.T)
            qv = rng.standard_normal(n)

            class QP:
                def value_and_gradient(self, x):
                    v, g = po.QuadraticForm(Q, qv).value_and_gradient(x.numpy())
                    return v, torch.as_tensor(g)

            gamma = 0.95 / np.max(np.abs(ev))
            x, it = pa.PANOC(tol=1e-4)(x0=np.zeros(n), f=QP(), g=pa.IndBox(-1.0, 1.0))
            z = np.minimum(1.0, np.maximum(-1.0, x - gamma * (Q @ x + qv)))
            assert np.max(np.abs(x - z)) / gamma <= 1e-4 and it < 1000


def test_finite_extrapolation_sequence_is_only_missed_when_really_needed(emu, golden):
    """The fixed-stepsize path draws beta one iteration ahead of the reference (it fuses the next extrapolation into the current pass).
    A finite user sequence must not fail before the iteration that really needs the missing coefficient."""
    import itertools

    T = np.float64
    A, b, lam, _ = _lasso_4x5(golden, T)
    Lf = T(np.linalg.norm(A, 2) ** 2)
    kw = dict(x0=np.zeros(5, T), f=pa.LeastSquares(A, b), g=pa.NormL1(lam), Lf=Lf)
    # the reference draws one coefficient per step: iterations 2, 3, 4 of a maxit = 4 run need 3 of them
    z, k = pa.FastForwardBackward(tol=-1.0, maxit=4, driver="python")(extrapolation_sequence=iter([0.1, 0.2, 0.3]), **kw)
    zr, kr = pa.FastForwardBackward(tol=-1.0, maxit=4, driver="python")(extrapolation_sequence=itertools.chain([0.1, 0.2, 0.3], itertools.repeat(0.9)), **kw)
    assert k == kr == 4 and np.array_equal(z, zr)
    with pytest.raises(RuntimeError, match="exhausted"):
        pa.FastForwardBackward(tol=-1.0, maxit=6, driver="python")(extrapolation_sequence=iter([0.1, 0.2, 0.3]), **kw)
    # a bounded itertools.repeat is not mistaken for the constant sequence of the native driver
    from proxb200.algorithms import _native_sequence

    class _It:
        extrapolation_sequence = itertools.repeat(0.5, 3)

    assert _native_sequence(_It, T) is None
    _It.extrapolation_sequence = itertools.repeat(0.5)
    assert _native_sequence(_It, T) is not None
