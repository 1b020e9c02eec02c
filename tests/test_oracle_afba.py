"""Pins the AFBA / Vu-Condat / Chambolle-Pock restatement (oracle/afba_oracle.py) against the reference's known answers:
test/problems/test_lasso_small.jl:233-275 and test/problems/test_elasticnet.jl:56-113.  CPU only."""
import numpy as np
import pytest

from oracle import afba_oracle as ao
from oracle import fb_oracle as o
from oracle import panoc_oracle as po

TYPES = [np.float64, np.float32]


def _data(golden, T):
    d = golden("unit_lasso_4x5")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    return A, b, T(T(0.1) * np.max(np.abs(A.T @ b))), d["xstar"].astype(T)


@pytest.mark.parametrize("T", TYPES)
def test_afba_lasso_small_three_formulations(golden, T):
    A, b, lam, xstar = _data(golden, T)
    n, m = 5, 4
    fA = o.LeastSquares(A, b)
    beta_f = T(np.linalg.norm(A, 2) ** 2)
    x0 = np.zeros(n, T)
    (x, y), it = ao.afba(x0, np.zeros(n, T), f=fA, g=o.NormL1(lam), beta_f=beta_f, theta=1, mu=1, tol=T(1e-6))
    assert x.dtype == T and y.dtype == T and np.max(np.abs(x - xstar)) <= 1e-4 and it <= 80 and not x0.any()
    (x, y), it = ao.afba(x0, np.zeros(n, T), f=fA, h=o.NormL1(lam), beta_f=beta_f, theta=1, mu=1, tol=T(1e-6))
    assert np.max(np.abs(x - xstar)) <= 1e-4 and it <= 100
    (x, y), it = ao.afba(x0, np.zeros(m, T), h=po.SqrNormL2Translated(b, 1.0), L=A, g=o.NormL1(lam), theta=1, mu=1, tol=T(1e-6))
    assert np.max(np.abs(x - xstar)) <= 1e-4 and it <= 150


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("theta,mu,maxit", [(2, 0, 130), (1, 1, 2000), (0, 1, 320), (0, 0, 194), (1, 0, 130)])
def test_afba_elastic_net(golden, T, theta, mu, maxit):
    A, b, _, _ = _data(golden, T)
    xstar = golden("unit_elasticnet")["xstar"].astype(T)
    (x, y), it = ao.afba(np.zeros(5, T), np.zeros(4, T), f=ao.SqrNormL2Smooth(1.0), g=o.NormL1(T(1)), h=po.SqrNormL2Translated(b, 1.0),
                         L=A, beta_f=1, theta=theta, mu=mu, tol=T(1e-6))
    assert np.max(np.abs(x - xstar)) <= 1e-4 and it <= maxit
    rng = np.random.default_rng(0)
    (x, y), it = ao.afba(rng.standard_normal(5).astype(T), rng.standard_normal(4).astype(T), f=ao.SqrNormL2Smooth(1.0), g=o.NormL1(T(1)),
                         h=po.SqrNormL2Translated(b, 1.0), L=A, beta_f=1, theta=theta, mu=mu, tol=T(1e-6))
    assert np.max(np.abs(x - xstar)) <= 1e-4


def test_afba_argument_errors_and_aliases(golden):
    A, b, lam, xstar = _data(golden, np.float64)
    with pytest.raises(ValueError):
        ao.AFBAIteration(np.zeros(5), np.zeros(5), f=o.LeastSquares(A, b))            # beta_f missing
    with pytest.raises(ValueError):
        ao.AFBAIteration(np.zeros(5), np.zeros(5), lambda_=0.5)                       # stepsizes needed
    with pytest.raises(ValueError):
        ao.AFBAIteration(np.zeros(5), np.zeros(4), h=o.NormL1(1.0), L=A, theta=0.5, mu=0.3)
    (x, y), it = ao.chambolle_pock(np.zeros(5), np.zeros(4), h=po.SqrNormL2Translated(b, 1.0), L=A, g=o.NormL1(lam), tol=1e-7)
    assert np.max(np.abs(x - xstar)) <= 1e-4
    (x, y), it = ao.vu_condat(np.zeros(5), np.zeros(5), f=o.LeastSquares(A, b), g=o.NormL1(lam), beta_f=np.linalg.norm(A, 2) ** 2, tol=1e-7)
    assert np.max(np.abs(x - xstar)) <= 1e-4


def _lp_quality(d, x, y):
    A, b, c = d["A"], d["b"], d["c"]
    x, y = x.astype(np.float64), y.astype(np.float64)
    return (-min(0.0, x.min()), np.linalg.norm(A @ x - b), max(0.0, (-A.T @ y - c).max()), abs((c + A.T @ y) @ x))


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("solver", ["afba", "vu_condat"])
def test_linear_program_like_the_reference(golden, T, solver):
    # test/problems/test_linear_programs.jl:102-151: f = <c, .>, g = IndNonnegative, h = IndPoint(b), L = A, beta_f = 0
    d = golden("unit_linear_program")
    A, b, c = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T), d["c"].astype(T)
    tol, maxit = 100 * np.finfo(T).eps, 100_000
    x0, y0 = np.zeros(10, T), np.zeros(8, T)
    (x, y), it = getattr(ao, solver)(x0, y0, f=ao.LinearSmooth(c), g=o.IndBox(T(0), T(np.inf)), h=o.IndBox(b, b), L=A, beta_f=0, tol=tol, maxit=maxit)
    assert x.dtype == T and y.dtype == T and it <= maxit and not x0.any() and not y0.any()
    assert all(q <= 1000 * tol for q in _lp_quality(d, x, y)), _lp_quality(d, x, y)
