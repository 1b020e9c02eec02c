"""GPU parity of row f4 (AFBA / Vu-Condat / Chambolle-Pock, src/algorithms/primal_dual.jl) against oracle/afba_oracle.py, on the
reference's own test problems (test/problems/test_lasso_small.jl:233-275, test_elasticnet.jl:56-113).  Bars: `pb_conj_prox`
bit-exact; iterates within 1e-11 (fp64) / 1e-5 (fp32) of the oracle for the first iterations (the products with L round
differently from BLAS); every x_star / iteration bound the reference asserts; iteration count within 5 % of the oracle."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import afba_oracle as ao  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402
from oracle import panoc_oracle as po  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import ptr  # noqa: E402

from conftest import load_golden  # noqa: E402
from gpu_util import ctx, dev, dt, prox_desc  # noqa: E402

TYPES = [np.float64, np.float32]


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("n", [0, 1, 5, 1000, 100_003])
def test_conj_prox_bit_exact(T, n):
    rng = np.random.default_rng(n + 2)
    v, b = rng.standard_normal(n).astype(T), rng.standard_normal(n).astype(T)
    c = ctx()
    bd = dev(b)
    for gamma in (T(0.3), T(2.5)):
        for desc, term in ((prox_desc(L.PB_PROX_L1, 0.7), o.NormL1(T(0.7))), (prox_desc(L.PB_PROX_SQRL2, 1.5, v0=bd), po.SqrNormL2Translated(b, 1.5)),
                           (prox_desc(L.PB_PROX_BOX, -0.2, 0.4), o.IndBox(-0.2, 0.4)), (prox_desc(L.PB_PROX_ZERO), o.ZeroFn())):
            out = torch.empty(n, dtype=dev(v).dtype, device="cuda")
            L.check(c.lib.pb_conj_prox(c.h, dt(T), n, ptr(dev(v)), float(gamma), C.byref(desc), ptr(out)))
            assert np.array_equal(out.cpu().numpy(), ao.conj_prox(term, v, gamma))


def _cases(T):
    d = load_golden("unit_lasso_4x5")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    lam, xstar = T(T(0.1) * np.max(np.abs(A.T @ b))), d["xstar"].astype(T)
    beta_f = T(np.linalg.norm(A, 2) ** 2)
    xs_en = load_golden("unit_elasticnet")["xstar"].astype(T)
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
def test_afba_like_the_reference(T):
    for name, kw_o, kw_p, xstar, bound in _cases(T):
        it_o, it_p = ao.AFBAIteration(**kw_o), pa.AFBAIteration(**kw_p)
        assert it_p.gamma == it_o.gamma, name
        tol = 1e-11 if T is np.float64 else 1e-5
        for k, (so, sp) in enumerate(zip(it_o, it_p)):
            assert np.max(np.abs(sp.xbar.cpu().numpy() - so.xbar)) <= tol and np.max(np.abs(sp.ybar.cpu().numpy() - so.ybar)) <= tol, (name, k)
            assert np.max(np.abs(sp.x.cpu().numpy() - so.x)) <= tol and np.max(np.abs(sp.y.cpu().numpy() - so.y)) <= tol, (name, k)
            if k == 10:
                break
        alg_kw = {k_: v for k_, v in kw_p.items() if k_ in ("theta", "mu")}
        call_kw = {k_: v for k_, v in kw_p.items() if k_ not in ("theta", "mu")}
        x0 = call_kw["x0"].copy()
        (x, y), it = pa.AFBA(tol=T(1e-6), **alg_kw)(**call_kw)
        (xo, yo), ito = ao.afba(tol=T(1e-6), **kw_o)
        assert isinstance(x, np.ndarray) and x.dtype == T and y.dtype == T
        assert np.max(np.abs(x - xstar)) <= 1e-4 and it <= bound, (name, it)
        assert abs(it - ito) <= max(2, ito // 20), (name, it, ito)
        assert np.array_equal(call_kw["x0"], x0)


@pytest.mark.parametrize("T", TYPES)
def test_chambolle_pock_and_vu_condat(T):
    d = load_golden("unit_lasso_4x5")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    lam, xstar = T(T(0.1) * np.max(np.abs(A.T @ b))), d["xstar"].astype(T)
    (x, y), it = pa.ChambollePock(tol=T(1e-6))(x0=np.zeros(5, T), y0=np.zeros(4, T), h=pa.SqrNormL2(1.0, b), L=A, g=pa.NormL1(lam))
    assert np.max(np.abs(x - xstar)) <= 1e-4
    (x, y), it = pa.VuCondat(tol=T(1e-6))(x0=np.zeros(5, T), y0=np.zeros(5, T), f=pa.LeastSquares(A, b), g=pa.NormL1(lam),
                                         beta_f=T(np.linalg.norm(A, 2) ** 2))
    assert np.max(np.abs(x - xstar)) <= 1e-4
    with pytest.raises(ValueError):
        pa.AFBAIteration(np.zeros(5, T), np.zeros(5, T), f=pa.LeastSquares(A, b))
