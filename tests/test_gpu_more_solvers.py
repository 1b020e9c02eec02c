"""GPU tests of the two-phase / user-prox value hand-over (`_Engine.pre_resolve`), AFBA / VuCondat on the reference's linear program and the
native PANOC driver.  Written at the end of round 1 without a hardware run (then opt-in); they ran green on B200 at the start of round 2
and are part of the normal suite since.  CPU twins: tests/test_host_emulated.py."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import afba_oracle as ao  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402

from conftest import load_golden  # noqa: E402

TYPES = [np.float64, np.float32]


@pytest.mark.parametrize("T", TYPES)
def test_two_phase_prox_with_adaptive_stepsize_and_panoc(T):
    """IndBallL2's two-phase step runs `pb_forward`, which reuses the AUX slot carrying a built-in f's value; the engine fetches f
    first (`_Engine.pre_resolve`).  The adaptive line searches and PANOC consume that value (host-logic twin of this test:
    tests/test_host_emulated.py::test_two_phase_and_user_prox_keep_the_smooth_value)."""
    from oracle import panoc_oracle as po

    d = load_golden("lasso_small")
    A, b = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T)
    n = A.shape[1]
    tol = T(1e-6 if T == np.float64 else 1e-4)
    close = 1e-6 if T == np.float64 else 2e-3
    for mk, mko in ((pa.FastForwardBackward, o.fast_forward_backward), (pa.ForwardBackward, o.forward_backward)):
        z, it = mk(tol=tol)(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=pa.IndBallL2(T(0.5)))
        zo, ito = mko(np.zeros(n, T), o.LeastSquares(A, b), o.IndBallL2(T(0.5)), tol=tol)
        assert abs(it - ito) <= max(5, ito // 20), (mk.__name__, it, ito)
        assert np.max(np.abs(z - zo)) <= close
    x, it = pa.PANOC(tol=tol)(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=pa.IndBallL2(T(0.5)))
    xo, ito = po.panoc(np.zeros(n, T), f=o.LeastSquares(A, b), g=o.IndBallL2(T(0.5)), tol=tol)
    assert it <= 2 * ito + 5 and np.max(np.abs(x - xo)) <= 10 * close, (it, ito)


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("solver", ["AFBA", "VuCondat"])
def test_linear_program_like_the_reference(T, solver):
    """test/problems/test_linear_programs.jl:102-151: f = <c, .>, g = IndNonnegative (= IndBox(0, inf)), h = IndPoint(b) (= IndBox(b, b)
    with per-element bounds), L = A, beta_f = 0; the four optimality measures of assert_lp_solution to 1000 * 100 eps."""
    d = load_golden("unit_linear_program")
    A, b, c = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T), d["c"].astype(T)
    tol, maxit = 100 * np.finfo(T).eps, 100_000
    bt = torch.as_tensor(b).cuda()
    x0, y0 = np.zeros(10, T), np.zeros(8, T)
    (x, y), it = getattr(pa, solver)(tol=tol, maxit=maxit)(x0=x0, y0=y0, f=pa.LinearFunction(torch.as_tensor(c).cuda()),
                                                           g=pa.IndBox(0.0, float("inf")), h=pa.IndBox(bt, bt), L=A, beta_f=0)
    (xo, yo), ito = getattr(ao, "afba" if solver == "AFBA" else "vu_condat")(x0, y0, f=ao.LinearSmooth(c), g=o.IndBox(T(0), T(np.inf)),
                                                                            h=o.IndBox(b, b), L=A, beta_f=0, tol=tol, maxit=maxit)
    assert x.dtype == T and y.dtype == T and it <= maxit and not x0.any() and not y0.any()
    x64, y64 = x.astype(np.float64), y.astype(np.float64)
    A64, b64, c64 = d["A"], d["b"], d["c"]
    quality = (-min(0.0, x64.min()), np.linalg.norm(A64 @ x64 - b64), max(0.0, (-A64.T @ y64 - c64).max()), abs((c64 + A64.T @ y64) @ x64))
    assert all(q <= 1000 * tol for q in quality), quality
    assert it <= 3 * ito + 100, (it, ito)          # same convergence regime as the oracle (the reference only asserts it <= maxit)


@pytest.mark.parametrize("T", TYPES)
def test_native_panoc_driver_equals_python_host(T):
    """pb_panoc_solve (csrc/panoc_solve.cu) is a native twin of panoc.py: same kernels, same scalar arithmetic in R -> identical
    iteration counts, backtrack counters and bit-identical solutions."""
    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T), T(d["lam"])
    n = A.shape[1]
    Lf = T(np.linalg.norm(d["A"], 2) ** 2)
    tol = T(1e-6 if T == np.float64 else 1e-4)
    cases = [
        dict(f=lambda: pa.LeastSquares(A, b), A=None, kw={}),                                   # adaptive, A = I, quadratic branch
        dict(f=lambda: pa.LeastSquares(A, b), A=None, kw=dict(Lf=Lf)),                          # fixed stepsize
        dict(f=lambda: pa.SquaredDistance(b), A=A, kw={}),                                      # matrix A, general branch
        dict(f=lambda: pa.SquaredDistance(b), A=A, kw=dict(Lf=Lf, directions=pa.NoAcceleration())),
        dict(f=lambda: pa.LeastSquares(A, b), A=None, kw=dict(directions=pa.LBFGS(2), max_backtracks=3)),
    ]
    for case in cases:
        for g in (pa.NormL1(lam), pa.IndBox(T(-0.05), T(0.05)), pa.NormL21(lam, 4)):
            kw = dict(case["kw"])
            if case["A"] is not None:
                kw["A"] = case["A"]
            sp, sn = pa.PANOC(tol=tol, maxit=400), pa.PANOC(tol=tol, maxit=400, driver="native")
            zp, itp = sp(x0=np.zeros(n, T), f=case["f"](), g=g, **kw)
            zn, itn = sn(x0=np.zeros(n, T), f=case["f"](), g=g, **kw)
            assert sn.last_driver == "native" and itn == itp, (itn, itp, type(g).__name__, list(kw))
            assert np.array_equal(zn, zp)
            assert sn.last_iteration.tau_backtracks == sp.last_iteration.tau_backtracks
            assert sn.last_iteration.backtracks == sp.last_iteration.backtracks
