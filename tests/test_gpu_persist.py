"""The persistent on-device solver (csrc/persist.cu: the whole ForwardBackward / FastForwardBackward solve in ONE cooperative
kernel launch, chosen automatically by pb_solve for cache-resident dense least squares) against (i) the one-kernel-per-operation
path -- bit-identical iterates, scalars and iteration counts for every CTA count -- and (ii) the oracle's iteration counts on the
reference's benchmark fixtures (benchmark/benchmarks.jl:47-61: 480 / 788 / 3912 FFB, 10000 / 1251 / 2811 FB)."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context  # noqa: E402

from conftest import load_golden  # noqa: E402

ORACLE_COUNTS = {("tiny", "ffb"): 480, ("small", "ffb"): 788, ("medium", "ffb"): 3912,
                 ("tiny", "fb"): 10000, ("small", "fb"): 1251, ("medium", "fb"): 2811}


def _mode(v):
    c = Context.get()
    L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_PERSISTENT, v))


@pytest.fixture(autouse=True)
def _restore_mode():
    yield
    _mode(0)


def _solve(mk, mode, **kw):
    _mode(mode)
    s = mk()
    z, it = s(**kw)
    st = s.last_state
    return z, it, s.last_persistent_ctas, (st.gamma, st.f_x, st.g_z, float(st.res_norm_inf)), (st.x.clone(), st.grad_f_x.clone(), st.z.clone()), s


def _same(a, b):
    (za, ia, _, sa, va, _), (zb, ib, _, sb, vb, _) = a, b
    assert ia == ib, (ia, ib)
    assert np.array_equal(za, zb, equal_nan=True)
    assert np.array_equal(np.array(sa, dtype=np.float64), np.array(sb, dtype=np.float64), equal_nan=True), (sa, sb)
    for u, v in zip(va, vb):
        assert torch.equal(torch.nan_to_num(u, nan=1e300 if u.dtype == torch.float64 else 1e30), torch.nan_to_num(v, nan=1e300 if v.dtype == torch.float64 else 1e30))


@pytest.mark.parametrize("name", ["tiny", "small", "medium"])
@pytest.mark.parametrize("alg", ["ffb", "fb"])
def test_fixtures_one_launch_same_bits_same_counts(name, alg):
    d = load_golden(f"lasso_{name}")
    A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
    n = A.shape[1]
    mk = lambda: (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6, driver="native")   # noqa: E731
    kw = dict(x0=np.zeros(n), f=pa.LeastSquares(A, b), g=pa.NormL1(lam))
    ctx = Context.get()
    ref = _solve(mk, -1, **kw)                       # one kernel per operation
    assert ref[2] == 0
    l0 = ctx.launches()
    auto = _solve(mk, 0, **kw)
    assert auto[2] >= 1 and ctx.launches() - l0 == 1, "the whole solve must be one kernel launch"
    _same(ref, auto)
    assert auto[1] == ORACLE_COUNTS[(name, alg)]
    for G in (1, 2, 3, 8, 32):
        got = _solve(mk, G, **kw)
        assert 1 <= got[2] <= G
        _same(ref, got)
    # converged to the fixture's own solution (objective gap, north_star: <= 1e-6 relative)
    obj = lambda v: 0.5 * np.sum((d["A"] @ v - b) ** 2) + lam * np.sum(np.abs(v))   # noqa: E731
    if auto[1] < 10000:
        assert abs(obj(auto[0]) - obj(d["xstar"])) <= 1e-6 * obj(d["xstar"])


@pytest.mark.parametrize("T", [np.float64, np.float32])
def test_every_variant_matches_the_multi_kernel_path(T):
    rng = np.random.default_rng(5)
    shapes = [(4, 5), (5, 10), (50, 100), (64, 96), (33, 257), (200, 500), (130, 70), (256, 40)]
    for (m, n) in shapes:
        A = np.asfortranarray(rng.standard_normal((m, n)).astype(T) / np.sqrt(m))
        xt = np.zeros(n)
        xt[rng.choice(n, max(1, n // 10), replace=False)] = rng.standard_normal(max(1, n // 10))
        b = (A.astype(np.float64) @ xt + 0.01 * rng.standard_normal(m)).astype(T)
        lam = T(0.1 * np.max(np.abs(A.T.astype(np.float64) @ b)))
        Lf = T(np.linalg.norm(A.astype(np.float64), 2) ** 2 * 1.01)
        tol = T(1e-6 if T == np.float64 else 1e-4)
        lo_v = torch.as_tensor((-0.05 - 0.01 * rng.random(n)).astype(T)).cuda()
        hi_v = torch.as_tensor((0.05 + 0.01 * rng.random(n)).astype(T)).cuda()
        variants = [
            (pa.FastForwardBackward, {}),
            (pa.FastForwardBackward, dict(increase_gamma=T(1.01))),
            (pa.FastForwardBackward, dict(Lf=Lf)),
            (pa.FastForwardBackward, dict(Lf=Lf, mf=T(0.01))),
            (pa.FastForwardBackward, dict(Lf=Lf, extrapolation_sequence=pa.FixedNesterovSequence(T))),
            (pa.FastForwardBackward, dict(Lf=Lf, extrapolation_sequence=pa.SimpleNesterovSequence(T))),
            (pa.FastForwardBackward, dict(Lf=Lf, extrapolation_sequence=pa.ConstantNesterovSequence(T(0.05), T(1) / Lf))),
            (pa.ForwardBackward, {}),
            (pa.ForwardBackward, dict(increase_gamma=T(1.01))),
            (pa.ForwardBackward, dict(Lf=Lf)),
        ]
        gs = [pa.NormL1(lam), pa.IndBox(T(-0.05), T(0.05)), pa.IndBox(lo_v, hi_v), pa.Zero()]
        for vi, (mkc, kw) in enumerate(variants):
            g = gs[vi % len(gs)] if (m, n) != (50, 100) else None
            for gg in ([g] if g is not None else gs):
                mk = lambda: mkc(tol=tol, maxit=400, driver="native")   # noqa: E731
                args = dict(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=gg, **kw)
                ref = _solve(mk, -1, **args)
                auto = _solve(mk, 0, **args)
                assert ref[2] == 0 and auto[2] >= 1, (m, n)
                _same(ref, auto)
                if (m, n) in ((200, 500), (33, 257)):
                    _same(ref, _solve(mk, 5, **args))


def test_maxit_edges_and_python_host_agreement():
    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
    n = A.shape[1]
    for maxit in (1, 2, 3, 17):
        for mkc in (pa.FastForwardBackward, pa.ForwardBackward):
            for kw in ({}, dict(Lf=np.linalg.norm(A, 2) ** 2)):
                args = dict(x0=np.zeros(n), f=pa.LeastSquares(A, b), g=pa.NormL1(lam), **kw)
                _mode(0)
                sn, sp = mkc(tol=-1.0, maxit=maxit, driver="native"), mkc(tol=-1.0, maxit=maxit, driver="python")
                zn, kn = sn(**args)
                zp, kp = sp(**args)
                assert kn == kp == maxit and sn.last_persistent_ctas >= 1
                assert np.array_equal(zn, zp)
                assert (sn.last_state.gamma, sn.last_state.f_x, sn.last_state.g_z) == (sp.last_state.gamma, sp.last_state.f_x, sp.last_state.g_z)


def test_faster_than_one_kernel_per_operation():
    """The point of the persistent kernel: no launch / synchronisation latency per iteration.  Conservative bar (2x); the measured
    numbers live in profiles/ (tools/perf_fixtures.py)."""
    import time

    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"]), d["b"], float(d["lam"])
    f = pa.LeastSquares(A, b)
    args = dict(x0=np.zeros(A.shape[1]), f=f, g=pa.NormL1(lam))
    t = {}
    for mode in (-1, 0):
        _mode(mode)
        s = pa.FastForwardBackward(tol=1e-6, driver="native")
        s(**args)
        t0 = time.perf_counter()
        for _ in range(3):
            s(**args)
        t[mode] = (time.perf_counter() - t0) / 3
    assert t[0] < 0.5 * t[-1], t


def test_ineligible_problems_fall_back():
    rng = np.random.default_rng(0)
    T = np.float64
    A = np.asfortranarray(rng.standard_normal((60, 120)))
    b = rng.standard_normal(60)
    _mode(0)
    s = pa.FastForwardBackward(tol=1e-6, maxit=50, driver="native")
    s(x0=np.zeros(120), f=pa.LeastSquares(A, b), g=pa.NormL21(0.1, 4))      # group prox: one kernel per operation
    assert s.last_persistent_ctas == 0
    big = np.asfortranarray(rng.standard_normal((1100, 1000)))              # 8.8 MB > the cache-resident limit
    s(x0=np.zeros(1000), f=pa.LeastSquares(big, rng.standard_normal(1100)), g=pa.NormL1(0.1))
    assert s.last_persistent_ctas == 0
    _ = (C, o, T)
