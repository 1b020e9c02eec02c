"""csrc/lsq_fista.cu -- fixed-stepsize FastForwardBackward on a block-diagonal least-squares term with ONE sweep of A per iteration
(gradient, fused step and the next residual's partial products from the same tiles) -- against the residual + gradient + step kernels
(two sweeps): same iteration counts, bit-identical iterates, scalars and final state.  BASELINE.json configs[1] structure."""
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


def _mode(v):
    c = Context.get()
    L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_LSQ_FISTA, v))


@pytest.fixture(autouse=True)
def _restore():
    yield
    _mode(0)


def _problem(T, nblk, mb, nb, seed):
    rng = np.random.default_rng(seed)
    blocks = (rng.standard_normal((nblk, mb, nb)) / np.sqrt(mb)).astype(T)
    xt = np.zeros(nblk * nb)
    idx = rng.choice(nblk * nb, max(4, nblk * nb // 200), replace=False)
    xt[idx] = rng.standard_normal(idx.size)
    b = (np.einsum("bij,bj->bi", blocks.astype(np.float64), xt.reshape(nblk, nb)).reshape(-1) + 0.01 * rng.standard_normal(nblk * mb)).astype(T)
    lam = T(0.1 * np.max(np.abs(np.einsum("bij,bi->bj", blocks.astype(np.float64), b.reshape(nblk, mb).astype(np.float64)))))
    Lf = T(1.05 * max(np.linalg.norm(blocks[k].astype(np.float64), 2) ** 2 for k in range(nblk)))
    return blocks, b, lam, Lf


def _run(mode, f, g, x0, Lf, tol, maxit, **kw):
    _mode(mode)
    s = pa.FastForwardBackward(tol=tol, maxit=maxit, driver="native")
    s.pipeline = True
    z, k = s(x0=x0, f=f, g=g, Lf=Lf, **kw)
    st = s.last_state
    return z, k, (float(st.gamma), float(st.f_x), float(st.g_z), float(st.res_norm_inf)), (st.x.clone(), st.z.clone(), st.z_prev.clone(), st.grad_f_x.clone()), dict(s.last_parity)


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("nblk,mb,nb", [(6, 64, 400), (5, 100, 2000), (3, 128, 5004), (2, 72, 36), (1, 100, 10_000)])
def test_one_sweep_per_iteration_same_bits(T, nblk, mb, nb):
    if T == np.float64 and mb > 128:
        pytest.skip("row packs of a column must fit one CTA")
    blocks, b, lam, Lf = _problem(T, nblk, mb, nb, nblk * 100 + mb)
    f = pa.BlockDiagLeastSquares.from_numpy(blocks, b)
    x0 = np.zeros(nblk * nb, T)
    tol = T(1e-6 if T == np.float64 else 1e-4)
    ctx = Context.get()
    for g in (pa.NormL1(lam), pa.IndBox(T(-0.02), T(0.05)), pa.Zero()):
        for maxit, kw in ((300, {}), (1, {}), (2, {}), (7, dict(extrapolation_sequence=pa.FixedNesterovSequence(T)))):
            ref = _run(-1, f, g, x0, Lf, tol, maxit, **kw)
            l0 = ctx.launches()
            got = _run(1, f, g, x0, Lf, tol, maxit, **kw)
            nl = ctx.launches() - l0
            assert got[1] == ref[1] and 1 <= got[1] <= maxit
            assert nl <= 2 * (got[1] + 1) + 8, "two launches per iteration (sweep + combine), one iteration launched ahead"
            assert np.array_equal(got[0], ref[0], equal_nan=True)
            assert got[2] == ref[2], (got[2], ref[2])
            for u, v in zip(got[3], ref[3]):
                assert torch.equal(u, v)
            assert got[4] == ref[4]


def test_matches_the_oracle_on_a_config1_like_problem():
    T = np.float64
    blocks, b, lam, Lf = _problem(T, 6, 64, 400, 2)
    z_o, it_o = o.fast_forward_backward(np.zeros(6 * 400, T), o.BlockDiagLeastSquares(blocks, b), o.NormL1(lam), tol=1e-6, Lf=Lf)
    _mode(1)
    z, it = pa.FastForwardBackward(tol=1e-6)(x0=np.zeros(6 * 400, T), f=pa.BlockDiagLeastSquares.from_numpy(blocks, b), g=pa.NormL1(lam), Lf=Lf)
    assert it == it_o and np.max(np.abs(z - z_o)) <= 1e-9


@pytest.mark.parametrize("T", [np.float32, np.float64])
def test_look_ahead_through_the_exchange_and_without_it(T):
    """The driver loop launches the sweep of iteration k+1 before it examines iteration k (csrc/solve.cu: run_ffb_bd_fista_ahead).  Three
    ways for the scalars to reach the host -- synchronous read-back (no spare vectors), pinned slots + events (spares, no exchange), the
    device exchange (spares, world 1) -- must stop at the same iteration with the same bits, for stops in the middle and at maxit."""
    from proxb200.host import DeviceExchangeComm

    blocks, b, lam, Lf = _problem(T, 5, 100, 2000, 77)
    f = pa.BlockDiagLeastSquares.from_numpy(blocks, b)
    x0 = np.zeros(5 * 2000, T)
    g = pa.NormL1(lam)
    for tol, maxit in ((T(1e-3), 500), (T(1e-5 if T == np.float64 else 1e-4), 500), (T(-1.0), 7), (T(-1.0), 1)):
        runs = []
        for how in ("sync", "slots", "exchange"):
            _mode(1)
            s = pa.FastForwardBackward(tol=tol, maxit=maxit, driver="native")
            s.pipeline = how != "sync"
            comm = DeviceExchangeComm(Context.get()) if how == "exchange" else None
            try:
                z, k = s(x0=x0, f=f, g=g, Lf=Lf, **({"comm": comm} if comm is not None else {}))
            finally:
                if comm is not None:
                    comm.close()
            st = s.last_state
            runs.append((k, z, float(st.f_x), float(st.g_z), float(st.res_norm_inf), st.x.cpu().numpy().copy(), st.z_prev.cpu().numpy().copy(),
                         st.grad_f_x.cpu().numpy().copy(), dict(s.last_parity)))
        for r in runs[1:]:
            assert r[0] == runs[0][0] and r[2:5] == runs[0][2:5] and r[8] == runs[0][8], (tol, maxit)
            for u, v in zip((r[1],) + r[5:8], (runs[0][1],) + runs[0][5:8]):
                assert np.array_equal(u, v, equal_nan=True)
        assert 1 <= runs[0][0] <= maxit
