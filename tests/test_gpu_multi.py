"""The persistent multi-iteration step kernel (csrc/step_multi.cu: fixed-stepsize FastForwardBackward with an element-wise gradient
source runs as ONE kernel that loops over the iterations, service CTA + streaming CTAs, in-kernel stop test and exchange) against the
one-launch-per-iteration loop of pb_solve: same iteration counts, bit-identical iterates, scalars and final state -- single GPU, and
two row shards driven from one process."""
import ctypes as C
import itertools
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, shard_bounds  # noqa: E402


def _mode(v):
    c = Context.get()
    L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_MULTI_ITER, v))


@pytest.fixture(autouse=True)
def _restore():
    yield
    _mode(0)


def _run(mode, tol, maxit, **kw):
    _mode(mode)
    s = pa.FastForwardBackward(tol=tol, maxit=maxit, driver="native")
    if "extrapolation_sequence" in kw:
        kw = dict(kw, extrapolation_sequence=kw["extrapolation_sequence"]())
    z, k = s(**kw)
    st = s.last_state
    return (z, k, s.last_multi_iter_kernel, (float(st.gamma), float(st.f_x), float(st.g_z), float(st.res_norm_inf)),
            (st.x.clone(), st.z.clone(), st.z_prev.clone(), st.grad_f_x.clone()), dict(s.last_parity))


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("n", [1000, 200_003, 3_000_000])
def test_one_kernel_for_all_iterations_same_bits(T, n):
    rng = np.random.default_rng(n)
    c = torch.as_tensor(rng.standard_normal(n).astype(T)).cuda()
    b = torch.as_tensor(rng.standard_normal(n).astype(T)).cuda()
    xs = rng.standard_normal(n).astype(T)
    lo_v = torch.as_tensor((-0.5 - 0.1 * rng.random(n)).astype(T)).cuda()
    hi_v = torch.as_tensor((0.5 + 0.1 * rng.random(n)).astype(T)).cuda()
    cases = [
        (dict(x0=xs, f=pa.LinearFunction(c), g=pa.NormL1(T(1)), gamma=T(0.1), extrapolation_sequence=lambda: itertools.repeat(T(0.5))), T(-1), 23),
        (dict(x0=xs, f=pa.LinearFunction(c), g=pa.IndBox(T(-0.5), T(0.5)), gamma=T(0.1)), T(1e-5), 400),
        (dict(x0=xs, f=pa.SquaredDistance(b), g=pa.NormL1(T(0.3)), gamma=T(0.7)), T(1e-5), 500),
        (dict(x0=xs, f=pa.SquaredDistance(b), g=pa.IndBox(lo_v, hi_v), gamma=T(0.9), extrapolation_sequence=lambda: pa.FixedNesterovSequence(T)), T(1e-6), 300),
        (dict(x0=xs, f=pa.SquaredDistance(b), g=pa.Zero(), gamma=T(0.5), extrapolation_sequence=lambda: pa.SimpleNesterovSequence(T)), T(1e-4), 300),
        (dict(x0=xs, f=pa.SquaredDistance(b), g=pa.NormL1(T(0.3)), gamma=T(0.7), mf=T(0.5)), T(-1), 1),
        (dict(x0=xs, f=pa.SquaredDistance(b), g=pa.NormL1(T(0.3)), gamma=T(0.7)), T(-1), 2),
        (dict(x0=xs, f=pa.SquaredDistance(b), g=pa.NormL1(T(0.3)), gamma=T(0.7)), T(-1), 3),
    ]
    ctx = Context.get()
    stopped = []
    for kw, tol, maxit in cases:
        ref = _run(-1, tol, maxit, **kw)
        l0 = ctx.launches()
        got = _run(0, tol, maxit, **kw)
        assert not ref[2] and got[2]
        assert ctx.launches() - l0 <= 4, "iterations must not be separate launches"
        assert got[1] == ref[1] and 1 <= got[1] <= maxit
        assert np.array_equal(got[0], ref[0], equal_nan=True)
        assert got[3] == ref[3], (got[3], ref[3])
        for u, v in zip(got[4], ref[4]):
            assert torch.equal(u, v)
        assert got[5] == ref[5]
        stopped.append(tol > 0 and got[1] < maxit)
    assert sum(stopped) >= 2, "some runs must have been ended by the in-kernel stop test"


def test_two_shards_in_one_process_equal_the_unsharded_solve():
    """P = 2 contexts on one GPU, each running the persistent kernel on half of the SMs' slots (ctas_per_sm = 1), exchanging their scalar
    blocks inside the kernels every iteration: same iteration count and bits as one context on the whole vector."""
    T = np.float64
    rng = np.random.default_rng(11)
    n = 600_007
    b = rng.standard_normal(n).astype(T)
    x0 = rng.standard_normal(n).astype(T)
    lib = L.lib()

    def world(P):
        hs = []
        for r in range(P):
            h = C.c_void_p()
            L.check(lib.pb_ctx_create(0, None, 0, C.byref(h)))
            hs.append(h)
        for r in range(P):
            L.check(lib.pb_xchg_init(hs[r], r, P, None))
        L.check(lib.pb_xchg_connect_local((C.c_void_p * P)(*[h.value for h in hs]), P))
        for h in hs:
            L.check(lib.pb_ctx_set_option(h, L.PB_OPT_FUSED_EXCHANGE, 1))
            L.check(lib.pb_ctx_set_option(h, L.PB_OPT_MULTI_ITER, 1))          # contexts share the GPU: force it ...
            L.check(lib.pb_ctx_set_option(h, L.PB_OPT_CTAS_PER_SM, 1))         # ... and leave room for the peer's CTAs
        return hs

    def run(hs, r, lo, hi, out, ready):
        m = hi - lo
        bd = torch.as_tensor(b[lo:hi]).cuda()
        x = torch.as_tensor(x0[lo:hi]).cuda()
        bufs = [torch.empty(m, dtype=torch.float64, device="cuda") for _ in range(8)]
        grad, z, zprev, xnext, gradz, scratch, sx, sz = bufs
        torch.cuda.synchronize()
        ready.wait()
        f = L.pb_smooth(L.PB_F_SQDIST, 0, 0, m, 0, 0, 0, 0, None, bd.data_ptr(), None)
        g = L.pb_prox(L.PB_PROX_L1, 0, 0.3, 0.0, None, None)
        o = L.pb_solve_opts(L.PB_ALG_FFB, 0, L.PB_SEQ_ADAPTIVE, 0, 400, n, 1e-7, 0.7, 0.0, 0.0, 1e-7, 0.5, 1.0, sx.data_ptr(), sz.data_ptr(), scratch.data_ptr())
        res = L.pb_solve_result()
        rc = lib.pb_solve(hs[r], L.PB_F64, m, C.byref(f), C.byref(g), C.byref(o), C.c_void_p(x.data_ptr()), C.c_void_p(grad.data_ptr()),
                          C.c_void_p(z.data_ptr()), C.c_void_p(zprev.data_ptr()), C.c_void_p(xnext.data_ptr()), C.c_void_p(gradz.data_ptr()),
                          C.c_void_p(scratch.data_ptr()), C.byref(res))
        if rc != 0:
            out[r] = RuntimeError(lib.pb_last_error().decode())
            ready.abort()
            return
        keep = {t.data_ptr(): t for t in [x] + bufs}
        ready.wait()
        out[r] = (int(res.iterations), res.gamma, res.f_x, res.g_z, res.res_inf, res.res_sq, res.gdr, int(res.multi_iter_kernel), keep[res.z].cpu().numpy().copy())

    def solve(P):
        hs = world(P)
        try:
            out = [None] * P
            ready = threading.Barrier(P)
            ths = [threading.Thread(target=run, args=(hs, r, lo, hi, out, ready)) for r, (lo, hi) in enumerate(shard_bounds(n, P))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            for o_ in out:
                if isinstance(o_, Exception):
                    raise o_
            return out
        finally:
            for h in hs:
                L.check(lib.pb_ctx_destroy(h))

    whole = solve(1)[0]
    shards = solve(2)
    assert whole[7] == 1 and all(s[7] == 1 for s in shards)
    for s in shards:
        assert s[:7] == whole[:7], (s[:7], whole[:7])
    assert np.array_equal(np.concatenate([s[8] for s in shards]), whole[8])
    assert 1 < whole[0] < 400
