"""Single-process, multi-context form of the row-sharded path (SURVEY.md section 8b "Threading": ONE host process drives all GPUs,
one host thread per context; `pb_xchg_connect_local`).  The contexts may share a device, so this runs -- and is observed -- on a
one-GPU box: P shards of a vector, each on its own context and stream, exchange their scalar blocks inside the step kernels and
every context ends up with the same P rows; a whole sharded pb_solve (one host thread per context) takes the same iterations and
produces the same bits as the unsharded solve.  On a box with >= 2 GPUs the contexts are spread over the devices (peer access)."""
import ctypes as C
import threading

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Scalars, shard_bounds  # noqa: E402


class World:
    """P raw contexts (own streams) connected in-process."""

    def __init__(self, P, spread=True):
        self.lib = L.lib()
        ndev = torch.cuda.device_count()
        self.P = P
        self.devs = [(r % ndev) if spread else 0 for r in range(P)]
        self.h = []
        for r in range(P):
            h = C.c_void_p()
            L.check(self.lib.pb_ctx_create(self.devs[r], None, 0, C.byref(h)))
            self.h.append(h)
        for r in range(P):
            L.check(self.lib.pb_xchg_init(self.h[r], r, P, None))
        arr = (C.c_void_p * P)(*[h.value for h in self.h])
        L.check(self.lib.pb_xchg_connect_local(arr, P))
        for r in range(P):
            L.check(self.lib.pb_ctx_set_option(self.h[r], L.PB_OPT_FUSED_EXCHANGE, 1))

    def close(self):
        for h in self.h:
            L.check(self.lib.pb_ctx_destroy(h))

    def wait(self, r):
        rows = (C.c_double * (self.P * L.PB_NSCALARS))()
        L.check(self.lib.pb_exchange_wait(self.h[r], rows, 20.0))
        return np.frombuffer(rows, dtype=np.float64).reshape(self.P, L.PB_NSCALARS).copy()


def _p(t):
    return C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("T", [np.float64, np.float32])
@pytest.mark.parametrize("P", [2, 3, 8])
def test_sharded_step_scalars_equal_the_unsharded_ones(T, P):
    """One fused FISTA step on P row shards (in-kernel exchange) vs the same step on the whole vector: element-wise outputs equal, and
    the combined reductions equal BIT FOR BIT -- for float data too (pack-wise double-double accumulation, step_common.cuh)."""
    rng = np.random.default_rng(P)
    n = 3_000_017
    x, g, zp = (rng.standard_normal(n).astype(T) for _ in range(3))
    dt = L.PB_F32 if T == np.float32 else L.PB_F64
    desc = L.pb_prox(L.PB_PROX_L1, 0, 0.7, 0.0, None, None)
    w1 = World(1)
    try:
        dev0 = torch.device("cuda", w1.devs[0])
        xd, gd, zd = (torch.as_tensor(a).to(dev0) for a in (x, g, zp))
        z1, xn1 = torch.empty_like(xd), torch.empty_like(xd)
        torch.cuda.synchronize()
        L.check(w1.lib.pb_ffb_step(w1.h[0], dt, n, _p(xd), _p(gd), _p(zd), 0.1, 0.5, C.byref(desc), None, _p(z1), None, _p(xn1)))
        whole = Scalars(w1.wait(0))
    finally:
        w1.close()
    w = World(P)
    try:
        bounds = shard_bounds(n, P)
        parts = []
        for r, (lo, hi) in enumerate(bounds):
            dev = torch.device("cuda", w.devs[r])
            parts.append(tuple(torch.as_tensor(a[lo:hi]).to(dev) for a in (x, g, zp)) + (torch.empty(hi - lo, dtype=xd.dtype, device=dev), torch.empty(hi - lo, dtype=xd.dtype, device=dev)))
        torch.cuda.synchronize()
        for r, (lo, hi) in enumerate(bounds):          # launch on every context first (asynchronous), then wait
            xs, gs, zs, zo, xo = parts[r]
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_ffb_step(w.h[r], dt, hi - lo, _p(xs), _p(gs), _p(zs), 0.1, 0.5, C.byref(desc), None, _p(zo), None, _p(xo)))
        rows = []
        for r in range(P):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            rows.append(w.wait(r))
        torch.cuda.set_device(0)
        for r in range(1, P):
            assert np.array_equal(rows[0], rows[r])     # every context received the same P rows
        comb = Scalars(rows[0])
        assert (comb.gsum, comb.res_sq, comb.gdr, comb.res_inf) == (whole.gsum, whole.res_sq, whole.gdr, whole.res_inf)
        zcat = torch.cat([p_[3].cpu() for p_ in parts])
        xcat = torch.cat([p_[4].cpu() for p_ in parts])
        assert torch.equal(zcat, z1.cpu()) and torch.equal(xcat, xn1.cpu())
    finally:
        w.close()


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("alg,adaptive", [(L.PB_ALG_FFB, 0), (L.PB_ALG_FFB, 1), (L.PB_ALG_FB, 1)])
def test_sharded_native_solve_in_one_process_equals_unsharded(P, alg, adaptive):
    """pb_solve on P contexts from P host threads of ONE process (box-constrained least squares toward b, f = SquaredDistance):
    same iteration count, same backtracks, bit-identical solution as one context on the whole vector."""
    T = np.float64
    rng = np.random.default_rng(7)
    n = 400_003
    b = rng.standard_normal(n).astype(T)
    x0 = rng.standard_normal(n).astype(T)

    def run(world, r, lo, hi, out, ready):
        dev = torch.device("cuda", world.devs[r])
        m = hi - lo
        bd = torch.as_tensor(b[lo:hi]).to(dev)
        x = torch.as_tensor(x0[lo:hi]).to(dev)
        bufs = [torch.empty(m, dtype=torch.float64, device=dev) for _ in range(8)]
        grad, z, zprev, xnext, gradz, scratch, sx, sz = bufs
        torch.cuda.synchronize(dev)
        # every rank finishes its set-up (allocations, copies, device-wide synchronisation) BEFORE any rank enters the solve: a peer
        # spinning inside an exchange must never be waited for by a device-wide operation of another rank (same rule as for NCCL)
        ready.wait()
        f = L.pb_smooth(L.PB_F_SQDIST, 0, 0, m, 0, 0, 0, 0, None, bd.data_ptr(), None)
        g = L.pb_prox(L.PB_PROX_BOX, 0, -0.5, 0.5, None, None)
        pipelined = alg == L.PB_ALG_FFB and not adaptive
        o = L.pb_solve_opts(alg, adaptive, L.PB_SEQ_ADAPTIVE, 0, 200, n, 1e-9, 0.0 if adaptive else 0.9, 0.0, 0.0, 1e-7, 0.5, 1.0,
                            sx.data_ptr() if pipelined else None, sz.data_ptr() if pipelined else None, scratch.data_ptr() if pipelined else None)
        res = L.pb_solve_result()
        rc = world.lib.pb_solve(world.h[r], L.PB_F64, m, C.byref(f), C.byref(g), C.byref(o), _p(x), _p(grad), _p(z), _p(zprev), _p(xnext),
                                _p(gradz), _p(scratch), C.byref(res))
        if rc != 0:
            out[r] = RuntimeError(world.lib.pb_last_error().decode())
            ready.abort()
            return
        keep = {t.data_ptr(): t for t in [x] + bufs}
        ready.wait()
        out[r] = (int(res.iterations), int(res.backtracks), res.gamma, res.f_x, res.res_inf, keep[res.z].cpu().numpy().copy())

    def solve(P_):
        w = World(P_)
        try:
            out = [None] * P_
            ready = threading.Barrier(P_)
            ths = [threading.Thread(target=run, args=(w, r, lo, hi, out, ready)) for r, (lo, hi) in enumerate(shard_bounds(n, P_))]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            for o_ in out:
                if isinstance(o_, Exception):
                    raise o_
            return out
        finally:
            w.close()

    whole = solve(1)[0]
    shards = solve(P)
    for s in shards:
        assert s[:5] == whole[:5], (s[:5], whole[:5])
    assert np.array_equal(np.concatenate([s[5] for s in shards]), whole[5])
    assert 1 < whole[0] <= 200
