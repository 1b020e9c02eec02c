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


def _dense_bounds(dt, m, n, P):
    lib = L.lib()
    cc = int(lib.pb_lsq_dense_chunk_cols(dt, m, n))
    nch = (n + cc - 1) // cc
    base, rem = divmod(nch, P)
    out, c = [], 0
    for r in range(P):
        k = base + (1 if r < rem else 0)
        out.append((min(n, c * cc), min(n, (c + k) * cc)))
        c += k
    return out


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("P", [2, 3, 8])
@pytest.mark.parametrize("m,n", [(300, 5000), (40, 3000), (1000, 20000), (7, 100)])
def test_column_sharded_dense_residual_equals_unsharded(T, P, m, n):
    """C2 (csrc/lsq_kernels.cu: k_gemv_n_combine_x): P column shards of a dense A, each on its own context; one kernel per rank combines
    its chunk partials, pushes them to the peers and folds ALL chunks in global chunk order.  r, ||r||^2 (on every rank) and the
    concatenated A_p' r equal the single-context product BIT FOR BIT; ranks without columns (fewer chunks than ranks) take part too."""
    rng = np.random.default_rng(m + n + P)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(T))
    x = rng.standard_normal(n).astype(T)
    b = rng.standard_normal(m).astype(T)
    dt = L.PB_F32 if T == np.float32 else L.PB_F64
    td = torch.float32 if T == np.float32 else torch.float64

    def aux(lib, h):
        row = (C.c_double * L.PB_NSCALARS)()
        L.check(lib.pb_read_scalars(h, row))
        return row[L.PB_S_AUX], row[L.PB_S_AUX + 1]

    w1 = World(1)
    try:
        dev = torch.device("cuda", w1.devs[0])
        Ad = torch.as_tensor(np.ascontiguousarray(A.T)).to(dev)          # (n, m) row-major == column-major A
        xd, bd = torch.as_tensor(x).to(dev), torch.as_tensor(b).to(dev)
        r1, g1 = torch.empty(m, dtype=td, device=dev), torch.empty(n, dtype=td, device=dev)
        L.check(w1.lib.pb_lsq_dense_residual(w1.h[0], dt, m, n, _p(Ad), m, _p(xd), _p(bd), _p(r1)))
        L.check(w1.lib.pb_lsq_dense_gradient(w1.h[0], dt, m, n, _p(Ad), m, _p(r1), _p(g1)))
        aux1 = aux(w1.lib, w1.h[0])
        r1, g1 = r1.cpu(), g1.cpu()
    finally:
        w1.close()
    w = World(P)
    try:
        bounds = _dense_bounds(dt, m, n, P)
        assert bounds[0][0] == 0 and bounds[-1][1] == n
        parts = []
        for r, (lo, hi) in enumerate(bounds):
            dev = torch.device("cuda", w.devs[r])
            Ar = torch.as_tensor(np.ascontiguousarray(A[:, lo:hi].T)).to(dev)
            parts.append((Ar, torch.as_tensor(x[lo:hi]).to(dev), torch.as_tensor(b).to(dev), torch.empty(m, dtype=td, device=dev),
                          torch.empty(hi - lo, dtype=td, device=dev)))
            # warm-up: the context's scratch is allocated before any rank polls (a cudaFree / cudaMalloc of a peer context on the same
            # device must not wait for a spinning kernel)
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_lsq_dense_residual(w.h[r], dt, m, max(hi - lo, 1) if hi > lo else 0, _p(Ar) if hi > lo else None, m,
                                                _p(parts[-1][1]) if hi > lo else None, None, _p(parts[-1][3])))
        torch.cuda.synchronize()
        for rep in range(3):                                  # three exchanges: both parities and a reuse of the first
            for r, (lo, hi) in enumerate(bounds):             # launch on every context first (asynchronous), then wait
                Ar, xr, br, rr, gr = parts[r]
                L.check(w.lib.pb_ctx_make_current(w.h[r]))
                L.check(w.lib.pb_lsq_dense_residual_sharded(w.h[r], dt, m, hi - lo, _p(Ar) if hi > lo else None, m, _p(xr) if hi > lo else None,
                                                            _p(br), _p(rr), n, lo, 0))
                if hi > lo:
                    L.check(w.lib.pb_lsq_dense_gradient(w.h[r], dt, m, hi - lo, _p(Ar), m, _p(rr), _p(gr)))
            for r in range(P):
                L.check(w.lib.pb_ctx_make_current(w.h[r]))
                L.check(w.lib.pb_ctx_sync(w.h[r]))
                assert torch.equal(parts[r][3].cpu(), r1), (rep, r)
            torch.cuda.set_device(0)
            gcat = torch.cat([p_[4].cpu() for p_ in parts])
            assert torch.equal(gcat, g1)
        # AUX = ||r||^2 replicated; with flags = 1 only rank 0 carries it
        for r in range(P):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_lsq_dense_residual_sharded(w.h[r], dt, m, bounds[r][1] - bounds[r][0], _p(parts[r][0]) if bounds[r][1] > bounds[r][0] else None,
                                                        m, _p(parts[r][1]) if bounds[r][1] > bounds[r][0] else None, _p(parts[r][2]), _p(parts[r][3]), n,
                                                        bounds[r][0], 0))
        for r in range(P):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            assert aux(w.lib, w.h[r]) == aux1
        for r in range(P):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_lsq_dense_residual_sharded(w.h[r], dt, m, bounds[r][1] - bounds[r][0], _p(parts[r][0]) if bounds[r][1] > bounds[r][0] else None,
                                                        m, _p(parts[r][1]) if bounds[r][1] > bounds[r][0] else None, _p(parts[r][2]), _p(parts[r][3]), n,
                                                        bounds[r][0], 1))
        for r in range(P):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            assert aux(w.lib, w.h[r]) == (aux1 if r == 0 else (0.0, 0.0))
        torch.cuda.set_device(0)
        # misaligned shard boundaries are refused
        if bounds[0][1] > 1 and bounds[0][1] < n:
            rc = w.lib.pb_lsq_dense_residual_sharded(w.h[0], dt, m, bounds[0][1] - 1, _p(parts[0][0]), m, _p(parts[0][1]), _p(parts[0][2]), _p(parts[0][3]), n, 0, 0)
            assert rc != 0 and b"aligned" in w.lib.pb_last_error()
    finally:
        w.close()


def _warm_combine_x(m, n, T):
    """Load k_gemv_n_combine_x<T> (and the partial-product kernel of this row count) from ONE host thread: launch on both contexts, then wait."""
    dt = L.PB_F32 if T == np.float32 else L.PB_F64
    td = torch.float32 if T == np.float32 else torch.float64
    w = World(2)
    try:
        bounds = _dense_bounds(dt, m, n, 2)
        bufs = []
        for r, (lo, hi) in enumerate(bounds):
            dev = torch.device("cuda", w.devs[r])
            bufs.append((torch.zeros(max(hi - lo, 1), m, dtype=td, device=dev), torch.zeros(max(hi - lo, 1), dtype=td, device=dev),
                         torch.zeros(m, dtype=td, device=dev), torch.empty(m, dtype=td, device=dev)))
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_lsq_dense_residual(w.h[r], dt, m, hi - lo, _p(bufs[-1][0]), m, _p(bufs[-1][1]), None, _p(bufs[-1][3])))
        torch.cuda.synchronize()
        for r, (lo, hi) in enumerate(bounds):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_lsq_dense_residual_sharded(w.h[r], dt, m, hi - lo, _p(bufs[r][0]), m, _p(bufs[r][1]), _p(bufs[r][2]), _p(bufs[r][3]), n, lo, 0))
        for r in range(2):
            L.check(w.lib.pb_ctx_make_current(w.h[r]))
            L.check(w.lib.pb_ctx_sync(w.h[r]))
        torch.cuda.set_device(0)
    finally:
        w.close()


@pytest.mark.parametrize("P", [2, 4])
@pytest.mark.parametrize("alg,adaptive", [(L.PB_ALG_FFB, 1), (L.PB_ALG_FFB, 0), (L.PB_ALG_FB, 1)])
def test_column_sharded_dense_lasso_solve_equals_unsharded(P, alg, adaptive):
    """pb_solve on a dense Lasso, A column-sharded over P contexts of one process (one host thread each): the in-kernel all-gather of the
    chunk partials makes f, every line-search decision, the iteration count and the solution bit-identical to the unsharded solve."""
    T = np.float64
    rng = np.random.default_rng(11)
    m, n = 120, 3000
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(m))
    xt = np.zeros(n)
    xt[rng.choice(n, 40, replace=False)] = rng.standard_normal(40)
    b = A @ xt + 0.01 * rng.standard_normal(m)
    lam = 0.1 * np.max(np.abs(A.T @ b))
    Lf = 1.05 * np.linalg.norm(A, 2) ** 2

    def run(world, r, lo, hi, out, ready):
        dev = torch.device("cuda", world.devs[r])
        k = hi - lo
        Ad = torch.as_tensor(np.ascontiguousarray(A[:, lo:hi].T)).to(dev)
        bd = torch.as_tensor(b).to(dev)
        rd = torch.empty(m, dtype=torch.float64, device=dev)
        x = torch.zeros(k, dtype=torch.float64, device=dev)
        bufs = [torch.empty(max(k, 1), dtype=torch.float64, device=dev) for _ in range(8)]
        grad, z, zprev, xnext, gradz, scratch, sx, sz = bufs
        L.check(world.lib.pb_ctx_make_current(world.h[r]))
        L.check(world.lib.pb_lsq_dense_residual(world.h[r], L.PB_F64, m, k, _p(Ad) if k else None, m, _p(x) if k else None, None, _p(rd)))   # scratch warm-up
        torch.cuda.synchronize(dev)
        ready.wait()
        sharded = world.P > 1
        f = L.pb_smooth(L.PB_F_LSQ_DENSE, 0, m, k, m, lo if sharded else 0, 0, n if sharded else 0, Ad.data_ptr() if k else None, bd.data_ptr(), rd.data_ptr())
        g = L.pb_prox(L.PB_PROX_L1, 0, lam, 0.0, None, None)
        o = L.pb_solve_opts(alg, adaptive, L.PB_SEQ_ADAPTIVE, 0, 400, n, 1e-8, 0.0 if adaptive else 1.0 / Lf, 0.0, 0.0, 1e-7, 0.5, 1.0, None, None, None)
        res = L.pb_solve_result()
        rc = world.lib.pb_solve(world.h[r], L.PB_F64, k, C.byref(f), C.byref(g), C.byref(o), _p(x), _p(grad), _p(z), _p(zprev), _p(xnext),
                                _p(gradz), _p(scratch), C.byref(res))
        if rc != 0:
            out[r] = RuntimeError(world.lib.pb_last_error().decode())
            ready.abort()
            return
        keep = {t.data_ptr(): t for t in [x] + bufs}
        ready.wait()
        out[r] = (int(res.iterations), int(res.backtracks), res.gamma, res.f_x, res.g_z, res.res_inf, keep[res.z].cpu().numpy()[:k].copy())

    def solve(P_, persistent=0):
        w = World(P_)
        try:
            for h in w.h:
                L.check(w.lib.pb_ctx_set_option(h, L.PB_OPT_PERSISTENT, persistent))
            out = [None] * P_
            ready = threading.Barrier(P_)
            bounds = _dense_bounds(L.PB_F64, m, n, P_) if P_ > 1 else [(0, n)]
            ths = [threading.Thread(target=run, args=(w, r, lo, hi, out, ready)) for r, (lo, hi) in enumerate(bounds)]
            for t in ths:
                t.start()
            for t in ths:
                t.join()
            errs = [f"rank {r}: {o_}" for r, o_ in enumerate(out) if isinstance(o_, Exception)]
            if errs:
                raise RuntimeError("; ".join(errs))
            return out
        finally:
            w.close()

    # Contexts that share ONE device also share its CUDA context: the first launch of a kernel loads it (lazy module loading) under a
    # context-wide lock, waiting for the device to go idle -- if that happens in one host thread while its previous kernel polls for a
    # peer whose thread now cannot launch, the exchange times out.  (One context per device -- the real configuration -- has no such
    # coupling.)  So every kernel of the sharded solve is loaded beforehand: the multi-kernel path by an unsharded solve with the
    # persistent solver switched off, the fused combine + all-gather kernel by one single-threaded sharded product.
    whole = solve(1, persistent=-1)[0]
    _warm_combine_x(m, 256, T)
    assert solve(1)[0][:6] == whole[:6]                    # (and the persistent solver agrees bit for bit)
    shards = solve(P)
    for s in shards:
        assert s[:6] == whole[:6], (s[:6], whole[:6])
    assert np.array_equal(np.concatenate([s[6] for s in shards]), whole[6])
    assert 3 < whole[0] <= 400
