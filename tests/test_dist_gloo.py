"""world_size-2 test of the N > 1 host path on CPU (gloo): the per-iteration exchange (all-gather of the scalar block +
rank-ordered double-double fold) gives every rank the same scalars as an unsharded reduction.  No kernels run here: each
rank fills its scalar block with numpy (test code), exactly the layout the kernels write."""
import math
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _pair_sum(values):
    hi, lo = 0.0, 0.0
    for v in values:
        s = hi + v
        bb = s - hi
        lo += (hi - (s - bb)) + (v - bb)
        hi = s
    return hi, lo


def _worker(rank, world, port, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from proxb200 import _lib as L
        from proxb200.host import TorchDistComm, shard_bounds

        rng = np.random.default_rng(123)
        res = rng.standard_normal(n) * 10.0 ** rng.integers(-6, 6, n)
        grad = rng.standard_normal(n)
        z = rng.standard_normal(n)
        lo, hi = shard_bounds(n, world)[rank]

        class FakeCtx:          # stands in for the device context: only the scalar-block tensor is used by exchange()
            scal = torch.zeros(L.PB_NSCALARS, dtype=torch.float64)

        row = FakeCtx.scal.numpy()
        row[L.PB_S_GSUM:L.PB_S_GSUM + 2] = _pair_sum(np.abs(z[lo:hi]).tolist())
        row[L.PB_S_RESSQ:L.PB_S_RESSQ + 2] = _pair_sum((res[lo:hi] ** 2).tolist())
        row[L.PB_S_GDR:L.PB_S_GDR + 2] = _pair_sum((grad[lo:hi] * res[lo:hi]).tolist())
        row[L.PB_S_RESINF] = np.max(np.abs(res[lo:hi])) if hi > lo else 0.0
        comm = TorchDistComm()
        sc = comm.exchange(FakeCtx)
        assert sc.parts.shape == (world, L.PB_NSCALARS) and comm.rank == rank and comm.size == world
        gathered = comm.allgather_vector(torch.full((3,), float(rank)))
        assert gathered.shape == (world, 3) and [float(v) for v in gathered[:, 0]] == list(range(world))
        q.put((rank, sc.gsum, sc.res_sq, sc.gdr, sc.res_inf,
               math.fsum(np.abs(z).tolist()), math.fsum((res ** 2).tolist()), math.fsum((grad * res).tolist()), float(np.max(np.abs(res)))))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [1000, 31, 100_003])
def test_scalar_exchange_world2(n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    out.sort()
    for rank, gsum, res_sq, gdr, res_inf, e_g, e_r, e_d, e_i in out:
        assert gsum == e_g and res_sq == e_r and res_inf == e_i          # exactly rounded, identical on both ranks
        assert abs(gdr - e_d) <= 2 * np.spacing(abs(e_d)) + 1e-25
    assert out[0][1:5] == out[1][1:5]


# ---------------------------------------------------------------------------------------------------------------------
# N > 1 host path of the row-sharded TV iteration (tv.py): neighbour discovery over the process group and the halo-row
# addresses handed to the kernel.  The device buffers are stood in by numpy arrays shared through the object all-gather.
# ---------------------------------------------------------------------------------------------------------------------


def _tv_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from proxb200.tv import halo_plan

        H, W = 13, 8
        bounds = [(0, 5), (5, 13)]
        rows = bounds[rank]
        mine = (rank, rows[0], rows[1] - rows[0], W, f"buf0-of-{rank}", f"buf1-of-{rank}")
        allinfo = [None] * world
        dist.all_gather_object(allinfo, mine)
        prev, nxt = halo_plan(allinfo, rank, H, 4)
        q.put((rank, prev, nxt))
    finally:
        dist.destroy_process_group()


def test_tv_halo_plan_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_tv_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = dict((r, (a, b)) for r, a, b in (q.get(timeout=120) for _ in range(world)))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    W, es = 8, 4
    # boundary between global rows 4 and 5: row 4 is even -> the pair (4, 5) belongs to copy 3
    assert out[0][0] is None and out[1][1] is None
    handles, off = out[0][1]                       # rank 0 reads row 5 = local row 0 of rank 1 (8 rows), copy 3
    assert handles == ("buf0-of-1", "buf1-of-1") and off == (3 * 8 * W) * es
    handles, off = out[1][0]                       # rank 1 reads row 4 = last local row of rank 0 (5 rows), copy 3
    assert handles == ("buf0-of-0", "buf1-of-0") and off == (3 * 5 * W + 4 * W) * es


def test_tv_halo_plan_matches_the_oracle_sharding():
    """The (copy, row) chosen by halo_plan is the one the oracle's sharded prox needs (oracle/tv_oracle.py), for both parities
    of the boundary, and bad tilings are refused."""
    from oracle import tv_oracle as tvo
    from proxb200.tv import halo_plan

    rng = np.random.default_rng(0)
    H, W = 11, 6
    b = rng.standard_normal((H, W)).astype(np.float32)
    X = rng.standard_normal((5, H, W)).astype(np.float32)
    full, _ = tvo.TVSplit(b, np.float32(0.3), (H, W)).prox(X.reshape(-1), np.float32(0.9))
    full = full.reshape(5, H, W)
    for split in (4, 5, 6):
        info = [(0, 0, split, W, "a0", "a1"), (1, split, H - split, W, "b0", "b1")]
        top, bot = np.ascontiguousarray(X[:, :split]), np.ascontiguousarray(X[:, split:])
        (_, off_next) = halo_plan(info, 0, H, 4)[1]
        (_, off_prev) = halo_plan(info, 1, H, 4)[0]
        halo_next = bot.reshape(-1)[off_next // 4: off_next // 4 + W]
        halo_prev = top.reshape(-1)[off_prev // 4: off_prev // 4 + W]
        ft = tvo.TVSplit(b[:split], np.float32(0.3), (split, W), 0, H)
        fb = tvo.TVSplit(b[split:], np.float32(0.3), (H - split, W), split, H)
        ft.halo_next, fb.halo_prev = halo_next, halo_prev
        yt, _ = ft.prox(top.reshape(-1), np.float32(0.9))
        yb, _ = fb.prox(bot.reshape(-1), np.float32(0.9))
        assert np.array_equal(yt.reshape(5, split, W), full[:, :split]) and np.array_equal(yb.reshape(5, H - split, W), full[:, split:])
    with pytest.raises(ValueError):
        halo_plan([(0, 0, 4, W, "a", "a"), (1, 5, 6, W, "b", "b")], 0, H, 4)        # gap
    with pytest.raises(ValueError):
        halo_plan([(0, 0, 4, W, "a", "a"), (1, 4, 6, W, "b", "b")], 0, H, 4)        # does not cover H


# ---------------------------------------------------------------------------------------------------------------------
# N > 1 host path of a whole sharded solve: two gloo ranks, each running the product's FastForwardBackward / ForwardBackward host
# logic on its row shard with the kernels replaced by the numpy ABI emulation (tests/emu_lib.py), exchange = TorchDistComm.
# Adaptive stepsize: the Lipschitz estimate needs n_global, the line search needs f summed over the shards.
# ---------------------------------------------------------------------------------------------------------------------


def _emulated_solver_setup():
    import proxb200 as pa
    from proxb200 import accel, algorithms, functions, host

    from emu_lib import EmuContext

    ctx = EmuContext()
    ctx.scal = torch.from_numpy(ctx.lib.scal)            # TorchDistComm all-gathers the context's scalar-block tensor
    host.Context.get = classmethod(lambda cls, device=None: ctx)

    def check_vec(t_, n=None, dtype=None):
        assert t_.dim() == 1

    for mod in (host, functions, algorithms, accel):
        mod.check_vec = check_vec
    return pa, ctx


def _problem(n=600):
    rng = np.random.default_rng(42)
    b = rng.standard_normal(n) * (rng.random(n) < 0.3) * 3.0
    return b, 0.7


def _solve_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        pa, ctx = _emulated_solver_setup()
        from proxb200.host import TorchDistComm, shard_bounds

        b, lam = _problem()
        lo, hi = shard_bounds(b.size, world)[rank]
        comm = TorchDistComm()
        out = {}
        for name, mk in (("ffb", pa.FastForwardBackward), ("fb", pa.ForwardBackward)):
            z, it = mk(tol=1e-9, maxit=500, driver="python")(x0=np.full(hi - lo, 2.0), f=pa.SquaredDistance(b[lo:hi]), g=pa.NormL1(lam),
                                                           comm=comm, n_global=b.size)
            out[name] = (it, z)
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_sharded_solve_world2_matches_unsharded_host_logic():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solve_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # the same problem unsharded, in a fresh interpreter state of this process's emulation
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from proxb200 import accel, algorithms, functions, host

    saved = (host.Context.__dict__["get"], [(m, m.check_vec) for m in (host, functions, algorithms, accel)])
    try:
        pa, _ = _emulated_solver_setup()
        b, lam = _problem()
        for name, mk in (("ffb", pa.FastForwardBackward), ("fb", pa.ForwardBackward)):
            z1, it1 = mk(tol=1e-9, maxit=500, driver="python")(x0=np.full(b.size, 2.0), f=pa.SquaredDistance(b), g=pa.NormL1(lam))
            z2 = np.concatenate([res[r][name][1] for r in range(world)])
            assert res[0][name][0] == res[1][name][0] == it1 and it1 < 500          # same iteration count on both ranks and unsharded
            assert np.array_equal(z2, z1)                                           # and the same bits
            want = np.sign(b) * np.maximum(np.abs(b) - lam, 0)                      # closed form: soft threshold
            assert np.max(np.abs(z1 - want)) <= 1e-8
    finally:
        host.Context.get = saved[0]
        for m, f in saved[1]:
            m.check_vec = f


# ---------------------------------------------------------------------------------------------------------------------
# N > 1 host path of a COLUMN-sharded dense A (C2, csrc/lsq_kernels.cu: k_gemv_n_combine_x): shard boundaries from
# host.dense_shard_bounds (multiples of the column chunk of the residual order), per-chunk partial products, all-gather, fold of ALL
# chunks in GLOBAL chunk order.  The kernels are stood in by numpy (any deterministic per-chunk partial will do: the property under test
# is the fold rule); the exchange is gloo.  Aligned shards reproduce the unsharded float32 bits, which is what the device path asserts
# on hardware (tests/test_gpu_local_world.py); a fold of per-rank sums in rank order -- the plain all-gather of rank partials -- does not.
# ---------------------------------------------------------------------------------------------------------------------


def _chunk_partials(A, x, lo, hi, cc):
    """float32 partial products of the chunks [lo, hi) (chunk-aligned lo), one row of the result per chunk."""
    out = []
    for c0 in range(lo, hi, cc):
        c1 = min(hi, c0 + cc)
        blk = np.ascontiguousarray(A[:, c0:c1]) * np.ascontiguousarray(x[c0:c1])[None, :]
        s = np.zeros(A.shape[0], np.float32)
        for j in range(c1 - c0):                   # sequential float32 chain per row, like one column lane of the kernel
            s = (s + blk[:, j]).astype(np.float32)
        out.append(s)
    return np.stack(out) if out else np.zeros((0, A.shape[0]), np.float32)


def _fold(parts):
    s = parts[0].copy()
    for p in parts[1:]:
        s = (s + p).astype(np.float32)
    return s


def _dense_worker(rank, world, port, m, n, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from proxb200 import _lib as L
        from proxb200.host import TorchDistComm, dense_shard_bounds

        rng = np.random.default_rng(5)
        A = rng.standard_normal((m, n)).astype(np.float32)
        x = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3, n)).astype(np.float32)
        cc = int(L.lib().pb_lsq_dense_chunk_cols(L.PB_F32, m, n))
        bounds = dense_shard_bounds(np.float32, m, n, world)
        lo, hi = bounds[rank]
        mine = _chunk_partials(A, x, lo, hi, cc)
        nch = (n + cc - 1) // cc
        per = max((b_[1] - b_[0] + cc - 1) // cc for b_ in bounds)
        padded = np.zeros((per, m), np.float32)
        padded[:mine.shape[0]] = mine
        comm = TorchDistComm()
        allp = comm.allgather_vector(torch.from_numpy(padded.reshape(-1))).numpy().reshape(world, per, m)
        chunks = [allp[r][k] for r in range(world) for k in range((bounds[r][1] - bounds[r][0] + cc - 1) // cc)]
        assert len(chunks) == nch
        r_global_order = _fold(chunks)                                        # what k_gemv_n_combine_x does
        rank_sums = [_fold([allp[r][k] for k in range((bounds[r][1] - bounds[r][0] + cc - 1) // cc)]) for r in range(world) if bounds[r][1] > bounds[r][0]]
        r_rank_order = _fold(rank_sums)                                       # the all-gather of rank partials it replaces
        q.put((rank, r_global_order, r_rank_order, cc, bounds))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("m,n", [(40, 3000), (17, 10_000)])
def test_column_sharded_dense_fold_rule_world2(lib_built, m, n):
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_dense_worker, args=(r, world, port, m, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    out = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(5)
    A = rng.standard_normal((m, n)).astype(np.float32)
    x = (rng.standard_normal(n) * 10.0 ** rng.integers(-3, 3, n)).astype(np.float32)
    cc, bounds = out[0][3], out[0][4]
    whole = _fold(list(_chunk_partials(A, x, 0, n, cc)))                      # one GPU: all chunks in order
    assert all(lo % cc == 0 for lo, _ in bounds) and bounds[0][0] == 0 and bounds[-1][1] == n
    for _, r_global, r_rank, _, _ in out:
        assert np.array_equal(r_global, whole)                                # same bits as unsharded, on every rank
    assert np.array_equal(out[0][1], out[1][1]) and np.array_equal(out[0][2], out[1][2])
    assert not np.array_equal(out[0][2], whole)                               # rank sums folded in rank order: deterministic, but other bits
