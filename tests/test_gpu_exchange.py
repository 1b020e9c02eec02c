"""C1 on the device (csrc/xchg.cuh): the in-kernel exchange delivers exactly the scalar block that the memcpy path reads,
a solve driven by it takes the same iterations, and (on a box with >= 2 GPUs) a 2-rank row-sharded solve reproduces the
single-GPU run bit for bit."""
import ctypes as C
import os
import socket
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, DeviceExchangeComm, ptr  # noqa: E402

from conftest import ROOT, load_golden  # noqa: E402


def test_device_exchange_world1_matches_memcpy_readback():
    ctx = Context.get()
    comm = DeviceExchangeComm(ctx)
    try:
        rng = np.random.default_rng(0)
        n = 1_000_003
        x, g, zp = (torch.as_tensor(rng.standard_normal(n).astype(np.float32)).cuda() for _ in range(3))
        z, xn = torch.empty_like(x), torch.empty_like(x)
        desc = L.pb_prox(L.PB_PROX_L1, 0, 0.7, 0.0, None, None)
        for k in range(5):
            L.check(ctx.lib.pb_ffb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(g), ptr(zp), 0.1 + 0.01 * k, 0.5, C.byref(desc), None, ptr(z), None, ptr(xn)))
            sc = comm.exchange(ctx)                       # fused: pushed by the step kernel's last CTA
            row = ctx.read_scalars()                      # reference path: cudaMemcpy of the device block
            assert sc.parts.shape == (1, L.PB_NSCALARS)
            assert np.array_equal(sc.parts[0], row)
        # a read that does not follow a fused step launches the stand-alone exchange kernel
        L.check(ctx.lib.pb_nrm2sq(ctx.h, L.PB_F32, n, ptr(x)))
        sc = comm.exchange(ctx)
        assert np.array_equal(sc.parts[0], ctx.read_scalars())
        # whole solves through the exchange: same iteration counts as the default path / the oracle
        d = load_golden("lasso_small")
        for alg, want in (("ffb", 788), ("fb", 1251)):
            solver = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6)
            zsol, it = solver(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(float(d["lam"])), comm=comm)
            assert it == want
        zsol2, it2 = pa.FastForwardBackward(tol=1e-6)(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(float(d["lam"])))
        zsol1, it1 = pa.FastForwardBackward(tol=1e-6)(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(float(d["lam"])), comm=comm)
        assert it1 == it2 and np.array_equal(zsol1, zsol2)
    finally:
        comm.close()
    # after close the context is back on the memcpy path
    assert ctx.lib.pb_exchange(ctx.h) == 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, exchange, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import proxb200 as pa_
        from proxb200.host import Context as Ctx, DeviceExchangeComm as DX, TorchDistComm, shard_bounds

        ctx = Ctx.get()
        comm = DX(ctx) if exchange == "device" else TorchDistComm()
        rng = np.random.default_rng(7)
        nblk, mb, nb = 8, 16, 64
        blocks = rng.standard_normal((nblk, mb, nb)) / np.sqrt(mb)
        b = rng.standard_normal(nblk * mb)
        lam = 0.1 * np.max(np.abs(np.einsum("bij,bi->bj", blocks, b.reshape(nblk, mb))))
        Lf = 1.01 * max(np.linalg.norm(blocks[k], 2) ** 2 for k in range(nblk))
        per = nblk // world
        sl_b = slice(rank * per, (rank + 1) * per)
        f = pa_.BlockDiagLeastSquares.from_numpy(blocks[sl_b], b[rank * per * mb:(rank + 1) * per * mb], comm=comm)
        n = nblk * nb
        x0 = np.zeros(per * nb)
        out = {}
        for name, kw in (("ffb_adaptive", {}), ("fb_adaptive", {}), ("ffb_fixed", dict(Lf=Lf))):
            mk = pa_.ForwardBackward if name.startswith("fb") else pa_.FastForwardBackward
            z, it = mk(tol=1e-7, maxit=5000)(x0=x0, f=f, g=pa_.NormL1(lam), comm=comm, n_global=n, **kw)
            out[name] = (it, z)
        if exchange == "device":
            # the one-sweep FISTA kernel (csrc/lsq_fista.cu) forced on this small A: its combine kernel exchanges the scalar block in-kernel
            # and the driver loop runs one iteration ahead through the exchange
            from proxb200 import _lib as L_

            blocks2, b2, lam2, Lf2 = _bd64_problem()
            per2 = blocks2.shape[0] // world
            f2 = pa_.BlockDiagLeastSquares.from_numpy(blocks2[rank * per2:(rank + 1) * per2], b2[rank * per2 * 64:(rank + 1) * per2 * 64], comm=comm)
            L_.check(ctx.lib.pb_ctx_set_option(ctx.h, L_.PB_OPT_LSQ_FISTA, 1))
            alg = pa_.FastForwardBackward(tol=1e-7, maxit=5000)
            z, it = alg(x0=np.zeros(per2 * blocks2.shape[2]), f=f2, g=pa_.NormL1(lam2), comm=comm, n_global=blocks2.shape[0] * blocks2.shape[2], Lf=Lf2)
            out["ffb_fixed_one_sweep"] = (it, z)
            L_.check(ctx.lib.pb_ctx_set_option(ctx.h, L_.PB_OPT_LSQ_FISTA, 0))
        if exchange == "device":
            # row-sharded PANOC + L-BFGS(5): every dot of the two-loop recursion is summed over the ranks inside its kernel
            for name, g_, kw in (("panoc_l1", pa_.NormL1(lam), {}), ("panoc_l21_fixed", pa_.NormL21(lam, 4), dict(Lf=Lf)),
                                 ("panoc_l1_lbfgs2", pa_.NormL1(lam), dict(directions=pa_.LBFGS(2)))):
                alg = pa_.PANOC(tol=1e-7, maxit=2000)
                z, it = alg(x0=x0, f=f, g=g_, comm=comm, n_global=n, **kw)
                out[name] = (it, z, alg.last_iteration.tau_backtracks, alg.last_iteration.backtracks)
        if exchange == "device":
            # dense A, column-sharded: the chunk partials of A x are all-gathered inside the combine kernel (C2), native driver loop
            from proxb200.host import dense_shard_bounds

            Ad, bd, lamd, Lfd = _dense_problem()
            lo, hi = dense_shard_bounds(np.float64, Ad.shape[0], Ad.shape[1], world)[rank]
            fd = pa_.LeastSquares(Ad[:, lo:hi], bd, comm=comm, n_global=Ad.shape[1], col_offset=lo)
            for name, kw in (("dense_ffb_adaptive", {}), ("dense_fb_adaptive", {}), ("dense_ffb_fixed", dict(Lf=Lfd))):
                mk = pa_.ForwardBackward if "fb_" in name and "ffb" not in name else pa_.FastForwardBackward
                alg = mk(tol=1e-7, maxit=3000)
                z, it = alg(x0=np.zeros(hi - lo), f=fd, g=pa_.NormL1(lamd), comm=comm, n_global=Ad.shape[1], **kw)
                out[name] = (it, z, alg.last_driver)
        q.put((rank, out))
        dist.barrier()
    finally:
        dist.destroy_process_group()


def _bd64_problem():
    """Block-diagonal Lasso whose blocks are tall enough (64 rows) for the one-sweep FISTA kernel."""
    rng = np.random.default_rng(33)
    nblk, mb, nb = 4, 64, 400
    blocks = rng.standard_normal((nblk, mb, nb)) / np.sqrt(mb)
    b = rng.standard_normal(nblk * mb)
    lam = 0.1 * np.max(np.abs(np.einsum("bij,bi->bj", blocks, b.reshape(nblk, mb))))
    return blocks, b, lam, 1.05 * max(np.linalg.norm(blocks[k], 2) ** 2 for k in range(nblk))


def _dense_problem():
    rng = np.random.default_rng(21)
    m, n = 150, 4000
    A = np.asfortranarray(rng.standard_normal((m, n)) / np.sqrt(m))
    xt = np.zeros(n)
    xt[rng.choice(n, 30, replace=False)] = rng.standard_normal(30)
    b = A @ xt + 0.01 * rng.standard_normal(m)
    return A, b, 0.1 * np.max(np.abs(A.T @ b)), 1.05 * np.linalg.norm(A, 2) ** 2


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
@pytest.mark.parametrize("exchange", ["device", "nccl"])
def test_two_rank_sharded_solve_equals_single_gpu(exchange):
    import torch.multiprocessing as mp

    world = 2
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, exchange, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = dict(q.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single-GPU run of the same problem in this process
    rng = np.random.default_rng(7)
    nblk, mb, nb = 8, 16, 64
    blocks = rng.standard_normal((nblk, mb, nb)) / np.sqrt(mb)
    b = rng.standard_normal(nblk * mb)
    lam = 0.1 * np.max(np.abs(np.einsum("bij,bi->bj", blocks, b.reshape(nblk, mb))))
    Lf = 1.01 * max(np.linalg.norm(blocks[k], 2) ** 2 for k in range(nblk))
    f = pa.BlockDiagLeastSquares.from_numpy(blocks, b)
    fo = o.BlockDiagLeastSquares(blocks, b)
    for name, kw in (("ffb_adaptive", {}), ("fb_adaptive", {}), ("ffb_fixed", dict(Lf=Lf))):
        mk = pa.ForwardBackward if name.startswith("fb") else pa.FastForwardBackward
        z1, it1 = mk(tol=1e-7, maxit=5000)(x0=np.zeros(nblk * nb), f=f, g=pa.NormL1(lam), **kw)
        z2 = np.concatenate([res[r][name][1] for r in range(world)])
        assert res[0][name][0] == res[1][name][0] == it1          # same iteration count on every rank and as 1 GPU
        assert np.array_equal(z2, z1), (name, float(np.max(np.abs(z2 - z1))), int(np.argmax(np.abs(z2 - z1))))   # and the same bits
        mk_o = o.forward_backward if name.startswith("fb") else o.fast_forward_backward
        z_o, it_o = mk_o(np.zeros(nblk * nb), fo, o.NormL1(lam), tol=1e-7, maxit=5000, **kw)
        assert abs(it1 - it_o) <= max(2, it_o // 100) and np.max(np.abs(z1 - z_o)) <= 1e-8
    if exchange == "device":
        blocks2, b2, lam2, Lf2 = _bd64_problem()
        z1, it1 = pa.FastForwardBackward(tol=1e-7, maxit=5000)(x0=np.zeros(blocks2.shape[0] * blocks2.shape[2]), f=pa.BlockDiagLeastSquares.from_numpy(blocks2, b2),
                                                               g=pa.NormL1(lam2), Lf=Lf2)          # two sweeps per iteration (A is small)
        z2 = np.concatenate([res[r]["ffb_fixed_one_sweep"][1] for r in range(world)])
        assert res[0]["ffb_fixed_one_sweep"][0] == res[1]["ffb_fixed_one_sweep"][0] == it1 and 3 < it1 < 5000
        assert np.array_equal(z2, z1)
        Ad, bd, lamd, Lfd = _dense_problem()
        fd = pa.LeastSquares(Ad, bd)
        for name, kw in (("dense_ffb_adaptive", {}), ("dense_fb_adaptive", {}), ("dense_ffb_fixed", dict(Lf=Lfd))):
            mk = pa.ForwardBackward if "fb_" in name and "ffb" not in name else pa.FastForwardBackward
            z1, it1 = mk(tol=1e-7, maxit=3000)(x0=np.zeros(Ad.shape[1]), f=fd, g=pa.NormL1(lamd), **kw)
            z2 = np.concatenate([res[r][name][1] for r in range(world)])
            assert res[0][name][0] == res[1][name][0] == it1, (name, res[0][name][0], it1)
            assert res[0][name][2] == "native"
            assert np.array_equal(z2, z1), (name, float(np.max(np.abs(z2 - z1))))
            assert 3 < it1 <= 3000
        for name, g_, kw in (("panoc_l1", pa.NormL1(lam), {}), ("panoc_l21_fixed", pa.NormL21(lam, 4), dict(Lf=Lf)),
                             ("panoc_l1_lbfgs2", pa.NormL1(lam), dict(directions=pa.LBFGS(2)))):
            alg = pa.PANOC(tol=1e-7, maxit=2000, driver="python")
            z1, it1 = alg(x0=np.zeros(nblk * nb), f=f, g=g_, **kw)
            z2 = np.concatenate([res[r][name][1] for r in range(world)])
            assert res[0][name][0] == res[1][name][0] == it1, (name, res[0][name][0], it1)
            assert res[0][name][2:] == res[1][name][2:] == (alg.last_iteration.tau_backtracks, alg.last_iteration.backtracks)
            assert np.array_equal(z2, z1), (name, float(np.max(np.abs(z2 - z1))))
            assert 1 < it1 < 2000
