"""GPU parity of K10 (fused Douglas-Rachford iteration of consensus-form TV denoising, BASELINE.json configs[4]) against
oracle/tv_oracle.py.  TV is absent from the reference, so the oracle is pinned by properties (tests/test_oracle_tv.py) and
parity here is: every output of the pass BIT-EXACT vs the oracle, shard-with-halo == unsharded bit for bit, and (on a box
with >= 2 GPUs) a 2-rank run whose halo rows are read from peer memory over NVLink reproduces the single-GPU iterates."""
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
from oracle import panoc_oracle as po  # noqa: E402
from oracle import tv_oracle as tvo  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import ptr  # noqa: E402

from conftest import ROOT  # noqa: E402
from gpu_util import ctx, dev, dt  # noqa: E402

TYPES = [np.float32, np.float64]


def _oracle_pass(T, f, X, gamma):
    y, _ = f.prox(X, gamma)
    r = (T(2) * y - X).astype(T)
    z, _ = tvo.Consensus(5).prox(r, gamma)
    res = (y - z).astype(T)
    return y, z, res, (X - res).astype(T)


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("H,W", [(1, 1), (1, 8), (2, 2), (7, 5), (8, 8), (9, 12), (33, 64), (64, 1000), (5, 6), (130, 258)])
def test_tv_pass_bit_exact(T, H, W):
    rng = np.random.default_rng(H * 1000 + W)
    n = H * W
    b = rng.standard_normal(n).astype(T)
    X = rng.standard_normal(5 * n).astype(T)
    gamma, lam = T(0.8), T(0.35)
    c = ctx()
    Xd, bd = dev(X), dev(b)
    Xo, Yo, Zo = torch.empty_like(Xd), torch.empty_like(Xd), torch.empty(n, dtype=Xd.dtype, device="cuda")
    L.check(c.lib.pb_dr_tv_step(c.h, dt(T), H, W, ptr(Xd), ptr(bd), float(gamma), float(lam), ptr(Xo), ptr(Yo), ptr(Zo), 0, H, None, None))
    row = c.read_scalars()
    y, z, res, xn = _oracle_pass(T, tvo.TVSplit(b, lam, (H, W)), X, gamma)
    assert np.array_equal(Yo.cpu().numpy(), y)
    assert np.array_equal(Zo.cpu().numpy(), z[:n])
    assert np.array_equal(Xo.cpu().numpy(), xn)
    assert row[L.PB_S_RESINF] == float(np.max(np.abs(res)))
    # nothing materialised: same x_out
    Xo2 = torch.empty_like(Xd)
    L.check(c.lib.pb_dr_tv_step(c.h, dt(T), H, W, ptr(Xd), ptr(bd), float(gamma), float(lam), ptr(Xo2), None, None, 0, H, None, None))
    assert torch.equal(Xo2, Xo)
    assert c.lib.pb_dr_tv_step(c.h, dt(T), H, W, ptr(Xd), ptr(bd), float(gamma), float(lam), ptr(Xd), None, None, 0, H, None, None) == 1   # aliasing


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("H,W,split", [(12, 16, 6), (12, 16, 5), (9, 7, 4), (64, 128, 31)])
def test_tv_shards_with_halo_rows_equal_unsharded(T, H, W, split):
    """Two shards on ONE GPU: the halo pointers address the other shard's buffer, exactly as a peer mapping would."""
    rng = np.random.default_rng(H + W + split)
    b = rng.standard_normal((H, W)).astype(T)
    X = rng.standard_normal((5, H, W)).astype(T)
    gamma, lam = T(1.1), T(0.2)
    c = ctx()
    es = np.dtype(T).itemsize
    full_in, full_out = dev(X.reshape(-1)), torch.empty(5 * H * W, dtype=dev(b.reshape(-1)).dtype, device="cuda")
    L.check(c.lib.pb_dr_tv_step(c.h, dt(T), H, W, ptr(full_in), ptr(dev(b.reshape(-1))), float(gamma), float(lam), ptr(full_out), None, None, 0, H, None, None))
    want = full_out.cpu().numpy().reshape(5, H, W)
    Ht, Hb = split, H - split
    top_in, bot_in = dev(X[:, :split].reshape(-1)), dev(X[:, split:].reshape(-1))
    top_out, bot_out = torch.empty_like(top_in), torch.empty_like(bot_in)
    kc = 3 if (split - 1) % 2 == 0 else 4
    halo_next = C.c_void_p(bot_in.data_ptr() + (kc * Hb * W) * es)                  # row 0 of copy kc of the lower shard
    halo_prev = C.c_void_p(top_in.data_ptr() + (kc * Ht * W + (Ht - 1) * W) * es)   # last row of copy kc of the upper shard
    L.check(c.lib.pb_dr_tv_step(c.h, dt(T), Ht, W, ptr(top_in), ptr(dev(b[:split].reshape(-1))), float(gamma), float(lam), ptr(top_out), None, None, 0, H, None, halo_next))
    L.check(c.lib.pb_dr_tv_step(c.h, dt(T), Hb, W, ptr(bot_in), ptr(dev(b[split:].reshape(-1))), float(gamma), float(lam), ptr(bot_out), None, None, split, H, halo_prev, None))
    assert np.array_equal(top_out.cpu().numpy().reshape(5, Ht, W), want[:, :split])
    assert np.array_equal(bot_out.cpu().numpy().reshape(5, Hb, W), want[:, split:])


@pytest.mark.parametrize("T", TYPES)
def test_tv_denoising_solver(T):
    rng = np.random.default_rng(5)
    H, W = 48, 64
    img = np.zeros((H, W))
    img[10:30, 12:40] = 1.0
    img[30:44, 30:60] = -0.5
    b = (img + 0.15 * rng.standard_normal((H, W))).astype(T)
    lam, gamma = 0.3, T(1.0)
    f = pa.TVSplit(b, lam)
    alg = pa.DouglasRachford(tol=T(1e-4), maxit=4000)
    y, k = alg(x0=f.initial_point(), f=f, g=pa.IndConsensus(5), gamma=gamma)
    y_o, k_o = po.douglas_rachford(np.tile(b.reshape(-1), 5), f=tvo.TVSplit(b, T(lam), (H, W)), g=tvo.Consensus(5), gamma=gamma, tol=T(1e-4), maxit=4000)
    assert k == k_o and np.array_equal(y.cpu().numpy(), y_o)                  # element-wise arithmetic: identical iterates
    u = f.image(y).cpu().numpy()
    assert f.objective(u) < 0.7 * f.objective(b)
    assert np.mean((u - img) ** 2) < 0.35 * np.mean((b - img) ** 2)           # it denoises
    st = alg.last_state
    assert np.max(np.abs(st.y.cpu().numpy().reshape(5, -1) - st.z.cpu().numpy())) <= 2e-3   # consensus


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import proxb200 as pa_
        from proxb200.host import Context as Ctx, DeviceExchangeComm as DX

        comm = DX(Ctx.get())
        rng = np.random.default_rng(9)
        H, W = 96, 128
        b = np.kron(rng.standard_normal((6, 8)), np.ones((16, 16))).astype(np.float32) + 0.1 * rng.standard_normal((H, W)).astype(np.float32)
        rows = [(0, 47), (47, 96)][rank]                 # an odd boundary: the straddling pair is an EVEN vertical pair (copy 3)
        f = pa_.TVSplit(b[rows[0]:rows[1]], 0.3, comm=comm, row0=rows[0], Hglob=H)
        y, k = pa_.DouglasRachford(tol=-1.0, maxit=60)(x0=f.initial_point(), f=f, g=pa_.IndConsensus(5), gamma=np.float32(1.0))
        q.put((rank, k, y.cpu().numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_two_rank_tv_with_nvlink_halo_equals_single_gpu():
    import torch.multiprocessing as mp

    world = 2
    ctxmp = mp.get_context("spawn")
    q = ctxmp.Queue()
    port = _free_port()
    procs = [ctxmp.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = {r: (k, y) for r, k, y in (q.get(timeout=300) for _ in range(world))}
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(9)
    H, W = 96, 128
    b = np.kron(rng.standard_normal((6, 8)), np.ones((16, 16))).astype(np.float32) + 0.1 * rng.standard_normal((H, W)).astype(np.float32)
    f = pa.TVSplit(b, 0.3)
    y1, k1 = pa.DouglasRachford(tol=-1.0, maxit=60)(x0=f.initial_point(), f=f, g=pa.IndConsensus(5), gamma=np.float32(1.0))
    full = y1.cpu().numpy().reshape(5, H, W)
    assert res[0][0] == res[1][0] == k1 == 60
    assert np.array_equal(res[0][1].reshape(5, 47, W), full[:, :47])
    assert np.array_equal(res[1][1].reshape(5, 49, W), full[:, 47:])
