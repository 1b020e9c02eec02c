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
