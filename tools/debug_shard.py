"""2-GPU debug: where does a block-sharded solve diverge from the single-GPU one?"""
import os, sys
import numpy as np, torch
import torch.distributed as dist
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa
from proxb200.host import Context, DeviceExchangeComm, LocalComm, TorchDistComm

rank = int(os.environ["RANK"]); world = int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
ctx = Context.get()
comm = TorchDistComm()
rng = np.random.default_rng(7)
nblk, mb, nb = 8, 16, 64
blocks = rng.standard_normal((nblk, mb, nb)) / np.sqrt(mb)
b = rng.standard_normal(nblk * mb)
lam = 0.1 * np.max(np.abs(np.einsum("bij,bi->bj", blocks, b.reshape(nblk, mb))))
per = nblk // world
f_sh = pa.BlockDiagLeastSquares.from_numpy(blocks[rank * per:(rank + 1) * per], b[rank * per * mb:(rank + 1) * per * mb], comm=comm)
f_1 = pa.BlockDiagLeastSquares.from_numpy(blocks, b)
n = nblk * nb
sl = slice(rank * per * nb, (rank + 1) * per * nb)
for name, mk, kw in (("ffb_fixed", pa.FastForwardBackwardIteration, dict(Lf=4.0)), ("fb_fixed", pa.ForwardBackwardIteration, dict(Lf=4.0)),
                     ("ffb_adaptive", pa.FastForwardBackwardIteration, {}), ("fb_adaptive", pa.ForwardBackwardIteration, {})):
    it_s = iter(mk(x0=np.zeros(per * nb), f=f_sh, g=pa.NormL1(lam), comm=comm, n_global=n, **kw))
    it_1 = iter(mk(x0=np.zeros(n), f=f_1, g=pa.NormL1(lam), **kw))
    first = None
    for k in range(5000):
        ss, s1 = next(it_s), next(it_1)
        if float(s1.res_norm_inf / s1.gamma) <= 1e-7 or float(ss.res_norm_inf / ss.gamma) <= 1e-7:
            if rank == 0:
                print(name, 'stop at state', k, 'single', float(s1.res_norm_inf / s1.gamma), 'sharded', float(ss.res_norm_inf / ss.gamma), 'final z equal', bool(torch.equal(ss.z, s1.z[sl])), flush=True)
            break
        dz = float((ss.z - s1.z[sl]).abs().max()); dx = float((ss.x - s1.x[sl]).abs().max()); dg = float((ss.grad_f_x - s1.grad_f_x[sl]).abs().max())
        same_sc = (ss.gamma == s1.gamma, ss.f_x == s1.f_x, ss.g_z == s1.g_z, ss._sc.res_sq == s1._sc.res_sq, ss._sc.gdr == s1._sc.gdr, ss._sc.res_inf == s1._sc.res_inf)
        if first is None and (dz or dx or dg or not all(same_sc)):
            first = k
            if rank == 0:
                print(name, "first divergence at state", k, "dz", dz, "dx", dx, "dgrad", dg, "scalars equal (gamma,f_x,g_z,res_sq,gdr,res_inf):", same_sc,
                      "| f_x", float(ss.f_x), float(s1.f_x), "gdr", ss._sc.gdr, s1._sc.gdr, flush=True)
    if rank == 0 and first is None:
        print(name, "identical for all states", flush=True)
dist.barrier(); dist.destroy_process_group()
