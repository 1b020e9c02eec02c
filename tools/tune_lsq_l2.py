#!/usr/bin/env python
"""Experiment: does the gradient pass of a block-diagonal least-squares term hit L2 when the blocks are processed in small groups
(residual of group g, then gradient of group g) instead of residual over ALL blocks, then gradient over ALL blocks?  configs[1] shape:
100 blocks of 100 x 1e5 fp32 (40 MB each, 4 GB in total; L2 = 126 MB).  Uses the existing kernels on sub-ranges of the blocks.
-> gpurun_out/tune_lsq_l2.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context  # noqa: E402
from tune_step import timeit  # noqa: E402

import ctypes as C  # noqa: E402


def main():
    ctx = Context.get()
    lib, h = ctx.lib, ctx.h
    res = []
    for nblk, mb, nb in ((100, 100, 100_000), (100, 128, 100_000), (1000, 100, 10_000)):
        A = torch.randn(nblk, nb, mb, device="cuda") * 0.1
        b = torch.randn(nblk * mb, device="cuda")
        x = torch.randn(nblk * nb, device="cuda")
        r = torch.empty_like(b)
        grad = torch.empty_like(x)
        es = 4
        blk_bytes = mb * nb * es

        def both(G):
            for k0 in range(0, nblk, G):
                g_ = min(G, nblk - k0)
                Ap = C.c_void_p(A.data_ptr() + k0 * blk_bytes)
                xp = C.c_void_p(x.data_ptr() + k0 * nb * es)
                bp = C.c_void_p(b.data_ptr() + k0 * mb * es)
                rp = C.c_void_p(r.data_ptr() + k0 * mb * es)
                gp = C.c_void_p(grad.data_ptr() + k0 * nb * es)
                L.check(lib.pb_lsq_blockdiag_residual(h, L.PB_F32, g_, mb, nb, Ap, xp, bp, rp))
                L.check(lib.pb_lsq_blockdiag_gradient(h, L.PB_F32, g_, mb, nb, Ap, rp, gp))

        byt = A.numel() * es
        for G in (nblk, 50, 10, 4, 3, 2, 1):
            if G > nblk:
                continue
            ms = timeit(lambda: both(G), reps=5, warm=2)
            res.append(dict(nblk=nblk, mb=mb, nb=nb, group=G, group_mb=G * blk_bytes / 1e6, ms=ms, gbs_if_two_passes=2 * byt / ms / 1e6, gbs_if_one_pass=byt / ms / 1e6))
            print(res[-1], flush=True)
        # the same sequence replayed from a CUDA graph (no launch overhead on the host)
        for G in (4, 2, 1):
            s = torch.cuda.Stream()
            # library kernels run on the context's stream = torch's current stream at context creation: capture on that stream
            g = torch.cuda.CUDAGraph()
            try:
                with torch.cuda.graph(g, stream=torch.cuda.current_stream()):
                    both(G)
                ms = timeit(lambda: g.replay(), reps=5, warm=2)
                res.append(dict(nblk=nblk, mb=mb, nb=nb, group=G, graph=True, ms=ms, gbs_if_one_pass=byt / ms / 1e6))
                print(res[-1], flush=True)
            except Exception as e:          # capture needs a non-default stream; report and move on
                print("graph capture failed:", repr(e)[:200], flush=True)
                break
            del s
        del A, b, x, r, grad
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_lsq_l2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
