#!/usr/bin/env python
"""Block-diagonal value+gradient: two kernels (2 sweeps of A from HBM) vs the fused persistent kernel (1 sweep from HBM + 1 from L2,
csrc/lsq_fused.cu) with 1 / 2 / 3 blocks kept between the sweeps; then a whole configs[1] FISTA iteration.  -> gpurun_out/tune_lsq_fused.json"""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, ptr  # noqa: E402
from tune_step import timeit  # noqa: E402


def main():
    ctx = Context.get()
    lib, h = ctx.lib, ctx.h
    res = []
    for nblk, mb, nb in ((100, 100, 100_000), (100, 128, 100_000), (1000, 100, 10_000), (400, 64, 50_000)):
        A = torch.randn(nblk, nb, mb, device="cuda") * 0.1
        b = torch.randn(nblk * mb, device="cuda")
        x = torch.randn(nblk * nb, device="cuda")
        r, grad = torch.empty_like(b), torch.empty_like(x)
        byt = A.numel() * 4
        for mode in (-1, 1, 2, 3):
            L.check(lib.pb_ctx_set_option(h, L.PB_OPT_LSQ_FUSED, mode))
            ms = timeit(lambda: L.check(lib.pb_lsq_blockdiag_value_and_gradient(h, L.PB_F32, nblk, mb, nb, ptr(A), ptr(x), ptr(b), ptr(r), ptr(grad))), reps=10, warm=3)
            res.append(dict(nblk=nblk, mb=mb, nb=nb, mode="two kernels" if mode < 0 else f"fused live={mode}", ms=ms, gbs_of_one_sweep=byt / ms / 1e6,
                            gbs_of_two_sweeps=2 * byt / ms / 1e6))
            print(res[-1], flush=True)
        L.check(lib.pb_ctx_set_option(h, L.PB_OPT_LSQ_FUSED, 0))
        del A, b, x, r, grad
        torch.cuda.empty_cache()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_lsq_fused.json"), "w"), indent=1)
    _ = pa


if __name__ == "__main__":
    main()
