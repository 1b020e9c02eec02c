#!/usr/bin/env python
"""TV denoising of an 8192 x 8192 fp32 image in its primal-dual home: Chambolle-Pock through AFBAIteration
(src/algorithms/primal_dual.jl:173-211) with L = 2-D finite differences (K11 pb_fd2d_forward / adjoint), g = 0.5||u - b||^2
(PB_PROX_SQRL2 in the fused step), h = lam*||.||_1 (conjugate prox = box projection).  -> gpurun_out/perf_cp_tv.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200.host import Context  # noqa: E402


def main():
    side = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
    ctx = Context.get()
    gen = torch.Generator(device="cuda").manual_seed(5)
    b = torch.randn(side * side, device="cuda", generator=gen)
    K = 100
    alg = pa.ChambollePock(tol=-1.0, maxit=K)
    kw = dict(x0=torch.zeros(side * side, device="cuda"), y0=torch.zeros(2 * side * side, device="cuda"), g=pa.SqrNormL2(1.0, b), h=pa.NormL1(0.3),
              L=pa.FiniteDifference2D(side, side))
    alg(**kw)
    torch.cuda.synchronize()
    l0 = ctx.launches()
    t0 = time.perf_counter()
    (x, y), it = alg(**kw)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    npx = side * side
    # per iteration (algorithmic, fp32): L'y (read 2, write 1), primal fused step (read x, temp; b; write xbar [+ -FPR]) ~5, L xbar (read 1, write 2),
    # dual conj-prox (read 2+2, write 2) ~6, over-relaxation updates of x (3) and y (6)  ->  ~26 pixel-vectors
    out = dict(side=side, iterations=it, seconds=dt, it_per_s=it / dt, ms_per_iteration=1e3 * dt / it, launches_per_iteration=(ctx.launches() - l0) / it,
               approx_algorithmic_gbs=26 * 4 * npx * it / dt / 1e9)
    print(json.dumps(out), flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "perf_cp_tv.json"), "w"), indent=1)
    _ = np


if __name__ == "__main__":
    main()
