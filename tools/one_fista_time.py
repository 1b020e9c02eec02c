#!/usr/bin/env python
"""Time fixed-stepsize FISTA iterations on the configs[1] shape: python tools/one_fista_time.py [mode]  (mode: PB_OPT_LSQ_FISTA value)"""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 0
nblk, mb, nb = 100, 100, 100_000
A = torch.randn(nblk, nb, mb, device="cuda") / 10.0
b = torch.randn(nblk * mb, device="cuda")
f = pa.BlockDiagLeastSquares(A, b)
x0 = torch.zeros(nblk * nb, device="cuda")
ctx = Context.get()
L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_LSQ_FISTA, mode))
s = pa.FastForwardBackward(maxit=200, tol=-1.0)
s.pipeline = os.environ.get("PROXB200_NO_LOOKAHEAD", "0") != "1"      # A/B of the one-iteration look-ahead of the driver loop
s(x0=x0, f=f, g=pa.NormL1(0.5), Lf=1100.0)
torch.cuda.synchronize()
t0 = time.perf_counter()
z, it = s(x0=x0, f=f, g=pa.NormL1(0.5), Lf=1100.0)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
env = {k[10:]: v for k, v in os.environ.items() if k.startswith("PROXB200_LF_")}
env["lookahead"] = s.pipeline
print(f"mode {mode} {env}: {it / dt:.1f} it/s  {1e3 * dt / it:.4f} ms/it", flush=True)
