#!/usr/bin/env python
"""A few launches of the K10 pass at the configs[4] size, for `ncu --set full` captures, plus a host-loop timing split."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, ptr  # noqa: E402

ctx = Context.get()
side = 8192
bt = torch.randn(side, side, device="cuda")
f = pa.TVSplit(bt, 0.3)
X0 = f.initial_point()
X1 = torch.empty_like(X0)
for _ in range(2):
    L.check(ctx.lib.pb_dr_tv_step(ctx.h, L.PB_F32, side, side, ptr(X0), ptr(f.b), 1.0, 0.3, ptr(X1), None, None, 0, side, None, None))
torch.cuda.synchronize()
if "--loop" in sys.argv:
    del X1
    it = pa.DouglasRachfordIteration(X0, f=f, g=pa.IndConsensus(5), gamma=1.0)
    t0 = time.perf_counter()
    st = it.step(None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(50):
        st = it.step(st)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    y = st.y
    torch.cuda.synchronize()
    t3 = time.perf_counter()
    print(f"init+first step {1e3 * (t1 - t0):.2f} ms; steady step {1e3 * (t2 - t1) / 50:.3f} ms; materialise y,z {1e3 * (t3 - t2):.2f} ms", flush=True)
    x0 = torch.randn(side * side, device="cuda")
    it = pa.DouglasRachfordIteration(x0, f=pa.SqrNormL2(1.0, bt.view(-1)), g=pa.NormL1(0.3), gamma=0.7)
    st = it.step(None)
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    for _ in range(200):
        st = it.step(st)
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    print(f"element-wise DR steady step {1e3 * (t2 - t1) / 200:.3f} ms", flush=True)
