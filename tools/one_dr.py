#!/usr/bin/env python
"""A handful of launches of the K8 pass at the configs[4] size, for `ncu --set full` captures."""
import ctypes as C
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, ptr  # noqa: E402

ctx = Context.get()
npx = 8192 * 8192
b, x0, x1 = (torch.randn(npx, device="cuda") for _ in range(3))
fd, gd = pa.SqrNormL2(1.0, b).descriptor(np.float32), pa.NormL1(0.3).descriptor(np.float32)
fd2, gd2 = pa.NormL1(0.3).descriptor(np.float32), pa.IndBox(-0.5, 0.5).descriptor(np.float32)
for unroll in (1, 2):
    ctx.set_launch(unroll=unroll)
    L.check(ctx.lib.pb_dr_step(ctx.h, L.PB_F32, npx, ptr(x0), 0.7, C.byref(fd), C.byref(gd), ptr(x1), None, None, None, None))
    L.check(ctx.lib.pb_dr_step(ctx.h, L.PB_F32, npx, ptr(x0), 0.7, C.byref(fd2), C.byref(gd2), ptr(x0), None, None, None, None))
torch.cuda.synchronize()
