#!/usr/bin/env python
"""A few fixed-stepsize FISTA iterations on the configs[1] shape (100 blocks of 100 x 1e5 fp32) for `ncu -k regex:k_bd_fista`."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa  # noqa: E402

nblk, mb, nb = 100, 100, 100_000
A = torch.randn(nblk, nb, mb, device="cuda") / 10.0
b = torch.randn(nblk * mb, device="cuda")
f = pa.BlockDiagLeastSquares(A, b)
x0 = torch.zeros(nblk * nb, device="cuda")
z, it = pa.FastForwardBackward(maxit=int(sys.argv[1]) if len(sys.argv) > 1 else 6, tol=-1.0)(x0=x0, f=f, g=pa.NormL1(0.5), Lf=1100.0)
torch.cuda.synchronize()
print("iterations", it)
