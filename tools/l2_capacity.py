#!/usr/bin/env python
"""How many MB can be re-read from L2?  Repeatedly reduce (torch.sum) a buffer of S bytes: while S fits the effective L2 capacity the
passes after the first run at L2 speed, beyond it at HBM speed.  Informs the fused block-diagonal product."""
import torch

for mb in (16, 32, 48, 56, 64, 72, 80, 96, 112, 128, 160, 256, 1024):
    x = torch.ones(mb * 1024 * 1024 // 4, device="cuda")
    for _ in range(3):
        x.sum()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 30
    e0.record()
    for _ in range(reps):
        x.sum()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print(f"{mb:5d} MB  {ms*1e3:8.1f} us  {mb * 1.048576 / ms:8.1f} GB/s", flush=True)
    del x
