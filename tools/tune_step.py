#!/usr/bin/env python
"""Sweep launch shapes / implementations of the fused step on one B200 and report achieved algorithmic GB/s.
Writes gpurun_out/tune_step.json.  Timing: CUDA events around `reps` back-to-back launches after warm-up; the 2 GB
working set (n = 1e8 fp32, 5 streams) exceeds L2, so every launch streams from HBM."""
import ctypes as C
import json
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


def timeit(fn, reps=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    n = int(float(os.environ.get("TUNE_N", "1e8")))
    ctx = Context.get()
    res = []
    peak = 6567.4
    for T, dt, tdt in ((np.float32, L.PB_F32, torch.float32), (np.float64, L.PB_F64, torch.float64)):
        nn = n if T == np.float32 else n // 2
        x, g, zp = (torch.randn(nn, device="cuda", dtype=tdt) for _ in range(3))
        z, xn = torch.empty_like(x), torch.empty_like(x)
        es = 4 if T == np.float32 else 8
        for kind, name in ((L.PB_PROX_L1, "l1"), (L.PB_PROX_BOX, "box")):
            desc = L.pb_prox(kind, 0, 1.0 if kind == L.PB_PROX_L1 else -1.0, 1.0, None, None)
            for extrap in (True, False):
                if extrap:
                    fn = lambda: L.check(ctx.lib.pb_ffb_step(ctx.h, dt, nn, ptr(x), ptr(g), ptr(zp), 0.1, 0.5, C.byref(desc), None, ptr(z), None, ptr(xn)))
                    nbytes = 5 * es * nn
                else:
                    fn = lambda: L.check(ctx.lib.pb_fb_step(ctx.h, dt, nn, ptr(x), ptr(g), 0.1, C.byref(desc), None, ptr(z), None))
                    nbytes = 3 * es * nn
                configs = []
                if kind == L.PB_PROX_L1 and extrap:
                    for impl in (1, 2):
                        for ctas in ([2, 3, 4, 6, 8] if impl == 1 else [1, 2, 3]):
                            for hint in (0, 1):
                                for unroll in ([1, 2, 4, 8] if impl == 1 else [0]):
                                    configs.append((impl, ctas, hint, unroll))
                else:
                    configs = [(1, 4, 1, 4), (1, 8, 1, 2), (2, 3, 0, 0), (2, 2, 0, 0)]
                for impl, ctas, hint, unroll in configs:
                    ctx.set_launch(ctas, hint, unroll, impl)
                    try:
                        ms = timeit(fn)
                    except Exception as e:  # noqa: BLE001
                        res.append(dict(dtype=T.__name__, prox=name, extrap=extrap, impl=impl, ctas=ctas, hint=hint, unroll=unroll, error=str(e)))
                        continue
                    gbs = nbytes / ms / 1e6
                    res.append(dict(dtype=T.__name__, prox=name, extrap=extrap, impl=impl, ctas=ctas, hint=hint, unroll=unroll, ms=ms, gbs=gbs, frac=gbs / peak))
                    print(res[-1], flush=True)
        ctx.set_launch()
        del x, g, zp, z, xn
        torch.cuda.empty_cache()
    # reference points: torch copy and torch add at the same sizes (library kernels, for context only)
    a = torch.randn(n, device="cuda")
    b = torch.empty_like(a)
    ms = timeit(lambda: b.copy_(a))
    res.append(dict(name="torch.copy_ f32 n=%d" % n, ms=ms, gbs=8 * n / ms / 1e6))
    c = torch.randn(n, device="cuda")
    ms = timeit(lambda: torch.add(a, c, out=b))
    res.append(dict(name="torch.add f32 n=%d" % n, ms=ms, gbs=12 * n / ms / 1e6))
    print(res[-2], res[-1])
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_step.json"), "w"), indent=1)
    best = sorted([r for r in res if r.get("extrap") and r.get("prox") == "l1" and "gbs" in r and r["dtype"] == "float32"], key=lambda r: -r["gbs"])[:5]
    print("BEST ffb l1 f32:", json.dumps(best))


if __name__ == "__main__":
    main()
