#!/usr/bin/env python
"""Robust A/B of the fused-step launch shapes: candidates are timed in interleaved rounds (so drift hits all alike) and the
median over rounds is reported.  -> gpurun_out/tune_step2.json"""
import ctypes as C
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import Context, ptr  # noqa: E402
from tune_step import timeit  # noqa: E402

PEAK = 6567.4


def main():
    n = int(float(os.environ.get("TUNE_N", "1e8")))
    rounds = int(os.environ.get("TUNE_ROUNDS", "5"))
    ctx = Context.get()
    out = []
    for T, dt, tdt in ((np.float32, L.PB_F32, torch.float32), (np.float64, L.PB_F64, torch.float64)):
        nn = n if T == np.float32 else n // 2
        es = 4 if T == np.float32 else 8
        x, g, zp = (torch.randn(nn, device="cuda", dtype=tdt) for _ in range(3))
        z, xn = torch.empty_like(x), torch.empty_like(x)
        for kind, pname in ((L.PB_PROX_L1, "l1"), (L.PB_PROX_BOX, "box")):
            desc = L.pb_prox(kind, 0, 1.0 if kind == L.PB_PROX_L1 else -1.0, 1.0, None, None)
            for extrap in (True, False):
                if extrap:
                    fn = lambda: L.check(ctx.lib.pb_ffb_step(ctx.h, dt, nn, ptr(x), ptr(g), ptr(zp), 0.1, 0.5, C.byref(desc), None, ptr(z), None, ptr(xn)))
                    nbytes = 5 * es * nn
                else:
                    fn = lambda: L.check(ctx.lib.pb_fb_step(ctx.h, dt, nn, ptr(x), ptr(g), 0.1, C.byref(desc), None, ptr(z), None))
                    nbytes = 3 * es * nn
                cands = [(1, c, h, u) for c in (1, 2, 3, 4) for h in (0, 1) for u in (2, 4)] + [(1, 2, 1, 8), (1, 1, 1, 8)] + [(2, c, 0, 0) for c in (1, 2, 3)]
                if pname == "box":
                    cands = [(1, 2, 1, 4), (1, 2, 0, 2), (1, 3, 1, 4), (1, 4, 1, 2), (2, 2, 0, 0), (2, 3, 0, 0)]
                times = {c: [] for c in cands}
                for _ in range(rounds):
                    for c in cands:
                        ctx.set_launch(c[1], c[2], c[3], c[0])
                        times[c].append(timeit(fn, reps=20, warm=2))
                for c in cands:
                    ms = float(np.median(times[c]))
                    out.append(dict(dtype=T.__name__, prox=pname, extrap=extrap, impl=c[0], ctas=c[1], hint=c[2], unroll=c[3], ms=ms,
                                    ms_min=float(min(times[c])), gbs=nbytes / ms / 1e6, frac=nbytes / ms / 1e6 / PEAK))
                best = sorted([o for o in out if o["dtype"] == T.__name__ and o["prox"] == pname and o["extrap"] == extrap], key=lambda o: o["ms"])[:4]
                print(T.__name__, pname, "extrap" if extrap else "plain", [(b["impl"], b["ctas"], b["hint"], b["unroll"], round(b["gbs"]), round(b["frac"], 3)) for b in best], flush=True)
        ctx.set_launch()
        del x, g, zp, z, xn
        torch.cuda.empty_cache()
    json.dump(out, open(os.path.join(ROOT, "gpurun_out", "tune_step2.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
