import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa
from proxb200.host import Context, DeviceExchangeComm, LocalComm
from proxb200 import _lib as L

d = np.load(os.path.join(ROOT, "tests/golden/lasso_small.npz"))
ctx = Context.get()

class Spy:
    def __init__(self, inner, mode):
        self.inner, self.mode, self.rank, self.size, self.k, self.diffs = inner, mode, inner.rank, inner.size, 0, 0
    def exchange(self, c):
        sc = self.inner.exchange(c)
        self.k += 1
        if self.mode == "sync":
            torch.cuda.synchronize()
        elif self.mode == "compare":
            direct = c.read_scalars()
            if not np.array_equal(sc.parts[0], direct, equal_nan=True):
                self.diffs += 1
                if self.diffs <= 5:
                    print("  DIFF at exchange", self.k, "xchg", sc.parts[0][:11], "direct", direct[:11])
        elif self.mode == "compare_nosync_first":
            # look at the raw device block WITHOUT waiting for the stream first (async copy on a side stream would still order) -> skip
            pass
        return sc
    def allgather_vector(self, v):
        return self.inner.allgather_vector(v)

def solve(comm):
    return pa.FastForwardBackward(tol=1e-6)(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(1.0), comm=comm)[1]

print("local:", solve(LocalComm()))
dev = DeviceExchangeComm(ctx)
for mode in ("none", "sync", "compare", "none"):
    spy = Spy(dev, mode)
    try:
        it = solve(spy)
    except Exception as e:
        it = repr(e)
    print("device, spy mode", mode, "-> iterations", it, "exchanges", spy.k, "diffs", spy.diffs, flush=True)
# unfused exchange only (k_xchg after every kernel group), fused flag off
L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_FUSED_EXCHANGE, 0))
print("device, fused off:", solve(Spy(dev, "none")))
L.check(ctx.lib.pb_ctx_set_option(ctx.h, L.PB_OPT_FUSED_EXCHANGE, 1))
print("device, fused on again:", solve(Spy(dev, "none")))
dev.close()
