import ctypes as C, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proxb200 as pa
from proxb200.host import Context, DeviceExchangeComm, LocalComm, ptr
from proxb200 import _lib as L

mode = sys.argv[1] if len(sys.argv) > 1 else "full"
d = np.load(os.path.join(ROOT, "tests/golden/lasso_small.npz"))
ctx = Context.get()
comm = DeviceExchangeComm(ctx)
if mode in ("full", "part1_small", "part1_nocompare"):
    rng = np.random.default_rng(0)
    n = 1_000_003 if mode != "part1_small" else 1003
    x, g, zp = (torch.as_tensor(rng.standard_normal(n).astype(np.float32)).cuda() for _ in range(3))
    z, xn = torch.empty_like(x), torch.empty_like(x)
    desc = L.pb_prox(L.PB_PROX_L1, 0, 0.7, 0.0, None, None)
    for k in range(5):
        L.check(ctx.lib.pb_ffb_step(ctx.h, L.PB_F32, n, ptr(x), ptr(g), ptr(zp), 0.1 + 0.01 * k, 0.5, C.byref(desc), None, ptr(z), None, ptr(xn)))
        sc = comm.exchange(ctx)
        if mode != "part1_nocompare":
            row = ctx.read_scalars()
            assert np.array_equal(sc.parts[0], row)
    L.check(ctx.lib.pb_nrm2sq(ctx.h, L.PB_F32, n, ptr(x)))
    sc = comm.exchange(ctx)
    if mode != "part1_nocompare":
        assert np.array_equal(sc.parts[0], ctx.read_scalars())
cnt = {"k": 0, "diff": 0}
class Spy:
    rank, size = 0, 1
    def exchange(self, c):
        sc = comm.exchange(c)
        cnt["k"] += 1
        if mode.endswith("spy"):
            direct = c.read_scalars()
            if not np.array_equal(sc.parts[0], direct, equal_nan=True):
                cnt["diff"] += 1
                if cnt["diff"] <= 4:
                    print("DIFF at", cnt["k"], sc.parts[0][:11], direct[:11])
        return sc
for alg in ("ffb", "fb"):
    solver = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        zsol, it = solver(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(float(d["lam"])), comm=Spy())
    print(mode, alg, "iterations", it, "exchanges", cnt["k"], "diffs", cnt["diff"], flush=True)
comm.close()
