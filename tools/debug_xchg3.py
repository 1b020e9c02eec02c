import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
mode = sys.argv[1]
if mode == "oracle":
    from oracle import fb_oracle  # noqa
if mode == "scipy":
    import scipy.linalg.blas  # noqa
if mode == "callfn":
    import test_gpu_exchange as t
    t.test_device_exchange_world1_matches_memcpy_readback()
    print("callfn: test function passed when called directly")
    sys.exit(0)
import numpy as np, torch
import proxb200 as pa
from proxb200.host import Context, DeviceExchangeComm
from conftest import load_golden
ctx = Context.get(); comm = DeviceExchangeComm(ctx)
d = load_golden("lasso_small")
for alg in ("ffb", "fb"):
    solver = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6)
    zsol, it = solver(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(float(d["lam"])), comm=comm)
    print(mode, alg, it, flush=True)
comm.close()
