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
    def __init__(self, inner, tag):
        self.inner, self.tag, self.rank, self.size, self.k = inner, tag, inner.rank, inner.size, 0
    def exchange(self, c):
        sc = self.inner.exchange(c)
        direct = c.read_scalars()
        self.k += 1
        if self.k <= 14:
            same = np.array_equal(sc.parts[0], direct)
            print(self.tag, self.k, "same_as_memcpy" if same else "DIFF", [f"{v:.6g}" for v in sc.parts[0][[0, 2, 4, 6, 8]]], "| direct", [f"{v:.6g}" for v in direct[[0, 2, 4, 6, 8]]])
        return sc
    def allgather_vector(self, v):
        return self.inner.allgather_vector(v)

for tag, mk in (("local", lambda: LocalComm()), ("device", lambda: DeviceExchangeComm(ctx))):
    comm = mk()
    it = pa.FastForwardBackwardIteration(x0=np.zeros(100), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(1.0), comm=Spy(comm, tag))
    for k, st in enumerate(it):
        print(tag, "iter", k, float(st.gamma), float(st.f_x), float(st.g_z), float(st.res_norm_inf))
        if k >= 4:
            break
    if hasattr(comm, "close"):
        comm.close()
