import os, sys, torch
sys.path.insert(0, "/root/repo")
os.environ["PROXB200_LSQ_FUSED_PROF"] = "1"
from proxb200 import _lib as L
from proxb200.host import Context, ptr
ctx = Context.get(); lib, h = ctx.lib, ctx.h
nblk, mb, nb = 100, 100, 100_000
A = torch.randn(nblk, nb, mb, device="cuda") * 0.1
b = torch.randn(nblk * mb, device="cuda"); x = torch.randn(nblk * nb, device="cuda")
r, grad = torch.empty_like(b), torch.empty_like(x)
for mode in (2, 2, 3, 1):
    L.check(lib.pb_ctx_set_option(h, L.PB_OPT_LSQ_FUSED, mode))
    L.check(lib.pb_lsq_blockdiag_value_and_gradient(h, L.PB_F32, nblk, mb, nb, ptr(A), ptr(x), ptr(b), ptr(r), ptr(grad)))
    torch.cuda.synchronize()
