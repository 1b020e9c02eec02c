"""GPU parity of K11 (forward-difference operator and adjoint, csrc/stencil_kernels.cu) and Chambolle-Pock TV denoising built on it.
STAGED FOR ROUND 2: written without GPU access; bars: both kernels bit-exact vs oracle/stencil_oracle.py, adjoint identity to rounding,
Chambolle-Pock TV minimiser equal to the Douglas-Rachford TV splitting's."""
import ctypes as C

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import panoc_oracle as po  # noqa: E402
from oracle import tv_oracle as tvo  # noqa: E402
from oracle.stencil_oracle import FiniteDifference2D as FDo  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import ptr  # noqa: E402

from gpu_util import ctx, dev, dt  # noqa: E402


@pytest.mark.parametrize("T", [np.float32, np.float64])
@pytest.mark.parametrize("H,W", [(1, 1), (1, 9), (7, 1), (5, 6), (33, 64), (130, 257), (512, 1000)])
def test_fd2d_bit_exact_and_adjoint(T, H, W):
    rng = np.random.default_rng(H * 7 + W)
    u, pq = rng.standard_normal(H * W).astype(T), rng.standard_normal(2 * H * W).astype(T)
    c = ctx()
    out_f = torch.empty(2 * H * W, dtype=dev(u).dtype, device="cuda")
    out_a = torch.empty(H * W, dtype=dev(u).dtype, device="cuda")
    L.check(c.lib.pb_fd2d_forward(c.h, dt(T), H, W, ptr(dev(u)), ptr(out_f)))
    L.check(c.lib.pb_fd2d_adjoint(c.h, dt(T), H, W, ptr(dev(pq)), ptr(out_a)))
    assert np.array_equal(out_f.cpu().numpy(), FDo(H, W).mul(u))
    assert np.array_equal(out_a.cpu().numpy(), FDo(H, W).mul_t(pq))
    lhs = float(out_f.cpu().numpy().astype(np.float64) @ pq.astype(np.float64))
    rhs = float(u.astype(np.float64) @ out_a.cpu().numpy().astype(np.float64))
    assert abs(lhs - rhs) <= 64 * np.finfo(T).eps * max(1.0, np.sqrt(H * W)) * max(1.0, abs(lhs))
    op = pa.FiniteDifference2D(H, W)
    assert torch.equal(op.mul_into(torch.empty_like(out_f), dev(u)), out_f) and torch.equal(op.mul_t_into(torch.empty_like(out_a), dev(pq)), out_a)


def test_chambolle_pock_tv_matches_the_douglas_rachford_splitting():
    T = np.float64
    rng = np.random.default_rng(4)
    H, W = 24, 32
    img = np.zeros((H, W))
    img[5:15, 6:20] = 1.0
    b = (img + 0.1 * rng.standard_normal((H, W))).astype(T)
    lam = 0.2
    (x, y), it = pa.ChambollePock(tol=1e-8, maxit=50000)(x0=np.zeros(H * W, T), y0=np.zeros(2 * H * W, T), g=pa.SqrNormL2(1.0, b.reshape(-1)),
                                                        h=pa.NormL1(lam), L=pa.FiniteDifference2D(H, W))
    f = pa.TVSplit(b, lam)
    ydr, k = pa.DouglasRachford(tol=1e-8, maxit=50000)(x0=f.initial_point(), f=f, g=pa.IndConsensus(5), gamma=1.0)
    u = f.image(ydr).cpu().numpy().reshape(-1)
    assert it < 50000 and k < 50000
    fo = tvo.TVSplit(b, lam, (H, W))
    assert abs(fo.objective(x) - fo.objective(u)) <= 1e-6 * fo.objective(u) and np.max(np.abs(x - u)) <= 1e-4
    _ = po, C
