"""GPU parity of K4: least-squares value and gradient (dense column-major and block-diagonal) against BLAS on the host."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402
from proxb200 import _lib as L  # noqa: E402
from proxb200.host import ptr  # noqa: E402

import gpu_util as G  # noqa: E402
from conftest import load_golden  # noqa: E402

TYPES = [np.float32, np.float64]


def _tol(T, k):
    return (np.finfo(T).eps * 8 * max(1, np.sqrt(k)))


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("m,n", [(4, 5), (5, 10), (50, 100), (500, 1000), (200, 500), (1, 1), (129, 3), (3, 700), (1000, 33),
                                 (4, 100_000), (16, 5000), (48, 2500)])
def test_dense_value_and_gradient(T, m, n):
    rng = np.random.default_rng(m * 1000 + n)
    A = np.asfortranarray(rng.standard_normal((m, n)).astype(T))
    b = rng.standard_normal(m).astype(T)
    x = rng.standard_normal(n).astype(T)
    f = pa.LeastSquares(A, b)
    val, grad = f.value_and_gradient(G.dev(x))
    A64, b64, x64 = A.astype(np.float64), b.astype(np.float64), x.astype(np.float64)
    r64 = A64 @ x64 - b64
    scale_r = np.abs(A64) @ np.abs(x64) + np.abs(b64)
    assert np.all(np.abs(f.r.cpu().numpy() - r64) <= _tol(T, n) * scale_r)
    assert np.isclose(float(val), 0.5 * r64 @ r64, rtol=50 * _tol(T, n))
    g64 = A64.T @ r64
    scale_g = np.abs(A64.T) @ np.abs(r64)
    assert np.all(np.abs(grad.cpu().numpy() - g64) <= 4 * _tol(T, m + n) * scale_g + 1e-30)
    assert type(val) is T
    # value-only path gives the same value (gradient of the FFB line search is discarded by the reference)
    v2 = f.value_into(f.ctx, G.dev(x)).resolve(f.ctx.read_scalars(), None)
    assert v2 == val


def test_dense_value_matches_oracle_on_fixtures():
    """fp64 fixtures: f(x) and grad agree with the oracle's BLAS evaluation to a few ulp (the line search lives there)."""
    for name in ("tiny", "small", "medium"):
        d = load_golden("lasso_" + name)
        A, b = d["A"], d["b"]
        rng = np.random.default_rng(4)
        x = rng.standard_normal(A.shape[1]) * (rng.random(A.shape[1]) < 0.1)
        fo = o.LeastSquares(A, b)
        v_o, g_o = fo.value_and_gradient(x)
        f = pa.LeastSquares(A, b)
        v, g = f.value_and_gradient(G.dev(x))
        assert abs(float(v) - float(v_o)) <= 16 * np.spacing(float(v_o))
        assert np.allclose(g.cpu().numpy(), g_o, rtol=1e-12, atol=1e-12 * np.abs(g_o).max())


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("nblk,mb,nb", [(1, 4, 5), (7, 100, 1000), (3, 33, 77), (100, 10, 257), (2, 200, 4096),
                                        # short columns -> k_gemv_n_sub (1, 2, 4, 8 lanes per column), ragged chunks
                                        (5, 4, 3001), (3, 8, 12800), (9, 16, 700), (4, 32, 1023), (2, 60, 333), (3, 62, 4100), (1, 2, 9)])
def test_blockdiag_value_and_gradient(T, nblk, mb, nb):
    rng = np.random.default_rng(nblk + mb + nb)
    blocks = rng.standard_normal((nblk, mb, nb)).astype(T)
    b = rng.standard_normal(nblk * mb).astype(T)
    x = rng.standard_normal(nblk * nb).astype(T)
    f = pa.BlockDiagLeastSquares.from_numpy(blocks, b)
    val, grad = f.value_and_gradient(G.dev(x))
    B64 = blocks.astype(np.float64)
    r64 = np.einsum("bij,bj->bi", B64, x.astype(np.float64).reshape(nblk, nb)).reshape(-1) - b
    g64 = np.einsum("bij,bi->bj", B64, r64.reshape(nblk, mb)).reshape(-1)
    sr = np.einsum("bij,bj->bi", np.abs(B64), np.abs(x.astype(np.float64)).reshape(nblk, nb)).reshape(-1) + np.abs(b)
    sg = np.einsum("bij,bi->bj", np.abs(B64), np.abs(r64).reshape(nblk, mb)).reshape(-1)
    assert np.all(np.abs(f.r.cpu().numpy() - r64) <= _tol(T, nb) * sr)
    assert np.all(np.abs(grad.cpu().numpy() - g64) <= 4 * _tol(T, nb + mb) * sg + 1e-30)
    assert np.isclose(float(val), 0.5 * r64 @ r64, rtol=50 * _tol(T, nb))
    # same numbers as the oracle's restatement
    v_o, g_o = o.BlockDiagLeastSquares(blocks, b).value_and_gradient(x)
    assert np.isclose(float(val), float(v_o), rtol=1e-4 if T == np.float32 else 1e-12)


@pytest.mark.parametrize("T", TYPES)
def test_sqdist_and_linear(T):
    n = 70_001
    rng = np.random.default_rng(0)
    x, b = rng.standard_normal(n).astype(T), rng.standard_normal(n).astype(T)
    v, g = pa.SquaredDistance(G.dev(b)).value_and_gradient(G.dev(x))
    v_o, g_o = o.SquaredDistance(b).value_and_gradient(x)
    assert np.array_equal(g.cpu().numpy(), g_o)
    assert abs(float(v) - float(v_o)) <= 4 * np.finfo(T).eps * float(v_o)
    v, g = pa.LinearFunction(G.dev(b)).value_and_gradient(G.dev(x))
    assert np.array_equal(g.cpu().numpy(), b) and np.isclose(float(v), float(b.astype(np.float64) @ x.astype(np.float64)), rtol=1e-5, atol=1e-3)


def test_lsq_argument_errors():
    c = G.ctx()
    x = torch.zeros(8, device="cuda")
    assert c.lib.pb_lsq_dense_residual(c.h, 0, 4, 2, ptr(x), 3, ptr(x), ptr(x), ptr(x)) == 1      # lda < m
    assert c.lib.pb_lsq_dense_gradient(c.h, 0, 4, 2, None, 4, ptr(x), ptr(x)) == 1                 # null A
    assert c.lib.pb_lsq_blockdiag_residual(c.h, 0, -1, 2, 2, ptr(x), ptr(x), ptr(x), ptr(x)) == 1
    assert c.lib.pb_last_error()


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("nblk,mb,nb", [(1, 100, 1000), (1, 64, 333), (3, 128, 2000), (1, 500, 1000), (1, 1024, 70), (2, 200, 5003), (1, 10000, 300), (5, 72, 129)])
def test_row_pack_residual_kernel_is_bit_identical_to_thread_per_row(T, nblk, mb, nb):
    """k_gemv_n_partial_v (16-byte row packs) keeps the summation order of k_gemv_n_partial per row: identical r and ||r||^2."""
    rng = np.random.default_rng(mb * 7 + nb)
    A = torch.as_tensor(rng.standard_normal((nblk, nb, mb)).astype(T)).cuda()
    b = torch.as_tensor(rng.standard_normal(nblk * mb).astype(T)).cuda()
    x = torch.as_tensor(rng.standard_normal(nblk * nb).astype(T)).cuda()
    c = G.ctx()
    out = []
    for mode in (1, 0):
        L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_GEMV_SCALAR, mode))
        r = torch.empty_like(b)
        L.check(c.lib.pb_lsq_blockdiag_residual(c.h, G.dt(T), nblk, mb, nb, ptr(A), ptr(x), ptr(b), ptr(r)))
        row = c.read_scalars()
        out.append((r.clone(), row[L.PB_S_AUX], row[L.PB_S_AUX + 1]))
    L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_GEMV_SCALAR, 0))
    assert torch.equal(out[0][0], out[1][0])
    assert out[0][1:] == out[1][1:]
    r64 = np.einsum("kji,kj->ki", A.cpu().numpy().astype(np.float64), x.cpu().numpy().astype(np.float64).reshape(nblk, nb)).reshape(-1) - b.cpu().numpy()
    assert np.allclose(out[1][0].cpu().numpy(), r64, rtol=0, atol=_tol(T, nb) * 50 * np.max(np.abs(r64)))


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("nblk,mb,nb", [(5, 100, 5003), (7, 64, 2100), (4, 128, 40_000), (9, 200, 3000), (1, 100, 700), (12, 72, 64), (3, 100, 100_000)])
@pytest.mark.parametrize("live", [1, 2, 3])
def test_fused_blockdiag_value_and_gradient_is_bit_identical(T, nblk, mb, nb, live):
    """csrc/lsq_fused.cu (every block read from HBM once, second sweep from L2, dynamic N / T work units through a TMA ring) against the
    residual kernel followed by the gradient kernel: same r, grad and ||r||^2, bit for bit, whatever the throttle depth."""
    if T == np.float64 and mb == 200:
        pytest.skip("row packs of a column must fit one CTA (mb <= 128 in Float64): this shape takes the two-kernel path")
    rng = np.random.default_rng(mb + nb + nblk)
    A = torch.as_tensor(rng.standard_normal((nblk, nb, mb)).astype(T)).cuda()
    b = torch.as_tensor(rng.standard_normal(nblk * mb).astype(T)).cuda()
    x = torch.as_tensor(rng.standard_normal(nblk * nb).astype(T)).cuda()
    c = G.ctx()
    out = []
    for mode in (-1, live):
        L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_LSQ_FUSED, mode))
        r, grad = torch.full_like(b, float("nan")), torch.full_like(x, float("nan"))
        l0 = c.launches()
        L.check(c.lib.pb_lsq_blockdiag_value_and_gradient(c.h, G.dt(T), nblk, mb, nb, ptr(A), ptr(x), ptr(b), ptr(r), ptr(grad)))
        row = c.read_scalars()
        out.append((r.clone(), grad.clone(), row[L.PB_S_AUX], row[L.PB_S_AUX + 1], c.launches() - l0))
    L.check(c.lib.pb_ctx_set_option(c.h, L.PB_OPT_LSQ_FUSED, 0))
    assert out[1][4] == 1 and out[0][4] >= 2, "the fused form is one launch"
    assert torch.equal(out[0][0], out[1][0]) and torch.equal(out[0][1], out[1][1])
    assert out[0][2] + out[0][3] == out[1][2] + out[1][3]
