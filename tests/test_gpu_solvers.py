"""GPU parity, solver level: the product's ForwardBackward / FastForwardBackward (fused CUDA path through the C ABI)
against the CPU oracle on the reference's own problems.

Bars: fp64 -- IDENTICAL iteration count, converged iterate within 1e-9*max(1,||z||inf), relative objective gap <= 1e-10;
fp32 -- iterate within 1e-4 (the reference's own Float32 TOL, test/problems/test_lasso_small.jl:44), iteration count
within +-1 % (+-2)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

torch = pytest.importorskip("torch")
if not torch.cuda.is_available():
    pytest.skip("no CUDA device", allow_module_level=True)

import proxb200 as pa  # noqa: E402
from oracle import fb_oracle as o  # noqa: E402

from conftest import load_golden  # noqa: E402

TYPES = [np.float64, np.float32]


def _objective(A, b, lam, v):
    v = np.asarray(v, dtype=np.float64)
    r = A.astype(np.float64) @ v - b.astype(np.float64)
    return 0.5 * r @ r + float(lam) * np.abs(v).sum()


def _unit_problem(T):
    d = load_golden("unit_lasso_4x5")
    A = np.asfortranarray(d["A"].astype(T))
    b = d["b"].astype(T)
    lam = T(0.1) * o.norm_inf(A.T @ b)
    Lf = T(np.linalg.norm(d["A"], 2)) ** 2
    return A, b, lam, Lf, d["xstar"].astype(T)


CASES_UNIT = [
    ("fb_fixed", "fb", dict(use_Lf=True), 150),
    ("fb_adaptive", "fb", dict(adaptive=True), 300),
    ("fb_regret", "fb", dict(adaptive=True, increase_gamma=1.01), 150),
    ("ffb_fixed", "ffb", dict(use_Lf=True), 100),
    ("ffb_adaptive", "ffb", dict(adaptive=True), 200),
    ("ffb_regret", "ffb", dict(adaptive=True, increase_gamma=1.01), 100),
    ("ffb_custom", "ffb", dict(use_Lf=True, custom=True), 100),
]


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("name,alg,kw,bound", CASES_UNIT)
def test_lasso_small_like_the_reference(T, name, alg, kw, bound):
    """test/problems/test_lasso_small.jl:46-135 run on the GPU path, plus equality with the oracle's iteration count."""
    A, b, lam, Lf, xstar = _unit_problem(T)
    TOL = T(1e-4)
    kw_o, kw_g = {}, {}
    if kw.get("use_Lf"):
        kw_o["Lf"] = kw_g["Lf"] = Lf
    if kw.get("adaptive"):
        kw_o["adaptive"] = kw_g["adaptive"] = True
    if "increase_gamma" in kw:
        kw_o["increase_gamma"] = kw_g["increase_gamma"] = T(kw["increase_gamma"])
    if kw.get("custom"):
        kw_o["extrapolation_sequence"] = o.fixed_nesterov_sequence(T)
        kw_g["extrapolation_sequence"] = pa.FixedNesterovSequence(T)
    x0 = np.zeros(5, T)
    solver_o = o.fast_forward_backward if alg == "ffb" else o.forward_backward
    z_o, it_o = solver_o(x0, o.LeastSquares(A, b), o.NormL1(lam), tol=TOL, **kw_o)
    solver = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=TOL)
    x, it = solver(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), **kw_g)
    assert isinstance(x, np.ndarray) and x.dtype == T
    assert np.max(np.abs(x - xstar)) <= TOL
    assert it < bound
    assert np.all(x0 == 0)
    if T == np.float64:
        assert it == it_o
        assert np.max(np.abs(x - z_o)) <= 1e-9
    else:
        assert abs(it - it_o) <= max(2, it_o // 100)
        assert np.max(np.abs(x - z_o)) <= 1e-4


@pytest.mark.parametrize("T", TYPES)
def test_lasso_strongly_convex_like_the_reference(T):
    """test/problems/test_lasso_small_strongly_convex.jl:65-144."""
    d = load_golden("unit_lasso_sc_5x5")
    A = np.asfortranarray(d["A"].astype(T))
    b, xstar, x0 = d["b"].astype(T), d["xstar"].astype(T), d["x0"].astype(T)
    lam, mf, Lf = T(d["lam"]), T(d["mf"]), T(d["Lf"])
    TOL = T(1e-4)
    x0_backup = x0.copy()
    f, g = pa.LeastSquares(A, b), pa.NormL1(lam)
    runs = [
        (pa.ForwardBackward(tol=TOL), dict(Lf=Lf), 110),
        (pa.ForwardBackward(tol=TOL, adaptive=True), {}, 300),
        (pa.ForwardBackward(tol=TOL, adaptive=True, increase_gamma=T(1.01)), {}, 80),
        (pa.FastForwardBackward(tol=TOL), dict(Lf=Lf, mf=mf), 35),
        (pa.FastForwardBackward(tol=TOL, adaptive=True), {}, 100),
        (pa.FastForwardBackward(tol=TOL, adaptive=True, increase_gamma=T(1.01)), {}, 100),
        (pa.FastForwardBackward(tol=TOL), dict(gamma=T(1) / Lf, mf=mf,
                                              extrapolation_sequence=pa.ConstantNesterovSequence(mf, T(1) / Lf)), 35),
    ]
    for solver, kw, bound in runs:
        y, it = solver(x0=x0, f=f, g=g, **kw)
        assert y.dtype == T and np.max(np.abs(y - xstar)) <= TOL and it < bound
        assert np.array_equal(x0, x0_backup)


@pytest.mark.parametrize("name", ["tiny", "small", "medium"])
@pytest.mark.parametrize("alg", ["ffb", "fb"])
def test_benchmark_fixtures_iteration_count_and_objective(name, alg):
    """benchmark/benchmarks.jl:47-61 (fp64, tol 1e-6, x0 = 0, adaptive): identical iteration count to the oracle,
    objective gap <= 1e-10 vs the oracle and <= 1e-6 vs the fixture's xstar."""
    d = load_golden("lasso_" + name)
    A, b, lam, xstar = d["A"], d["b"], d["lam"], d["xstar"]
    n = A.shape[1]
    solver_o = o.fast_forward_backward if alg == "ffb" else o.forward_backward
    z_o, it_o = solver_o(np.zeros(n), o.LeastSquares(A, b), o.NormL1(lam), tol=1e-6)
    solver = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6)
    z, it = solver(x0=np.zeros(n), f=pa.LeastSquares(A, b), g=pa.NormL1(lam))
    assert it == it_o
    assert np.max(np.abs(z - z_o)) <= 1e-9 * max(1.0, np.max(np.abs(z_o)))
    obj, obj_o, obj_s = _objective(A, b, lam, z), _objective(A, b, lam, z_o), _objective(A, b, lam, xstar)
    assert abs(obj - obj_o) / obj_o <= 1e-10
    if it < 10000:
        assert abs(obj - obj_s) / obj_s <= 1e-6
        assert np.max(np.abs(z - xstar)) <= 2 * np.max(np.abs(z_o - xstar)) + 1e-12
    last = solver.last_state
    assert float(last.gamma) > 0 and last.z.is_cuda


@pytest.mark.parametrize("T", TYPES)
@pytest.mark.parametrize("alg", ["ffb", "fb"])
@pytest.mark.parametrize("adaptive", [False, True])
def test_state_by_state_equivalence_with_oracle(T, alg, adaptive):
    """The pattern of test/problems/test_equivalence.jl:71-83: step two implementations side by side and compare the
    states.  fp64: z, x, gamma agree to ~1e-13; lazily materialised y and res equal x - gamma*grad and x - z bit for bit."""
    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T), T(d["lam"])
    n = A.shape[1]
    kw = {} if adaptive else dict(Lf=T(np.linalg.norm(d["A"], 2) ** 2))
    It_o = o.FastForwardBackwardIteration if alg == "ffb" else o.ForwardBackwardIteration
    It_g = pa.FastForwardBackwardIteration if alg == "ffb" else pa.ForwardBackwardIteration
    it_o = iter(It_o(np.zeros(n, T), o.LeastSquares(A, b), o.NormL1(lam), **kw))
    it_g = iter(It_g(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=pa.NormL1(lam), **kw))
    tol = 1e-12 if T == np.float64 else 2e-5
    for k in range(25):
        so, sg = next(it_o), next(it_g)
        assert type(sg.gamma) is T and type(sg.f_x) is T and type(sg.g_z) is T
        assert np.isclose(float(sg.gamma), float(so.gamma), rtol=1e-12 if T == np.float64 else 1e-6)
        scale = max(1.0, float(np.max(np.abs(so.z))))
        assert np.max(np.abs(sg.z.cpu().numpy() - so.z)) <= tol * scale
        assert np.max(np.abs(sg.x.cpu().numpy() - so.x)) <= tol * scale
        assert np.isclose(float(sg.f_x), float(so.f_x), rtol=1e-11 if T == np.float64 else 1e-4)
        assert np.isclose(float(sg.g_z), float(so.g_z), rtol=1e-11 if T == np.float64 else 1e-4, atol=1e-30)
        xg, gg, zg = sg.x.cpu().numpy(), sg.grad_f_x.cpu().numpy(), sg.z.cpu().numpy()
        assert np.array_equal(sg.y.cpu().numpy(), xg - sg.gamma * gg)
        assert np.array_equal(sg.res.cpu().numpy(), xg - zg)
        assert float(sg.res_norm_inf) == float(np.max(np.abs(xg - zg)))


@pytest.mark.parametrize("T", TYPES)
def test_box_qp_doc_example_and_user_callbacks(T):
    """docs/src/guide/getting_started.jl:59-72 (analytic answer (2.3/3.4, 0)) with USER-DEFINED f and g plugged in through
    the callback contract (value_and_gradient / prox!), like docs/src/guide/custom_objectives.jl:50-61,115-124."""
    Q = torch.tensor([[3.4, 1.2], [1.2, 4.5]], dtype=torch.float64 if T == np.float64 else torch.float32, device="cuda")
    q = torch.tensor([-2.3, 9.9], dtype=Q.dtype, device="cuda")

    class Quadratic:                       # user smooth term: returns a fresh gradient, like the reference contract
        calls = 0

        def value_and_gradient(self, x):
            Quadratic.calls += 1
            g = Q @ x
            return float(0.5 * (x @ g) + q @ x), g + q

    class MyBox:                           # user proximable term: in-place prox!, returns g(z)
        def prox_(self, z, y, gamma):
            torch.clamp(y, 0.0, 1.0, out=z)
            return 0.0

    ffb = pa.FastForwardBackward(maxit=1000, tol=1e-5)
    sol, it = ffb(x0=np.ones(2, T), f=Quadratic(), g=MyBox())
    assert np.allclose(sol, [2.3 / 3.4, 0.0], atol=1e-4) and it < 1000 and Quadratic.calls > it
    # the same problem with the built-in fused IndBox gives the same iterates
    sol2, it2 = ffb(x0=np.ones(2, T), f=Quadratic(), g=pa.IndBox(0, 1))
    assert it2 == it and np.allclose(sol2, sol, atol=1e-6)
    # torch CUDA x0 in -> CUDA tensor out, x0 untouched
    x0 = torch.ones(2, dtype=Q.dtype, device="cuda")
    sol3, it3 = ffb(x0=x0, f=Quadratic(), g=pa.IndBox(0, 1))
    assert sol3.is_cuda and it3 == it and torch.all(x0 == 1)


@pytest.mark.parametrize("T", TYPES)
def test_ball_and_group_lasso_converge_to_oracle(T):
    rng = np.random.default_rng(0)
    m, n = 60, 256
    A = np.asfortranarray((rng.standard_normal((m, n)) / np.sqrt(m)).astype(T))
    b = rng.standard_normal(m).astype(T)
    Lf = T(np.linalg.norm(A.astype(np.float64), 2) ** 2)
    tol = T(1e-5 if T == np.float64 else 1e-4)
    for g_o, g_g in [(o.IndBallL2(T(0.5)), pa.IndBallL2(T(0.5))), (o.NormL21(T(0.3), 16), pa.NormL21(T(0.3), 16))]:
        z_o, it_o = o.fast_forward_backward(np.zeros(n, T), o.LeastSquares(A, b), g_o, tol=tol, Lf=Lf, maxit=3000)
        z, it = pa.FastForwardBackward(tol=tol, maxit=3000)(x0=np.zeros(n, T), f=pa.LeastSquares(A, b), g=g_g, Lf=Lf)
        assert abs(it - it_o) <= max(2, it_o // 50)
        assert np.max(np.abs(z - z_o)) <= (1e-7 if T == np.float64 else 2e-3)


def test_verbose_display_and_maxit(capsys):
    """test/problems/test_verbose.jl: verbose runs print `it | gamma | residual` rows; maxit caps the loop."""
    d = load_golden("lasso_tiny")
    solver = pa.ForwardBackward(tol=1e-6, maxit=50, verbose=True, freq=10)
    z, it = solver(x0=np.zeros(10), f=pa.LeastSquares(d["A"], d["b"]), g=pa.NormL1(d["lam"]))
    out = capsys.readouterr().out.strip().splitlines()
    assert it == 50 and len(out) == 5 and out[0].split("|")[0].strip() == "10"


def _m1_problem():
    """SURVEY.md section 8d M1 = BASELINE.json configs[0] as written: synthetic 200 x 500 fp64 Lasso, seed 1."""
    rng = np.random.default_rng(1)
    A = np.asfortranarray(rng.standard_normal((200, 500)))
    xt = np.zeros(500)
    idx = rng.choice(500, 25, replace=False)
    xt[idx] = rng.standard_normal(25)
    b = A @ xt + 0.01 * rng.standard_normal(200)
    lam = 0.1 * np.max(np.abs(A.T @ b))
    return A, b, lam


@pytest.mark.parametrize("alg", ["ffb", "fb"])
@pytest.mark.parametrize("mode", ["adaptive", "fixed"])
def test_config1_synthetic_200x500(alg, mode):
    A, b, lam = _m1_problem()
    kw = {} if mode == "adaptive" else dict(Lf=np.linalg.norm(A, 2) ** 2)
    so = o.fast_forward_backward if alg == "ffb" else o.forward_backward
    z_o, it_o = so(np.zeros(500), o.LeastSquares(A, b), o.NormL1(lam), tol=1e-6, **kw)
    sg = (pa.FastForwardBackward if alg == "ffb" else pa.ForwardBackward)(tol=1e-6)
    z, it = sg(x0=np.zeros(500), f=pa.LeastSquares(A, b), g=pa.NormL1(lam), **kw)
    assert it == it_o and it < 10000
    assert np.max(np.abs(z - z_o)) <= 1e-9 * max(1.0, np.max(np.abs(z_o)))
    assert abs(_objective(A, b, lam, z) - _objective(A, b, lam, z_o)) / _objective(A, b, lam, z_o) <= 1e-10


@pytest.mark.parametrize("T", TYPES)
def test_config2_blockdiag_lasso_fista(T):
    """BASELINE.json configs[1] structure (block-diagonal implicit A, FISTA + NormL1, fixed gamma = 1/L from 20 power
    iterations) at a size the oracle finishes in seconds."""
    rng = np.random.default_rng(2)
    nblk, mb, nb = 6, 20, 400
    blocks = (rng.standard_normal((nblk, mb, nb)) / np.sqrt(mb)).astype(T)
    xt = np.zeros(nblk * nb)
    idx = rng.choice(nblk * nb, 24, replace=False)
    xt[idx] = rng.standard_normal(24)
    b = (np.einsum("bij,bj->bi", blocks.astype(np.float64), xt.reshape(nblk, nb)).reshape(-1) + 0.01 * rng.standard_normal(nblk * mb)).astype(T)
    lam = T(0.1 * np.max(np.abs(np.einsum("bij,bi->bj", blocks.astype(np.float64), b.reshape(nblk, mb).astype(np.float64)))))
    v = rng.standard_normal(nblk * nb)
    for _ in range(20):
        w = np.einsum("bij,bi->bj", blocks.astype(np.float64), np.einsum("bij,bj->bi", blocks.astype(np.float64), v.reshape(nblk, nb))).reshape(-1)
        Lhat = np.linalg.norm(w) / np.linalg.norm(v)
        v = w / np.linalg.norm(w)
    Lf = T(1.05 * Lhat)
    tol = T(1e-6 if T == np.float64 else 1e-4)
    z_o, it_o = o.fast_forward_backward(np.zeros(nblk * nb, T), o.BlockDiagLeastSquares(blocks, b), o.NormL1(lam), tol=tol, Lf=Lf)
    z, it = pa.FastForwardBackward(tol=tol)(x0=np.zeros(nblk * nb, T), f=pa.BlockDiagLeastSquares.from_numpy(blocks, b), g=pa.NormL1(lam), Lf=Lf)
    if T == np.float64:
        assert it == it_o and np.max(np.abs(z - z_o)) <= 1e-9
    else:
        assert abs(it - it_o) <= max(2, it_o // 100) and np.max(np.abs(z - z_o)) <= 1e-4


@pytest.mark.parametrize("T", TYPES)
def test_native_driver_equals_python_host(T):
    """The in-library driver loop (pb_solve, csrc/solve.cu) and the Python loop issue the same kernels and the same scalar
    arithmetic in R: identical iteration counts and bit-identical solutions, for every algorithm variant."""
    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T), T(d["lam"])
    Lf = T(np.linalg.norm(d["A"], 2) ** 2)
    tol = T(1e-6 if T == np.float64 else 1e-4)
    x0 = np.zeros(A.shape[1], T)
    cases = [
        (pa.FastForwardBackward, {}),
        (pa.FastForwardBackward, dict(increase_gamma=T(1.01))),
        (pa.FastForwardBackward, dict(Lf=Lf)),
        (pa.FastForwardBackward, dict(Lf=Lf, mf=T(0.01))),
        (pa.FastForwardBackward, dict(Lf=Lf, extrapolation_sequence=pa.FixedNesterovSequence(T))),
        (pa.FastForwardBackward, dict(Lf=Lf, extrapolation_sequence=pa.SimpleNesterovSequence(T))),
        (pa.FastForwardBackward, dict(Lf=Lf, extrapolation_sequence=pa.ConstantNesterovSequence(T(0.05), T(1) / Lf))),
        (pa.ForwardBackward, {}),
        (pa.ForwardBackward, dict(increase_gamma=T(1.01))),
        (pa.ForwardBackward, dict(Lf=Lf)),
    ]
    for mk, kw in cases:
        for g in (pa.NormL1(lam), pa.IndBox(T(-0.05), T(0.05)), pa.NormL21(lam, 4), pa.Zero()):
            f = pa.LeastSquares(A, b)
            sol_p, sol_n = mk(tol=tol, maxit=600, driver="python"), mk(tol=tol, maxit=600, driver="native")
            zp, itp = sol_p(x0=x0, f=f, g=g, **kw)
            zn, itn = sol_n(x0=x0, f=f, g=g, **kw)
            assert sol_p.last_driver == "python" and sol_n.last_driver == "native"
            assert itn == itp, (mk.__name__, kw.keys(), type(g).__name__, itn, itp)
            assert np.array_equal(zn, zp, equal_nan=True)
            sn, sp = sol_n.last_state, sol_p.last_state
            assert sn.gamma == sp.gamma and sn.f_x == sp.f_x and sn.g_z == sp.g_z and float(sn.res_norm_inf) == float(sp.res_norm_inf)
            assert torch.equal(sn.x, sp.x) and torch.equal(sn.grad_f_x, sp.grad_f_x)
    # other built-in smooth terms
    rng = np.random.default_rng(0)
    bvec = rng.standard_normal(5000).astype(T)
    for f in (pa.SquaredDistance(torch.as_tensor(bvec).cuda()),
              pa.BlockDiagLeastSquares.from_numpy((rng.standard_normal((5, 8, 1000)) / 3).astype(T), rng.standard_normal(40).astype(T))):
        for mk, kw in ((pa.FastForwardBackward, {}), (pa.ForwardBackward, {}), (pa.FastForwardBackward, dict(gamma=T(0.005)))):
            zp, itp = mk(tol=tol, maxit=300, driver="python")(x0=np.zeros(5000, T), f=f, g=pa.NormL1(T(0.3)), **kw)
            zn, itn = mk(tol=tol, maxit=300, driver="native")(x0=np.zeros(5000, T), f=f, g=pa.NormL1(T(0.3)), **kw)
            assert itn == itp and np.array_equal(zn, zp, equal_nan=True) and np.all(np.isfinite(zn))
    # IndBallL2 on one GPU runs natively too (PB_PROX_BALL: norm pass + scale factor formed on the device, no host round trip between
    # the two phases) -- and the Python loop uses the same form: identical results
    for mk, kw in ((pa.FastForwardBackward, dict(Lf=Lf)), (pa.FastForwardBackward, {}), (pa.ForwardBackward, {}), (pa.ForwardBackward, dict(Lf=Lf))):
        sp, sn = mk(tol=tol, maxit=300, driver="python"), mk(tol=tol, maxit=300)
        zp, itp = sp(x0=x0, f=pa.LeastSquares(A, b), g=pa.IndBallL2(T(0.5)), **kw)
        zn, itn = sn(x0=x0, f=pa.LeastSquares(A, b), g=pa.IndBallL2(T(0.5)), **kw)
        assert sn.last_driver == "native" and itn == itp and np.array_equal(zn, zp, equal_nan=True)
        assert np.linalg.norm(zn.astype(np.float64)) <= 0.5 * (1 + 1e-5)
    # ... and equals the two-phase form with the scale factor computed on the host (what row shards use): pb_forward + PB_PROX_SCALE
    import ctypes as C

    from proxb200 import _lib as L
    from proxb200.host import Context, ptr

    ctx = Context.get()
    rng_ = np.random.default_rng(3)
    nv = 100_003
    xv, gv = (torch.as_tensor(rng_.standard_normal(nv).astype(T)).cuda() for _ in range(2))
    dt_ = L.PB_F32 if T == np.float32 else L.PB_F64
    for radius in (0.5, 1e6):
        ball = pa.IndBallL2(T(radius))
        z1, z2, ysc = torch.empty_like(xv), torch.empty_like(xv), torch.empty_like(xv)
        d1 = ball.ball_descriptor(T)
        L.check(ctx.lib.pb_fb_step(ctx.h, dt_, nv, ptr(xv), ptr(gv), 0.3, C.byref(d1), None, ptr(z1), None))
        row1 = ctx.read_scalars()
        L.check(ctx.lib.pb_forward(ctx.h, dt_, nv, ptr(xv), ptr(gv), 0.3, ptr(ysc)))
        ysq = ctx.read_scalars()
        d2 = ball.scale_descriptor(T, ysq[L.PB_S_AUX] + ysq[L.PB_S_AUX + 1])
        L.check(ctx.lib.pb_fb_step(ctx.h, dt_, nv, ptr(xv), ptr(gv), 0.3, C.byref(d2), None, ptr(z2), None))
        row2 = ctx.read_scalars()
        assert torch.equal(z1, z2) and row1[L.PB_S_RESSQ] == row2[L.PB_S_RESSQ] and row1[L.PB_S_RESINF] == row2[L.PB_S_RESINF]
    # what the native driver cannot run falls back (auto) or refuses (native): a user-defined proximable term
    class MyBall:
        def prox_(self, z, y, gamma):
            nrm = float(torch.linalg.vector_norm(y))
            z.copy_(y if nrm <= 0.5 else y * (0.5 / nrm))
            return 0.0

    auto = pa.FastForwardBackward(tol=tol, maxit=50)
    auto(x0=x0, f=pa.LeastSquares(A, b), g=MyBall(), Lf=Lf)
    assert auto.last_driver == "python"
    with pytest.raises(pa.ProxB200Error):
        pa.FastForwardBackward(tol=tol, driver="native")(x0=x0, f=pa.LeastSquares(A, b), g=MyBall(), Lf=Lf)


@pytest.mark.parametrize("T", TYPES)
def test_pipelined_native_loop_is_invisible(T):
    """pb_solve with spare vectors launches iteration k+1 before the scalars of iteration k reach the host (csrc/solve.cu:
    run_ffb_pipelined).  Same iterations, same bits, same final state as the unpipelined native loop, whether the stop test
    or maxit ends the run."""
    d = load_golden("lasso_small")
    A, b, lam = np.asfortranarray(d["A"].astype(T)), d["b"].astype(T), T(d["lam"])
    Lf = T(np.linalg.norm(d["A"], 2) ** 2)
    x0 = np.zeros(A.shape[1], T)
    rng = np.random.default_rng(1)
    n = 200_003
    c, xs = torch.as_tensor(rng.standard_normal(n).astype(T)).cuda(), rng.standard_normal(n).astype(T)
    problems = [
        (dict(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), Lf=Lf), T(1e-6 if T == np.float64 else 1e-4), 5000),
        (dict(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), Lf=Lf), T(-1), 37),
        (dict(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), Lf=Lf), T(-1), 1),
        (dict(x0=x0, f=pa.LeastSquares(A, b), g=pa.NormL1(lam), Lf=Lf), T(-1), 2),
        (dict(x0=xs, f=pa.LinearFunction(c), g=pa.NormL1(T(1)), gamma=T(0.1), extrapolation_sequence=pa.ConstantNesterovSequence(T(0.05), T(1))), T(-1), 25),
        (dict(x0=xs, f=pa.SquaredDistance(c), g=pa.IndBox(T(-0.5), T(0.5)), gamma=T(0.7)), T(1e-5), 500),
    ]
    for kw, tol, maxit in problems:
        a1, a2 = pa.FastForwardBackward(tol=tol, maxit=maxit, driver="native"), pa.FastForwardBackward(tol=tol, maxit=maxit, driver="native")
        a2.pipeline = False
        z1, k1 = a1(**kw)
        if "extrapolation_sequence" in kw:
            kw = dict(kw, extrapolation_sequence=pa.ConstantNesterovSequence(T(0.05), T(1)))
        z2, k2 = a2(**kw)
        assert k1 == k2 and np.array_equal(z1, z2)
        s1, s2 = a1.last_state, a2.last_state
        assert torch.equal(s1.x, s2.x) and torch.equal(s1.z_prev, s2.z_prev) and torch.equal(s1.grad_f_x, s2.grad_f_x)
        assert s1.f_x == s2.f_x and s1.g_z == s2.g_z and float(s1.res_norm_inf) == float(s2.res_norm_inf)
