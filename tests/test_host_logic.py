"""CPU tests of the host-side scalar logic (no kernels run here): Nesterov recurrences in strict R arithmetic,
double-double shard combination, shard partitioning, f_model from reductions."""
import itertools
import math

import numpy as np
import pytest

import proxb200 as pa
from oracle import fb_oracle as o
from proxb200 import _lib as L
from proxb200.algorithms import f_model
from proxb200.host import Scalars, dd_add, shard_bounds, two_sum

TYPES = [np.float64, np.float32]


@pytest.mark.parametrize("R", TYPES)
def test_sequences_match_oracle_bitwise(R):
    a = list(itertools.islice(iter(pa.FixedNesterovSequence(R)), 200))
    b = list(itertools.islice(o.fixed_nesterov_sequence(R), 200))
    assert a == b and all(type(v) is R for v in a)
    a = list(itertools.islice(iter(pa.SimpleNesterovSequence(R)), 50))
    b = list(itertools.islice(o.simple_nesterov_sequence(R), 50))
    assert a == b
    assert next(pa.ConstantNesterovSequence(R(1), R(0.1))) == next(o.constant_nesterov_sequence(R(1), R(0.1)))
    for m in (R(0), R(0.3)):
        s1, s2 = pa.AdaptiveNesterovSequence(m), o.AdaptiveNesterovSequence(m)
        gam = R(0.9)
        for k in range(100):
            if k % 17 == 16:
                gam = R(gam * R(0.5))
            pk = s1.peek(gam)
            v1, v2 = s1.next(gam), s2.next(gam)
            assert v1 == v2 == pk and type(v1) is R


@pytest.mark.parametrize("R", TYPES)
def test_reference_nesterov_identities(R):
    # test/accel/test_nesterov.jl:63-81 on the product's host implementation
    seq, fixed = pa.AdaptiveNesterovSequence(R(0)), iter(pa.FixedNesterovSequence(R))
    for _ in range(20):
        assert np.isclose(seq.next(R(1.7)), next(fixed), rtol=float(np.sqrt(np.finfo(R).eps)))
    seq, const = pa.AdaptiveNesterovSequence(R(1)), pa.ConstantNesterovSequence(R(1), R(0.5))
    for _ in range(20):
        assert np.isclose(seq.next(R(0.5)), next(const), rtol=float(np.sqrt(np.finfo(R).eps)))


def test_double_double_combination_is_partition_independent():
    rng = np.random.default_rng(0)
    v = (rng.standard_normal(20000) * 10.0 ** rng.integers(-8, 8, 20000))
    exact = math.fsum(v)

    def shard_pair(chunk):
        acc = (0.0, 0.0)
        for e in chunk.tolist():
            s, err = two_sum(acc[0], e)
            acc = (s, acc[1] + err)
        return acc

    for P in (1, 2, 3, 8):
        parts = np.zeros((P, L.PB_NSCALARS))
        for r, (lo, hi) in enumerate(shard_bounds(v.size, P)):
            parts[r, L.PB_S_RESSQ], parts[r, L.PB_S_RESSQ + 1] = shard_pair(v[lo:hi])
        assert Scalars(parts).res_sq == exact
    # max slots: NaN propagates, like Julia's norm(., Inf)
    parts = np.zeros((2, L.PB_NSCALARS))
    parts[0, L.PB_S_RESINF], parts[1, L.PB_S_RESINF] = 3.0, float("nan")
    assert math.isnan(Scalars(parts).res_inf)
    parts[1, L.PB_S_RESINF] = 5.0
    assert Scalars(parts).res_inf == 5.0
    assert dd_add((1.0, 1e-20), (2.0, -1e-20))[0] == 3.0


def test_shard_bounds():
    for n, P in [(10**8, 8), (1000, 3), (5, 8), (0, 2), (128 * 7, 4)]:
        b = shard_bounds(n, P, align=32)
        assert b[0][0] == 0 and b[-1][1] == n and len(b) == P
        for (lo, hi), (lo2, _) in zip(b, b[1:]):
            assert hi == lo2 and lo <= hi
        for lo, hi in b[:-1]:
            assert (lo % 32 == 0 or lo == n) and (hi % 32 == 0 or hi == n)
    b = shard_bounds(128 * 10, 4, align=128)      # NormL21: whole groups per shard
    assert all(lo % 128 == 0 for lo, _ in b)
    with pytest.raises(ValueError):
        shard_bounds(10, 0)


@pytest.mark.parametrize("R", TYPES)
def test_f_model_from_reductions_matches_oracle(R):
    rng = np.random.default_rng(3)
    g = rng.standard_normal(257).astype(R)
    res = rng.standard_normal(257).astype(R)
    f_x, Lc = R(3.25), R(7.5)
    want = o.f_model(f_x, g, res, Lc)
    gdr = math.fsum((g.astype(np.float64) * res.astype(np.float64)).tolist())
    rsq = math.fsum((res.astype(np.float64) ** 2).tolist())
    got = f_model(R, f_x, gdr, rsq, Lc)
    assert type(got) is R
    assert abs(float(got) - float(want)) <= 4 * np.finfo(R).eps * abs(float(want))


def test_iteration_parameter_defaults_follow_reference():
    x0 = np.zeros(4, np.float32)
    it = pa.FastForwardBackwardIteration(x0=x0)
    assert it.adaptive and it.gamma is None and type(it.minimum_gamma) is np.float32          # :50-52
    assert float(it.reduce_gamma) == 0.5 and float(it.increase_gamma) == 1.0 and float(it.mf) == 0.0
    it = pa.ForwardBackwardIteration(x0=x0, Lf=4.0)
    assert not it.adaptive and it.gamma == 0.25
    it = pa.ForwardBackwardIteration(x0=x0, Lf=4.0, adaptive=True)
    assert it.adaptive
    assert pa.ProximalGradient is pa.ForwardBackward and pa.FastProximalGradient is pa.FastForwardBackward
    alg = pa.FastForwardBackward(maxit=7, tol=1e-3, Lf=2.0)
    assert alg.maxit == 7 and alg.kwargs == {"Lf": 2.0} and alg.iterator_type is pa.FastForwardBackwardIteration
    with pytest.raises(TypeError):
        pa.ForwardBackwardIteration(x0=np.zeros(3, np.int32))


def test_tv_engine_connect_glue(monkeypatch):
    """TVDouglasRachfordEngine._connect (tv.py) end to end with stand-ins for the process group and the cudaIpc calls: the
    neighbour's two ping-pong buffers are opened and the halo addresses are base + offset of halo_plan."""
    import ctypes as C
    import types

    from proxb200 import tv

    W, es = 8, 4
    infos = [(0, 0, 5, W, b"h00", b"h01"), (1, 5, 8, W, b"h10", b"h11")]
    bases = {b"h00": 0x1000000, b"h01": 0x2000000, b"h10": 0x3000000, b"h11": 0x4000000}
    opened, closed, barriers = [], [], []

    class FakeLib:
        def pb_ipc_open(self, h, handle, pp):
            opened.append(handle)
            pp._obj.value = bases[handle]
            return 0

        def pb_ipc_close(self, h, p):
            closed.append(p.value)
            return 0

    class FakeDist:
        def all_gather_object(self, out, mine, group=None):
            assert mine[:4] == infos[1][:4]
            out[:] = infos

        def barrier(self, group=None):
            barriers.append(1)

    class FakeBuf:
        def __init__(self, h):
            self.h = h

        def handle(self):
            return self.h

    monkeypatch.setattr(tv, "torch", lambda: types.SimpleNamespace(cuda=types.SimpleNamespace(synchronize=lambda dev: None)))
    eng = object.__new__(tv.TVDouglasRachfordEngine)
    eng.f = types.SimpleNamespace(comm=types.SimpleNamespace(dist=FakeDist(), rank=1, size=2), R=np.float32, H=8, W=W, row0=5, Hglob=13)
    eng.ctx = types.SimpleNamespace(lib=FakeLib(), h=C.c_void_p(1), device="cuda:1")
    eng.bufs = [FakeBuf(b"h10"), FakeBuf(b"h11")]
    eng._opened = []
    eng._connect()
    assert opened == [b"h00", b"h01"] and barriers == [1]
    off = (3 * 5 * W + 4 * W) * es                      # last row of rank 0's copy 3 (the pair (4, 5) is an even pair)
    assert eng.halo == [(0x1000000 + off, None), (0x2000000 + off, None)]
    eng.bufs = None
    eng.close()
    assert closed == [0x1000000, 0x2000000]
