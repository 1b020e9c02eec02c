"""CPU checks of the drop-in boundary: the shared library builds for sm_100a without a GPU, loads, and exports every
symbol that include/proxb200.h declares; the ctypes table mirrors the header; the product fails loudly (no CPU path)."""
import ctypes
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import ROOT

HEADER = os.path.join(ROOT, "include", "proxb200.h")


def header_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(pb_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_path():
    names = header_functions()
    for required in ["pb_fb_step", "pb_ffb_step", "pb_prox_apply", "pb_lsq_dense_residual", "pb_lsq_dense_gradient",
                     "pb_lsq_blockdiag_residual", "pb_lsq_blockdiag_gradient", "pb_read_scalars", "pb_ctx_create",
                     "pb_ffb_step_host", "pb_last_error"]:
        assert required in names


def test_library_exports_every_declared_symbol(lib_built):
    lib = ctypes.CDLL(lib_built)
    for name in header_functions():
        assert hasattr(lib, name), f"{name} declared in proxb200.h but not exported by {lib_built}"


def test_ctypes_table_matches_header(lib_built):
    from proxb200 import _lib

    assert sorted(_lib.SIGNATURES) == header_functions()
    lib = _lib.load(lib_built)
    assert lib.pb_version().decode().startswith("0.")
    # argument counts agree with the header prototypes
    src = re.sub(r"/\*.*?\*/", "", open(HEADER).read(), flags=re.S)
    for name, (_, args) in _lib.SIGNATURES.items():
        m = re.search(r"\b" + name + r"\s*\(([^;]*?)\)\s*;", src, flags=re.S)
        assert m, name
        params = m.group(1).strip()
        n = 0 if params in ("void", "") else params.count(",") + 1
        assert n == len(args), (name, n, len(args))


def test_built_for_sm_100a(lib_built):
    out = subprocess.run(["cuobjdump", "--list-elf", lib_built], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    assert "sm_100a" in out.stdout


def test_error_reporting_without_gpu(lib_built):
    """Argument errors are reported through status codes + pb_last_error, never by crashing; with no device the context
    cannot be created and the Python package refuses to run (there is no CPU fallback)."""
    import torch

    from proxb200 import _lib
    import proxb200 as pa

    lib = _lib.load(lib_built)
    h = ctypes.c_void_p()
    if not torch.cuda.is_available():
        rc = lib.pb_ctx_create(0, None, 0, ctypes.byref(h))
        assert rc != 0 and lib.pb_last_error()
        with pytest.raises(pa.ProxB200Error):
            pa.LeastSquares(np.eye(3), np.zeros(3))
        with pytest.raises(pa.ProxB200Error):
            pa.FastForwardBackward()(x0=np.zeros(3), f=pa.Zero(), g=pa.NormL1(1.0))
    assert lib.pb_fb_step(None, 0, 4, None, None, 0.1, None, None, None, None) == 1   # PB_EINVAL: null context
    assert b"null context" in lib.pb_last_error()
    assert lib.pb_ctx_sync(None) == 1


def test_missing_library_is_loud(tmp_path):
    from proxb200 import _lib

    with pytest.raises(_lib.ProxB200Error, match="no fallback"):
        _lib.load(str(tmp_path / "nope.so"))


def test_dense_shard_bounds_follow_the_chunk_rule(lib_built):
    """Column shards of a dense A (C2) must start on the column chunks of the residual order (csrc/lsq_order.h: ceil(n/64) columns,
    clamped to [32, 4096], rounded up to a multiple of 4); pure host code, no device needed."""
    from proxb200 import _lib
    from proxb200.host import dense_shard_bounds

    lib = _lib.load(lib_built)
    for n in (1, 31, 100, 3000, 100_000, 262_144, 1_000_003):
        cc_ref = min(max((n + 63) // 64, 32), 4096)
        cc_ref = (cc_ref + 3) & ~3
        for dt in (_lib.PB_F32, _lib.PB_F64):
            assert lib.pb_lsq_dense_chunk_cols(dt, 500, n) == cc_ref
        for P in (1, 2, 3, 8):
            b = dense_shard_bounds(np.float32, 500, n, P)
            assert len(b) == P and b[0][0] == 0 and b[-1][1] == n
            for (lo, hi), (lo2, _) in zip(b, b[1:] + [(n, n)]):
                assert lo <= hi == lo2 and (lo % cc_ref == 0 or lo == n) and (hi % cc_ref == 0 or hi == n)
            sizes = [(hi - lo + cc_ref - 1) // cc_ref for lo, hi in b]
            assert max(sizes) - min(sizes) <= 1                      # chunks dealt evenly
