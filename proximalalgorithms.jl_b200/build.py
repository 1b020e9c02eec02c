"""Build libproxb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python proximalalgorithms.jl_b200/build.py [--force] [--verbose]

Objects and the shared library land in proximalalgorithms.jl_b200/lib/ (git-ignored, shipped to the GPU box by gpurun).
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
INCLUDE = os.path.join(ROOT, "include")
LIBNAME = "libproxb200.so"
SOURCES = ["ctx.cu", "step_kernels.cu", "step_tma.cu", "lsq_kernels.cu", "xchg.cu", "solve.cu", "qn_kernels.cu", "dr_kernels.cu", "lsq_prox.cu", "tv_kernels.cu", "stencil_kernels.cu", "panoc_solve.cu", "util_kernels.cu", "persist.cu", "step_multi.cu", "lsq_fused.cu", "lsq_fista.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-ffp-contract=off",
    "-Xptxas", "-v",
    "-I", INCLUDE, "-I", CSRC,
]


def _nvcc():
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libproxb200 cannot be built")


def _stamp():
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)) + ["../../include/proxb200.h"]:
        path = os.path.join(CSRC, name)
        if os.path.isfile(path):
            h.update(name.encode())
            h.update(open(path, "rb").read())
    h.update(" ".join(NVCC_FLAGS).encode())
    h.update(repr(sorted(EXTRA_FLAGS.items())).encode())
    return h.hexdigest()


def lib_path():
    return os.path.join(LIBDIR, LIBNAME)


# extra nvcc flags per source.  persist.cu runs the driver loop's scalar arithmetic (solve_scalar.h) ON the device: like the
# host build (-ffp-contract=off) it must round every product and sum separately; explicit fma() calls are unaffected.
EXTRA_FLAGS = {"persist.cu": ["-fmad=false"], "step_multi.cu": ["-fmad=false"]}


def _obj_stamp(src):
    """Hash of one source, every header it may include and its flags: only stale objects are recompiled."""
    h = hashlib.sha256()
    names = [src] + sorted(n for n in os.listdir(CSRC) if n.endswith((".cuh", ".h")))
    for name in names:
        h.update(name.encode())
        h.update(open(os.path.join(CSRC, name), "rb").read())
    h.update(open(os.path.join(INCLUDE, "proxb200.h"), "rb").read())
    h.update(" ".join(NVCC_FLAGS + EXTRA_FLAGS.get(src, [])).encode())
    return h.hexdigest()


def build(force=False, verbose=False):
    """Compile every .cu for sm_100a and link the shared library.  Returns the library path."""
    os.makedirs(LIBDIR, exist_ok=True)
    stamp_file = os.path.join(LIBDIR, "build.stamp")
    stamp = _stamp()
    if not force and os.path.exists(lib_path()) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return lib_path()
    nvcc = _nvcc()

    def compile_one(src):
        obj = os.path.join(LIBDIR, src.replace(".cu", ".o"))
        ostamp_file = obj + ".stamp"
        ostamp = _obj_stamp(src)
        if not force and os.path.exists(obj) and os.path.exists(ostamp_file) and open(ostamp_file).read() == ostamp:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *EXTRA_FLAGS.get(src, []), "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = r.stdout + r.stderr
        with open(os.path.join(LIBDIR, src.replace(".cu", ".ptxas.log")), "w") as fh:
            fh.write(log)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{log}")
        if verbose:
            print(log)
        with open(ostamp_file, "w") as fh:
            fh.write(ostamp)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path(), *objs, "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return lib_path()


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(p)
