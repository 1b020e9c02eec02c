"""Minimal reader for the JLD2 (HDF5-based) fixtures of the reference's benchmark suite.

The reference loads `benchmark/data/lasso_{tiny,small,medium}.jld2` with FileIO/JLD2 (benchmark/benchmarks.jl:30-45).  No HDF5
library exists in this image, so this module parses the subset of the HDF5 file format those files use: superblock version 2
(after JLD2's 512-byte text header), version-2 object headers (`OHDR`, with continuation blocks `OCHK`), hard links in the root
group, simple dataspaces, fixed-point / floating-point datatypes, contiguous and compact data layouts.  Anything else raises
`ValueError` naming the unsupported feature.  Host-side I/O only.

Julia arrays are column-major and JLD2 writes their dimensions reversed (HDF5 is row-major): `read_jld2` returns numpy arrays
in the Julia orientation (`A.shape == size(A)`), Fortran-ordered.
"""
from __future__ import annotations

import struct

import numpy as np

_SIG = b"\x89HDF\r\n\x1a\n"
_UNDEF = 0xFFFFFFFFFFFFFFFF


class _Reader:
    def __init__(self, buf):
        self.buf = buf
        base = buf.find(_SIG)
        if base < 0:
            raise ValueError("not an HDF5/JLD2 file (no superblock signature)")
        ver = buf[base + 8]
        if ver not in (2, 3):
            raise ValueError(f"unsupported superblock version {ver}")
        self.O, self.L = buf[base + 9], buf[base + 10]
        if self.O != 8 or self.L != 8:
            raise ValueError("only 8-byte offsets / lengths are supported")
        self.base_addr, _ext, _eof, self.root = struct.unpack_from("<QQQQ", buf, base + 12)

    # ---- object headers ------------------------------------------------------------------------------------------------
    def messages(self, addr):
        """Yield (type, payload_offset, payload_size) of every message of the version-2 object header at `addr`."""
        buf = self.buf
        addr += self.base_addr
        if buf[addr:addr + 4] != b"OHDR":
            raise ValueError(f"no version-2 object header at {addr}")
        if buf[addr + 4] != 2:
            raise ValueError("unsupported object header version")
        flags = buf[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16                      # access / modification / change / birth times
        if flags & 0x10:
            p += 4                       # max compact / min dense attributes
        nsz = 1 << (flags & 3)
        chunk0 = int.from_bytes(buf[p:p + nsz], "little")
        p += nsz
        track_order = bool(flags & 0x04)
        blocks = [(p, p + chunk0)]
        while blocks:
            q, end = blocks.pop(0)
            while q + 4 <= end:
                mtype, msize, _mflags = buf[q], int.from_bytes(buf[q + 1:q + 3], "little"), buf[q + 3]
                q += 4 + (2 if track_order else 0)
                if mtype == 0x10:        # continuation: offset, length of an OCHK block
                    off, ln = struct.unpack_from("<QQ", buf, q)
                    off += self.base_addr
                    if buf[off:off + 4] != b"OCHK":
                        raise ValueError("bad object header continuation block")
                    blocks.append((off + 4, off + ln - 4))
                elif mtype != 0:
                    yield mtype, q, msize
                q += msize

    def links(self, addr):
        """{name: object header address} of the hard links stored as link messages in the group at `addr`."""
        out = {}
        buf = self.buf
        for mtype, q, _ in self.messages(addr):
            if mtype == 0x02:            # link info: links live in link messages unless a fractal heap is given
                p = q + 2 + (8 if buf[q + 1] & 1 else 0)
                if struct.unpack_from("<Q", buf, p)[0] != _UNDEF:
                    raise ValueError("dense link storage (fractal heap) is not supported")
                continue
            if mtype != 0x06:
                continue
            if buf[q] != 1:
                raise ValueError("unsupported link message version")
            flags = buf[q + 1]
            p = q + 2
            ltype = 0
            if flags & 0x08:
                ltype = buf[p]
                p += 1
            if flags & 0x04:
                p += 8
            if flags & 0x10:
                p += 1
            nsz = 1 << (flags & 3)
            nlen = int.from_bytes(buf[p:p + nsz], "little")
            p += nsz
            name = buf[p:p + nlen].decode()
            p += nlen
            if ltype != 0:
                continue                 # soft / external links: not data
            out[name] = struct.unpack_from("<Q", buf, p)[0]
        return out

    # ---- datasets ---------------------------------------------------------------------------------------------------------
    def dataset(self, addr):
        buf = self.buf
        dims, dtype, data = None, None, None
        for mtype, q, msize in self.messages(addr):
            if mtype == 0x01:            # dataspace
                ver, rank, flags = buf[q], buf[q + 1], buf[q + 2]
                if ver == 1:
                    p = q + 8
                elif ver == 2:
                    p = q + 4
                else:
                    raise ValueError("unsupported dataspace version")
                dims = struct.unpack_from("<" + "Q" * rank, buf, p) if rank else ()
                _ = flags
            elif mtype == 0x03:          # datatype
                cls, size = buf[q] & 0x0F, struct.unpack_from("<I", buf, q + 4)[0]
                bits0 = buf[q + 1]
                if bits0 & 1:
                    raise ValueError("big-endian data is not supported")
                if cls == 1:
                    dtype = {4: "<f4", 8: "<f8"}.get(size)
                elif cls == 0:
                    signed = bool(bits0 & 0x08)
                    dtype = {1: "i1", 2: "<i2", 4: "<i4", 8: "<i8"}.get(size)
                    if dtype and not signed:
                        dtype = dtype.replace("i", "u")
                if dtype is None:
                    return None          # strings, compounds, references ...: not numeric data
            elif mtype == 0x08:          # data layout
                ver, lclass = buf[q], buf[q + 1]
                if ver not in (3, 4):
                    raise ValueError("unsupported data layout version")
                if lclass == 0:          # compact
                    size = struct.unpack_from("<H", buf, q + 2)[0]
                    data = (q + 4, size)
                elif lclass == 1:        # contiguous
                    off, size = struct.unpack_from("<QQ", buf, q + 2)
                    data = (None, 0) if off == _UNDEF else (off + self.base_addr, size)
                else:
                    raise ValueError("chunked / virtual layouts are not supported")
        if dtype is None or data is None or dims is None:
            return None
        count = int(np.prod(dims)) if dims else 1
        if data[0] is None:
            arr = np.zeros(count, dtype)
        else:
            arr = np.frombuffer(buf, dtype, count, data[0]).copy()
        if not dims:
            return arr[0]
        return np.asfortranarray(arr.reshape(dims).T)       # HDF5 dims are the reversed Julia dims


def read_jld2(path):
    """{name: numpy array or scalar} of the numeric datasets in the root group of a JLD2 file."""
    with open(path, "rb") as fh:
        r = _Reader(fh.read())
    out = {}
    for name, addr in r.links(r.root).items():
        try:
            v = r.dataset(addr)
        except ValueError:
            continue
        if v is not None:
            out[name] = v
    return out


def load_lasso_fixture(path):
    """(A, b, lam, xstar, ystar) of a `benchmark/data/lasso_*.jld2` file (benchmark/benchmarks.jl:33-36)."""
    d = read_jld2(path)
    return d["A"], d["b"], float(d["lambda"]), d.get("xstar"), d.get("ystar")


__all__ = ["read_jld2", "load_lasso_fixture"]
