"""Reader / writer of the CLSTATE1 snapshot file (include/clothgpu.h, clothgpu_save_state) in plain numpy — for tools and
tests: the CPU checker of tests/ and the Go harness of the reference (host/go/physics/dump_harness.go) exchange states with the GPU
library through this one format.  No physics here."""
import struct

import numpy as np

MAGIC = b"CLSTATE1"
_HDR = struct.Struct("<8sII4i4iQQ")      # 64 bytes
_FNV_BASIS, _FNV_PRIME, _M64 = 0xcbf29ce484222325, 0x100000001b3, (1 << 64) - 1


def fnv1a64(data, h=_FNV_BASIS):
    """FNV-1a 64 of a bytes-like object (vectorised: the hash of a block is a polynomial in the prime modulo 2^64)."""
    a = np.frombuffer(data, np.uint8)
    for i in range(0, len(a), 1 << 16):          # h' = (h ^ b) * p per byte; keep it simple and exact in chunks
        for b in a[i:i + (1 << 16)].tolist():
            h = ((h ^ b) * _FNV_PRIME) & _M64
    return h


def write(path, nx, ny, cloths, *, row0=0, rows=None, spacing=6, pos=(128, 164), frames=0):
    """cloths: list of (x, y, px, py, flags) arrays of rows*nx elements each."""
    rows = ny if rows is None else rows
    n = nx * rows
    out = bytearray(_HDR.pack(MAGIC, 64, 1, nx, ny, row0, rows, len(cloths), spacing, pos[0], pos[1], frames, 0))
    for x, y, px, py, fl in cloths:
        for a in (x, y, px, py):
            a = np.ascontiguousarray(a, "<f8")
            assert a.size == n
            out += a.tobytes()
        f = np.zeros((n + 7) // 8 * 8, np.uint8)
        f[:n] = np.asarray(fl, np.uint8)
        out += f.tobytes()
    out += struct.pack("<Q", fnv1a64(bytes(out)))
    with open(path, "wb") as fh:
        fh.write(out)


def read(path, verify=True):
    """-> (header dict, [(x, y, px, py, flags) per cloth])"""
    data = open(path, "rb").read()
    magic, hb, ver, nx, ny, row0, rows, ens, spacing, px0, py0, frames, _ = _HDR.unpack_from(data)
    if magic != MAGIC or hb != 64 or ver != 1:
        raise ValueError(f"{path} is not a CLSTATE1 file")
    n = nx * rows
    nfl = (n + 7) // 8 * 8
    if len(data) != 64 + ens * (32 * n + nfl) + 8:
        raise ValueError(f"{path} is truncated")
    if verify and struct.unpack_from("<Q", data, len(data) - 8)[0] != fnv1a64(data[:-8]):
        raise ValueError(f"{path}: checksum mismatch")
    cloths, at = [], 64
    for _ in range(ens):
        planes = []
        for _k in range(4):
            planes.append(np.frombuffer(data, "<f8", n, at).copy())
            at += 8 * n
        fl = np.frombuffer(data, np.uint8, n, at).copy()
        at += nfl
        cloths.append((*planes, fl))
    hdr = dict(nx=nx, ny=ny, row0=row0, rows=rows, ensemble=ens, spacing=spacing, pos=(px0, py0), frames=frames)
    return hdr, cloths
