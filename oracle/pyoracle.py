"""ctypes wrapper of oracle/libclothoracle.so — TEST INFRASTRUCTURE ONLY (see cloth_oracle.h).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
PARITY UNPINNED: the reference has no tests/fixtures and cannot be built here.
"""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(_HERE, "..", "cloth-physics_b200", "python"))
from clothgpu import _abi  # noqa: E402  (POD types only; does not load the CUDA library)

SEQ_FAITHFUL, SEQ_MASKED, COLOUR = 0, 1, 2

_lib = None


def build(force=False):
    so = os.path.join(_HERE, "libclothoracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("cloth_oracle.c", "cloth_oracle_impl.inc", "cloth_oracle.h")]
    srcs.append(os.path.join(_HERE, "..", "include", "clothgpu.h"))
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return so


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        dp, u8p = C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.oracle_create.restype = C.c_void_p
        L.oracle_create.argtypes = [C.POINTER(_abi.Desc), C.c_int]
        L.oracle_destroy.argtypes = [C.c_void_p]
        L.oracle_reset.argtypes = [C.c_void_p, C.c_int32, C.c_int32]
        L.oracle_step.argtypes = [C.c_void_p, C.POINTER(_abi.Frame), C.c_int]
        for n in ("oracle_nx", "oracle_ny", "oracle_held_rows", "oracle_held_row0"):
            getattr(L, n).restype = C.c_int32
            getattr(L, n).argtypes = [C.c_void_p]
        L.oracle_slice_len.restype = C.c_int64
        L.oracle_slice_len.argtypes = [C.c_void_p]
        L.oracle_relaxations.restype = C.c_uint64
        L.oracle_relaxations.argtypes = [C.c_void_p]
        L.oracle_torn_last.restype = C.c_uint64
        L.oracle_torn_last.argtypes = [C.c_void_p]
        L.oracle_read_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, dp, dp, dp, dp, u8p]
        L.oracle_write_rows.argtypes = [C.c_void_p, C.c_int32, C.c_int32, dp, dp, dp, dp, u8p]
        L.oracle_last_visit_order.restype = C.c_int64
        L.oracle_last_visit_order.argtypes = [C.c_void_p, C.POINTER(C.c_int64), C.c_int64]
        L.oracle_stick_update.argtypes = [dp, dp, dp, dp, C.c_int, C.c_int, C.c_double, C.c_int,
                                          C.c_double, C.c_double]
        L.oracle_segment_quad.argtypes = [C.POINTER(C.c_float), C.c_float, C.POINTER(C.c_float)]
        _lib = L
    return _lib


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


class OracleCloth:
    """NewCloth + Init of the reference (physics/cloth.go:31-81), stepped on the CPU."""

    def __init__(self, desc, mode=COLOUR):
        self._L = lib()
        self.desc = desc
        self.mode = mode
        self._h = self._L.oracle_create(C.byref(desc), mode)
        if not self._h:
            raise ValueError("oracle_create rejected the descriptor (the Go code would panic, "
                             "physics/cloth.go:72, or the strip/mode combination is unsupported)")
        self.nx = self._L.oracle_nx(self._h)
        self.ny = self._L.oracle_ny(self._h)
        self.held_row0 = self._L.oracle_held_row0(self._h)
        self.held_rows = self._L.oracle_held_rows(self._h)

    def close(self):
        if self._h:
            self._L.oracle_destroy(self._h)
            self._h = None

    __del__ = close

    def reset(self, pos_x, pos_y):
        self._L.oracle_reset(self._h, pos_x, pos_y)

    def step(self, frame, threads=1):
        self._L.oracle_step(self._h, C.byref(frame), threads)

    @property
    def slice_len(self):
        return int(self._L.oracle_slice_len(self._h))

    @property
    def relaxations(self):
        return int(self._L.oracle_relaxations(self._h))

    @property
    def torn_last(self):
        return int(self._L.oracle_torn_last(self._h))

    def read_rows(self, row0=None, rows=None):
        row0 = self.held_row0 if row0 is None else row0
        rows = self.held_rows if rows is None else rows
        n = rows * self.nx
        x, y, px, py = (np.empty(n, np.float64) for _ in range(4))
        fl = np.empty(n, np.uint8)
        rc = self._L.oracle_read_rows(self._h, row0, rows, _p(x, C.c_double), _p(y, C.c_double),
                                      _p(px, C.c_double), _p(py, C.c_double), _p(fl, C.c_uint8))
        if rc:
            raise ValueError("rows not held")
        return x, y, px, py, fl

    read_state = read_rows

    def write_rows(self, row0, rows, x=None, y=None, px=None, py=None, flags=None):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x, y, px, py)]
        fl = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        for a in arrs + [fl]:
            assert a is None or a.size == rows * self.nx
        rc = self._L.oracle_write_rows(self._h, row0, rows, *[_p(a, C.c_double) for a in arrs],
                                       _p(fl, C.c_uint8))
        if rc:
            raise ValueError("rows not held")

    def write_state(self, x=None, y=None, px=None, py=None, flags=None):
        self.write_rows(self.held_row0, self.held_rows, x, y, px, py, flags)

    def last_visit_order(self):
        n = self._L.oracle_last_visit_order(self._h, None, 0)
        ids = np.empty(max(n, 1), np.int64)
        self._L.oracle_last_visit_order(self._h, _p(ids, C.c_int64), n)
        return ids[:n]


def stick_update(x1, y1, x2, y2, pin1=False, pin2=False, length=6.0, dragging=False,
                 tear_length=150.0, relax_gain=0.35):
    """constraint.Update on one stick (physics/constraint.go:25-54)."""
    v = [C.c_double(a) for a in (x1, y1, x2, y2)]
    tore = lib().oracle_stick_update(*[C.byref(a) for a in v], int(pin1), int(pin2), length,
                                     int(dragging), tear_length, relax_gain)
    return tuple(a.value for a in v), bool(tore)


def segment_quads(xyxy, line_width=0.6):
    """addSegment (physics/cloth.go:150-157) for an (n, 4) float32 array of stick endpoints -> (n, 8) float32 corners."""
    xyxy = np.ascontiguousarray(xyxy, np.float32).reshape(-1, 4)
    out = np.empty((len(xyxy), 8), np.float32)
    L = lib()
    for i in range(len(xyxy)):
        L.oracle_segment_quad(xyxy[i].ctypes.data_as(C.POINTER(C.c_float)), line_width,
                              out[i].ctypes.data_as(C.POINTER(C.c_float)))
    return out
