"""clothgpu — ctypes binding of libclothgpu.so (include/clothgpu.h).

This is a test/bench harness over the C ABI, the same entry points the cgo shim of INTEGRATION.md
binds.  It never computes anything itself and has no CPU path: if the CUDA library is missing or no
B200 is visible, it raises.
"""
import ctypes as C
import os

import numpy as np

from . import _abi
from ._abi import (Counters, Desc, Frame, Info, default_desc, default_frame, grid_desc)  # noqa: F401
from ._abi import *  # noqa: F401,F403  (constants)

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.normpath(os.path.join(_PKG, "..", "..", "csrc", "libclothgpu.so"))
if os.environ.get("CLOTHGPU_LIB"):      # A/B timing of another build of the same library (tools/ab_time.sh); never a fallback
    LIB_PATH = os.path.abspath(os.environ["CLOTHGPU_LIB"])

# Every symbol include/clothgpu.h declares.
EXPORTS = [
    "clothgpu_frame_default", "clothgpu_desc_default", "clothgpu_create", "clothgpu_reset",
    "clothgpu_destroy", "clothgpu_step", "clothgpu_step_n", "clothgpu_step_each", "clothgpu_sync",
    "clothgpu_get_info", "clothgpu_set_schedule", "clothgpu_read_state", "clothgpu_write_state",
    "clothgpu_read_masks", "clothgpu_read_segments_f32", "clothgpu_read_quads_f32", "clothgpu_host_alloc", "clothgpu_host_free",
    "clothgpu_compact_segments",
    "clothgpu_get_counters",
    "clothgpu_halo_region", "clothgpu_strip_export", "clothgpu_strip_attach", "clothgpu_strip_push", "clothgpu_set_stream", "clothgpu_last_step_ms", "clothgpu_launch_count",
    "clothgpu_last_schedule", "clothgpu_selftest_stickmath", "clothgpu_selftest_stickmath_f32", "clothgpu_save_state", "clothgpu_load_state",
    "clothgpu_last_error", "clothgpu_version",
]

_lib = None


class ClothGpuError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"clothgpu error {code}: {msg}")
        self.code = code


def lib():
    """Load libclothgpu.so (built in-tree by __graft_entry__.build()).  No fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ClothGpuError(_abi.ERR_NODEVICE,
                                f"{LIB_PATH} not built; run python -c 'import __graft_entry__ as g; g.build()'")
        L = C.CDLL(LIB_PATH)
        dp, u8p, u32p = C.POINTER(C.c_double), C.POINTER(C.c_uint8), C.POINTER(C.c_uint32)
        vp = C.c_void_p
        L.clothgpu_frame_default.argtypes = [C.POINTER(Frame)]
        L.clothgpu_desc_default.argtypes = [C.POINTER(Desc)]
        L.clothgpu_create.argtypes = [C.POINTER(Desc), C.POINTER(vp)]
        L.clothgpu_reset.argtypes = [vp, C.c_int32, C.c_int32]
        L.clothgpu_destroy.argtypes = [vp]
        L.clothgpu_destroy.restype = None
        L.clothgpu_step.argtypes = [vp, C.POINTER(Frame)]
        L.clothgpu_step_n.argtypes = [vp, C.POINTER(Frame), C.c_int32]
        L.clothgpu_step_each.argtypes = [vp, C.POINTER(Frame), C.c_int32]
        L.clothgpu_sync.argtypes = [vp]
        L.clothgpu_get_info.argtypes = [vp, C.POINTER(Info)]
        L.clothgpu_set_schedule.argtypes = [vp, C.c_int32]
        L.clothgpu_read_state.argtypes = [vp, C.c_int32, dp, dp, dp, dp, u8p]
        L.clothgpu_write_state.argtypes = [vp, C.c_int32, dp, dp, dp, dp, u8p]
        if hasattr(L, "clothgpu_save_state"):
            L.clothgpu_save_state.argtypes = [vp, C.c_char_p]
            L.clothgpu_load_state.argtypes = [vp, C.c_char_p]
        L.clothgpu_read_masks.argtypes = [vp, C.c_int32, u32p, u32p, u32p, u32p, u32p]
        L.clothgpu_read_segments_f32.argtypes = [vp, C.c_int32, C.POINTER(C.c_float), u8p, u32p,
                                                 C.c_size_t, C.POINTER(C.c_size_t)]
        L.clothgpu_read_quads_f32.argtypes = [vp, C.c_int32, C.c_float, C.POINTER(C.c_float), u8p,
                                              C.c_size_t, C.POINTER(C.c_size_t)]
        L.clothgpu_compact_segments.argtypes = [vp, C.c_int32]
        if hasattr(L, "clothgpu_host_alloc"):
            L.clothgpu_host_alloc.argtypes = [C.c_size_t, C.POINTER(C.c_void_p)]
            L.clothgpu_host_free.argtypes = [C.c_void_p]
        L.clothgpu_get_counters.argtypes = [vp, C.POINTER(Counters)]
        L.clothgpu_halo_region.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp), C.POINTER(C.c_size_t)]
        L.clothgpu_strip_export.argtypes = [vp, vp, C.c_size_t, C.POINTER(C.c_size_t)]
        L.clothgpu_strip_attach.argtypes = [vp, vp, vp]
        L.clothgpu_strip_push.argtypes = [vp]
        L.clothgpu_set_stream.argtypes = [vp, vp]
        L.clothgpu_last_step_ms.argtypes = [vp, C.POINTER(C.c_float)]
        L.clothgpu_launch_count.argtypes = [vp, C.POINTER(C.c_uint64)]
        if hasattr(L, "clothgpu_last_schedule"):    # (absent from the round-1 build used for A/B timing)
            L.clothgpu_last_schedule.argtypes = [vp, C.POINTER(C.c_int32), C.POINTER(C.c_char_p)]
            L.clothgpu_selftest_stickmath.argtypes = [C.c_int32, C.c_uint64, C.c_uint64, dp, C.c_size_t,
                                                      C.POINTER(_abi.StickMathReport)]
        if hasattr(L, "clothgpu_selftest_stickmath_f32"):
            L.clothgpu_selftest_stickmath_f32.argtypes = [C.c_int32, C.POINTER(_abi.StickMathReport)]
        L.clothgpu_last_error.restype = C.c_char_p
        L.clothgpu_version.restype = C.c_char_p
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise ClothGpuError(rc, lib().clothgpu_last_error().decode())


def _p(a, t):
    return None if a is None else a.ctypes.data_as(C.POINTER(t))


def frame_array(frames):
    arr = (Frame * len(frames))()
    for i, f in enumerate(frames):
        C.memmove(C.byref(arr[i]), C.byref(f), C.sizeof(Frame))
    return arr


def selftest_stickmath(samples=0, s_list=None, seed=1, device=0):
    """Device self-test of the binary64 stick arithmetic against sqrt.rn / div.rn (include/clothgpu.h)."""
    rep = _abi.StickMathReport()
    arr = None if s_list is None else np.ascontiguousarray(s_list, np.float64)
    _check(lib().clothgpu_selftest_stickmath(device, seed, samples, _p(arr, C.c_double), 0 if arr is None else arr.size,
                                             C.byref(rep)))
    return rep.as_dict()


def selftest_stickmath_f32(device=0):
    """Exhaustive device self-test of the binary32 stick arithmetic (include/clothgpu.h)."""
    rep = _abi.StickMathReport()
    _check(lib().clothgpu_selftest_stickmath_f32(device, C.byref(rep)))
    return rep.as_dict()


class HostBuffer:
    """Page-locked host memory of clothgpu_host_alloc as a numpy array (the export writes such buffers directly)."""

    def __init__(self, shape, dtype):
        import numpy as np
        self._L = lib()
        self.shape, self.dtype = tuple(np.atleast_1d(shape)), np.dtype(dtype)
        self.nbytes = int(np.prod(self.shape)) * self.dtype.itemsize
        p = C.c_void_p()
        _check(self._L.clothgpu_host_alloc(self.nbytes, C.byref(p)))
        self.ptr = p.value
        self.array = np.ctypeslib.as_array((C.c_uint8 * max(1, self.nbytes)).from_address(self.ptr))[:self.nbytes].view(self.dtype).reshape(self.shape)

    def close(self):
        if self.ptr:
            self.array = None
            _check(self._L.clothgpu_host_free(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Cloth:
    """One device-resident cloth (or ensemble / strip): NewCloth + Init of the reference
    (physics/cloth.go:31-81) behind the C ABI."""

    def __init__(self, desc):
        self._L = lib()
        self._h = C.c_void_p()
        _check(self._L.clothgpu_create(C.byref(desc), C.byref(self._h)))
        self.desc = desc
        self.info = Info()
        _check(self._L.clothgpu_get_info(self._h, C.byref(self.info)))
        self.nx, self.ny = self.info.nx, self.info.ny
        self.rows, self.row0 = self.info.rows, self.info.row0

    def close(self):
        if getattr(self, "_h", None):
            self._L.clothgpu_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    # (*Cloth).Reset, physics/cloth.go:142-148
    def reset(self, pos_x, pos_y):
        _check(self._L.clothgpu_reset(self._h, pos_x, pos_y))

    # physics half of (*Cloth).Update, physics/cloth.go:92-100
    def step(self, frame):
        _check(self._L.clothgpu_step(self._h, C.byref(frame)))

    def step_n(self, frames):
        arr = frames if isinstance(frames, C.Array) else frame_array(frames)
        _check(self._L.clothgpu_step_n(self._h, arr, len(arr)))

    def step_each(self, frames):
        arr = frames if isinstance(frames, C.Array) else frame_array(frames)
        _check(self._L.clothgpu_step_each(self._h, arr, len(arr)))

    def sync(self):
        _check(self._L.clothgpu_sync(self._h))

    def set_schedule(self, s):
        _check(self._L.clothgpu_set_schedule(self._h, s))

    def set_stream(self, cuda_stream):
        _check(self._L.clothgpu_set_stream(self._h, C.c_void_p(cuda_stream)))

    def read_state(self, cloth=0):
        n = self.nx * self.rows
        x, y, px, py = (np.empty(n, np.float64) for _ in range(4))
        fl = np.empty(n, np.uint8)
        _check(self._L.clothgpu_read_state(self._h, cloth, _p(x, C.c_double), _p(y, C.c_double),
                                           _p(px, C.c_double), _p(py, C.c_double), _p(fl, C.c_uint8)))
        return x, y, px, py, fl

    def write_state(self, x=None, y=None, px=None, py=None, flags=None, cloth=0):
        arrs = [None if a is None else np.ascontiguousarray(a, np.float64) for a in (x, y, px, py)]
        fl = None if flags is None else np.ascontiguousarray(flags, np.uint8)
        for a in arrs + [fl]:
            assert a is None or a.size == self.nx * self.rows
        _check(self._L.clothgpu_write_state(self._h, cloth, *[_p(a, C.c_double) for a in arrs],
                                            _p(fl, C.c_uint8)))

    def save_state(self, path):
        """Snapshot of the whole handle on disk (CLSTATE1, include/clothgpu.h)."""
        _check(self._L.clothgpu_save_state(self._h, os.fsencode(path)))

    def load_state(self, path):
        _check(self._L.clothgpu_load_state(self._h, os.fsencode(path)))

    def read_masks(self, cloth=0):
        w = (self.nx * self.rows + 31) // 32
        planes = [np.zeros(w, np.uint32) for _ in range(5)]
        _check(self._L.clothgpu_read_masks(self._h, cloth, *[_p(a, C.c_uint32) for a in planes]))
        return dict(zip(("pin", "active", "highlighted", "alive_h", "alive_v"), planes))

    def read_segments(self, cloth=0, with_indices=False):
        cap = 2 * self.nx * self.rows
        xyxy = np.empty((cap, 4), np.float32)
        hl = np.empty(cap, np.uint8)
        idx = np.empty((cap, 2), np.uint32) if with_indices else None
        n = C.c_size_t(0)
        _check(self._L.clothgpu_read_segments_f32(self._h, cloth, _p(xyxy, C.c_float), _p(hl, C.c_uint8),
                                                  _p(idx, C.c_uint32), cap, C.byref(n)))
        if with_indices:
            return xyxy[:n.value], hl[:n.value], idx[:n.value]
        return xyxy[:n.value], hl[:n.value]

    def read_quads(self, cloth=0, line_width=0.6):
        """addSegment of every live stick (physics/cloth.go:150-157): (n, 8) float32 corners and the highlight bytes."""
        cap = 2 * self.nx * self.rows
        quads = np.empty((cap, 8), np.float32)
        hl = np.empty(cap, np.uint8)
        n = C.c_size_t(0)
        _check(self._L.clothgpu_read_quads_f32(self._h, cloth, C.c_float(line_width), _p(quads, C.c_float),
                                               _p(hl, C.c_uint8), cap, C.byref(n)))
        return quads[:n.value], hl[:n.value]

    def read_quads_into(self, quads_ptr, hl_ptr, cap, cloth=0, line_width=0.6):
        """clothgpu_read_quads_f32 into caller-owned (e.g. pinned) host memory given as raw addresses; returns n."""
        n = C.c_size_t(0)
        _check(self._L.clothgpu_read_quads_f32(self._h, cloth, C.c_float(line_width), C.cast(quads_ptr, C.POINTER(C.c_float)),
                                               C.cast(hl_ptr, C.POINTER(C.c_uint8)), cap, C.byref(n)))
        return n.value

    def read_segments_into(self, seg_ptr, hl_ptr, idx_ptr, cap, cloth=0):
        """clothgpu_read_segments_f32 into caller-owned (e.g. pinned) host memory given as raw addresses (0 = not wanted)."""
        n = C.c_size_t(0)
        _check(self._L.clothgpu_read_segments_f32(self._h, cloth, C.cast(seg_ptr, C.POINTER(C.c_float)), C.cast(hl_ptr, C.POINTER(C.c_uint8)),
                                                  C.cast(idx_ptr, C.POINTER(C.c_uint32)), cap, C.byref(n)))
        return n.value

    def compact_segments(self, cloth=0):
        """Device-side live-stick compaction only (asynchronous)."""
        _check(self._L.clothgpu_compact_segments(self._h, cloth))

    def count_segments(self, cloth=0):
        n = C.c_size_t(0)
        _check(self._L.clothgpu_read_segments_f32(self._h, cloth, None, None, None, 0, C.byref(n)))
        return n.value

    def counters(self):
        c = Counters()
        _check(self._L.clothgpu_get_counters(self._h, C.byref(c)))
        return c.as_dict()

    def halo_region(self, plane, which):
        ptr, nbytes = C.c_void_p(), C.c_size_t()
        _check(self._L.clothgpu_halo_region(self._h, plane, which, C.byref(ptr), C.byref(nbytes)))
        return ptr.value, nbytes.value

    # row strips over several GPUs (include/clothgpu.h)
    def strip_export(self):
        n = C.c_size_t()
        _check(self._L.clothgpu_strip_export(self._h, None, 0, C.byref(n)))
        buf = C.create_string_buffer(n.value)
        _check(self._L.clothgpu_strip_export(self._h, buf, n.value, C.byref(n)))
        return buf.raw[:n.value]

    def strip_attach(self, upper_blob=None, lower_blob=None):
        _check(self._L.clothgpu_strip_attach(self._h, upper_blob, lower_blob))

    def strip_push(self):
        _check(self._L.clothgpu_strip_push(self._h))

    def last_step_ms(self):
        ms = C.c_float()
        _check(self._L.clothgpu_last_step_ms(self._h, C.byref(ms)))
        return ms.value

    def last_schedule(self):
        """(clothgpu_schedule, kernel name) the last step actually ran."""
        sched, name = C.c_int32(), C.c_char_p()
        _check(self._L.clothgpu_last_schedule(self._h, C.byref(sched), C.byref(name)))
        return sched.value, name.value.decode()

    def launch_count(self):
        n = C.c_uint64()
        _check(self._L.clothgpu_launch_count(self._h, C.byref(n)))
        return n.value
