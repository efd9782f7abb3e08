"""ctypes mirror of include/clothgpu.h (POD types and constants only).

These are the inputs (*Cloth).Update reads in the reference (physics/cloth.go:85,
physics/particle.go:75-181, physics/mouse.go:9-19, gui/hud.go:86-97), flattened into the two PODs the
C ABI takes.
"""
import ctypes as C

# clothgpu_status
OK, ERR_INVALID, ERR_CUDA, ERR_NODEVICE, ERR_NOMEM, ERR_STATE, ERR_CAPACITY = 0, -1, -2, -3, -4, -5, -6

PIN_REFERENCE, PIN_TOP_ROW, PIN_NONE = 0, 1, 2
F64, F32 = 0, 1
SCHED_AUTO, SCHED_STREAMING, SCHED_FUSED, SCHED_ROWSTREAM, SCHED_WAVE, SCHED_RESIDENT = 0, 1, 2, 3, 4, 5

FLAG_PIN, FLAG_ACTIVE, FLAG_HIGHLIGHT, FLAG_ALIVE_H, FLAG_ALIVE_V = 1, 2, 4, 8, 16

MOUSE_LEFT, MOUSE_RIGHT, MOUSE_DRAGGING, MOUSE_CTRL = 1, 2, 4, 8

SLIDER_DRAG_FORCE, SLIDER_GRAVITY, SLIDER_STIFFNESS, SLIDER_FRICTION, SLIDER_TEAR_DISTANCE = range(5)


class Desc(C.Structure):
    _fields_ = [
        ("width", C.c_int32), ("height", C.c_int32), ("spacing", C.c_int32),
        ("pos_x", C.c_int32), ("pos_y", C.c_int32),
        ("pin_rule", C.c_int32), ("precision", C.c_int32), ("device", C.c_int32),
        ("ensemble", C.c_int32),
        ("strip_row0", C.c_int32), ("strip_rows", C.c_int32), ("max_iterations", C.c_int32),
        ("reserved", C.c_int32 * 4),
    ]


class Frame(C.Structure):
    _fields_ = [
        ("slider", C.c_float * 5), ("drag_force_max", C.c_float),
        ("mouse_x", C.c_double), ("mouse_y", C.c_double),
        ("mouse_px", C.c_double), ("mouse_py", C.c_double),
        ("mouse_force", C.c_double),
        ("scroll_y", C.c_float), ("buttons", C.c_uint32),
        ("win_w", C.c_int32), ("win_h", C.c_int32),
        ("win_off_x", C.c_double), ("win_off_y", C.c_double),
        ("dt", C.c_double),
        ("iterations", C.c_int32), ("reserved0", C.c_int32),
        ("tear_length", C.c_double), ("relax_gain", C.c_double),
    ]

    def copy(self):
        f = Frame()
        C.memmove(C.byref(f), C.byref(self), C.sizeof(Frame))
        return f


class Counters(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "frames", "live_sticks", "alive_sticks", "torn_last", "pinned", "inactive", "highlighted",
        "relaxations")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


class StickMathReport(C.Structure):
    _fields_ = [("tested", C.c_uint64), ("mismatches", C.c_uint64)] + [(n, C.c_double) for n in (
        "first_s", "first_L", "got_dist", "want_dist", "got_mul", "want_mul")]

    def as_dict(self):
        return {n: getattr(self, n) for n, _ in self._fields_}


class Info(C.Structure):
    _fields_ = [
        ("nx", C.c_int32), ("ny", C.c_int32), ("row0", C.c_int32), ("rows", C.c_int32),
        ("ghost", C.c_int32), ("pitch", C.c_int32), ("ensemble", C.c_int32), ("precision", C.c_int32),
        ("particles", C.c_int64), ("sticks", C.c_int64), ("device", C.c_int32), ("sm_count", C.c_int32),
    ]


def default_frame(iterations=1):
    """Reference defaults without needing the CUDA library (same values clothgpu_frame_default
    writes): HUD slider defaults gui/hud.go:86-92, consts/consts.go:9,11-12,18."""
    f = Frame()
    f.slider[:] = [2.0, 250.0, 30.0, 0.98, 15.0]
    f.drag_force_max = 15.0
    f.scroll_y = 50.0
    f.win_w, f.win_h = 1280, 820
    f.dt = 0.022
    f.iterations = iterations
    f.tear_length = 150.0
    f.relax_gain = 0.35
    return f


def default_desc():
    """The reference's default GUI cloth (main.go:149-166): 1024 x 211 px, spacing 6, at (128,164)."""
    d = Desc()
    d.width, d.height, d.spacing = 1024, 211, 6
    d.pos_x, d.pos_y = 128, 164
    d.pin_rule = PIN_REFERENCE
    d.precision = F64
    d.ensemble = 1
    d.max_iterations = 1
    return d


def grid_desc(nx, ny, spacing=6, pos_x=128, pos_y=164, pin_rule=PIN_TOP_ROW, precision=F64,
              ensemble=1, device=0, max_iterations=8, strip_row0=0, strip_rows=0):
    """Descriptor of an nx x ny particle cloth (width = (nx-1)*spacing as NewCloth would get)."""
    d = Desc()
    d.width, d.height, d.spacing = (nx - 1) * spacing, (ny - 1) * spacing, spacing
    d.pos_x, d.pos_y = pos_x, pos_y
    d.pin_rule, d.precision, d.device, d.ensemble = pin_rule, precision, device, ensemble
    d.strip_row0, d.strip_rows, d.max_iterations = strip_row0, strip_rows, max_iterations
    return d
