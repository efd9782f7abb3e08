"""GUI-sized cloths: clothgpu_step_n runs all its frames in ONE launch of a resident CTA that keeps the whole cloth in
shared memory (csrc/cloth_small.cuh, CLOTHGPU_SCHED_RESIDENT) — the persistent multi-frame loop of SURVEY §7 step 5.
Bit-exact against the oracle frame by frame, counters included."""
import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import FLAG_ACTIVE, FLAG_ALIVE_H, FLAG_ALIVE_V, FLAG_PIN, MOUSE_DRAGGING, SCHED_RESIDENT
from traces import c0_trace

from test_gpu_counters import census
from test_gpu_parity import assert_same_state, random_state

pytestmark = pytest.mark.gpu


def test_default_cloth_trace_in_multi_frame_launches(oracle):
    """BASELINE configs[0]: the 1000-frame scripted trace in calls of 1, 2, 7, 50, 240, 700 frames on the resident schedule;
    after every call the state, the counters and the launch count are checked."""
    desc = _abi.default_desc()
    g = clothgpu.Cloth(desc)
    g.set_schedule(SCHED_RESIDENT)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    frames = c0_trace(1000)
    at = 0
    for n in (1, 2, 7, 50, 240, 700):
        l0 = g.launch_count()
        g.step_n(frames[at:at + n])
        assert g.launch_count() - l0 == 1, "one launch per call"
        for f in frames[at:at + n]:
            o.step(f)
        at += n
        assert_same_state(g.read_state(), o.read_state(), f"after {at} frames")
        c = g.counters()
        want = census(171, 36, o.read_state()[4])
        assert {k: c[k] for k in want} == want, at
        assert (c["frames"], c["relaxations"], c["torn_last"]) == (at, o.relaxations, o.torn_last)
        assert g.last_schedule() == (SCHED_RESIDENT, "small_frames_kernel<double>")
    assert at == 1000 and c["pinned"] > 11 and c["inactive"] > 0


@pytest.mark.parametrize("nx,ny,K,precision", [(171, 36, 1, _abi.F64), (64, 100, 3, _abi.F64), (83, 84, 8, _abi.F64), (2, 2, 2, _abi.F64),
                                               (1, 9, 1, _abi.F64), (160, 70, 2, _abi.F32)])
def test_resident_schedule_on_random_states(oracle, nx, ny, K, precision):
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K, precision=precision)
    g = clothgpu.Cloth(desc)
    g.set_schedule(SCHED_RESIDENT)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    st = random_state(nx, ny, seed=nx * 7 + ny + K)
    if precision == _abi.F32:
        st = tuple(a.astype(np.float32).astype(np.float64) if a.dtype == np.float64 else a for a in st)
    g.write_state(*st)
    o.write_state(*st)
    fr = []
    for t in range(6):
        f = _abi.default_frame(K if t != 3 else K + 1)           # the sweep count may change from frame to frame
        f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20
        f.buttons = MOUSE_DRAGGING if t % 2 == 0 else _abi.MOUSE_RIGHT
        f.tear_length = 9.5
        f.mouse_x, f.mouse_y = 128.0 + 3 * nx + 5 * t, 164.0 + 3 * ny - 4 * t
        f.mouse_px, f.mouse_py = f.mouse_x - 8.0, f.mouse_y + 6.0
        fr.append(f)
    g.step(fr[0])                                                # a single step on the forced schedule
    g.step_n(fr[1:])
    for f in fr:
        o.step(f)
    assert g.last_schedule()[0] == SCHED_RESIDENT
    assert_same_state(g.read_state(), o.read_state(), f"resident {nx}x{ny} K={K}")
    c = g.counters()
    want = census(nx, ny, o.read_state()[4])
    assert {k: c[k] for k in want} == want
    assert c["relaxations"] == o.relaxations and c["torn_last"] == o.torn_last


def test_cloths_too_large_for_one_cta_fall_back():
    g = clothgpu.Cloth(_abi.grid_desc(300, 300, max_iterations=1))
    g.set_schedule(SCHED_RESIDENT)
    g.step_n([_abi.default_frame(1)] * 3)
    assert g.last_schedule()[0] != SCHED_RESIDENT and g.counters()["frames"] == 3
    e = clothgpu.Cloth(_abi.grid_desc(40, 40, max_iterations=1, ensemble=3))
    e.set_schedule(SCHED_RESIDENT)
    e.step_n([_abi.default_frame(1)] * 3)                        # ensembles keep their own kernels
    assert e.last_schedule()[0] != SCHED_RESIDENT
    a = clothgpu.Cloth(_abi.default_desc())
    a.step_n([_abi.default_frame(1)] * 3)                        # AUTO: one multi-CTA launch per frame is faster (measured)
    assert a.last_schedule()[0] != SCHED_RESIDENT
