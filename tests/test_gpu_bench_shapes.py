"""The benchmarked kernels at the benchmarked shapes, against the oracle (OpenMP over the host cores; the colour-ordered
mode does not depend on the thread count):

  * the register-streaming kernel (one sweep per frame) on 4096 x 4096 and on 16384 x 2048 — the shape one of eight strips
    of the 16384^2 cloth has, which takes the multi-wave chunk planner — with a tearing drag,
  * 60 frames of the scripted drag / tear / hole trace scaled x24 (BASELINE configs[2]) on 4096 x 4096, including the
    live-stick export of the final state,
  * the packed-ensemble instantiation of the wavefront kernel on 640 cloths of 128 x 128 (several CTA spans, the AUTO
    schedule's own choice).
"""
import os
import sys

import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import FLAG_ACTIVE, FLAG_ALIVE_H, FLAG_ALIVE_V, MOUSE_DRAGGING, MOUSE_LEFT, SCHED_ROWSTREAM

from test_gpu_parity import _expected_segments_fast, assert_same_state, random_state

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

pytestmark = pytest.mark.gpu
THREADS = max(1, min(32, os.cpu_count() or 1))


def jittered(nx, ny, seed):
    """random_state without the per-particle Python cost concerns: numpy only, same recipe."""
    return random_state(nx, ny, seed=seed, kill=0.005, pins=0.001, holes=0.002)


@pytest.mark.parametrize("nx,ny", [(4096, 4096), (16384, 2048)])
def test_rowstream_kernel_at_benchmark_shapes(oracle, nx, ny):
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=1)
    g = clothgpu.Cloth(desc)
    g.set_schedule(SCHED_ROWSTREAM)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    st = jittered(nx, ny, seed=nx + ny)
    g.write_state(*st)
    o.write_state(*st)
    f = _abi.default_frame(1)
    f.win_w, f.win_h = 128 + 6 * nx + 128, 164 + 6 * ny * 4
    f.buttons = MOUSE_LEFT | MOUSE_DRAGGING
    f.mouse_force = 13.0
    f.tear_length = 9.0                       # the jitter alone stretches sticks past it: tears all over the cloth
    for t in range(3):
        f.mouse_px, f.mouse_py = 128.0 + 3.0 * nx + 40 * t, 164.0 + 3.0 * ny + 25 * t
        f.mouse_x, f.mouse_y = f.mouse_px + 31.0, f.mouse_py - 17.0
        g.step(f)
        o.step(f, threads=THREADS)
    assert g.last_schedule()[1] == "stream1_frame_kernel<double,DRAG>"
    assert_same_state(g.read_state(), o.read_state(), f"rowstream {nx}x{ny}")
    c = g.counters()
    assert c["relaxations"] == o.relaxations
    assert c["torn_last"] == o.torn_last > 0


def test_scripted_trace_x24_on_4096_with_live_stick_export(oracle):
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    nx = ny = 4096
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=1)
    g = clothgpu.Cloth(desc)                                  # AUTO, as bench.py runs it
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    frames = bench.make_frames("c2t", 420, nx, ny, 1)         # window [380, 800) of the x24 trace
    # 20 frames from the end of the slow drag and the release, 20 of the right-button hole punch, 20 of the fast zig-zag
    picked = frames[0:20] + frames[100:120] + frames[240:260]
    torn = 0
    for f in picked:
        g.step(f)
        o.step(f, threads=THREADS)
        torn += o.torn_last
    gs, os_ = g.read_state(), o.read_state()
    assert_same_state(gs, os_, "c2t trace, 60 frames")
    c = g.counters()
    fl = os_[4]
    assert c["relaxations"] == o.relaxations
    assert c["inactive"] == int(((fl & FLAG_ACTIVE) == 0).sum()) > 0, "the right button punches holes"
    assert torn > 0 and c["alive_sticks"] == 2 * nx * ny - nx - ny - torn, "the drag tears sticks"
    want, want_hl = _expected_segments_fast(nx, ny, os_[0], os_[1], fl)
    got, got_hl = g.read_segments()
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got_hl, want_hl)
    assert c["live_sticks"] == len(want)


def test_packed_ensemble_of_640_cloths(oracle):
    nx = ny = 128
    E, K = 640, 8
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, ensemble=E, max_iterations=K))   # AUTO
    one = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K)
    checked = sorted({0, 1, 2, 147, 148, 295, 296, 444, 591, 638, 639} | set(range(17, E, 41)))
    oracles = {e: oracle.OracleCloth(one, oracle.COLOUR) for e in checked}
    for e in range(E):
        st = random_state(nx, ny, seed=7000 + e)
        g.write_state(*st, cloth=e)
        if e in oracles:
            oracles[e].write_state(*st)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 9.5
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 128.0 + 3 * nx, 164.0 + 3 * ny, 120.0 + 3 * nx, 170.0 + 3 * ny
    for t in range(3):
        g.step(f)
        for o in oracles.values():
            o.step(f)
    assert g.last_schedule()[1] == "wave_frame_kernel<double,DRAG,PACK>"
    for e, o in oracles.items():
        assert_same_state(g.read_state(cloth=e), o.read_state(), f"packed ensemble, cloth {e} of {E}")


def test_packed_ensemble_of_640_cloths_with_a_mouse_per_cloth(oracle):
    """BASELINE configs[3] as bench.py runs it (`c3`): clothgpu_step_each on the packed wavefront kernel, AUTO schedule,
    enough cloths for several CTA spans (a span starts in the middle of a cloth, so the entry warps pick up the frame of
    the cloth they land in).  The frames are bench.py's own per-cloth frames (phase = cloth id)."""
    import bench
    nx = ny = 128
    E, K = 640, 8
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, ensemble=E, max_iterations=K))
    one = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K)
    checked = sorted({0, 1, 2, 3, 147, 148, 149, 295, 296, 297, 444, 445, 591, 592, 636, 637, 638, 639} | set(range(16, E, 44)))
    oracles = {e: oracle.OracleCloth(one, oracle.COLOUR) for e in checked}
    for e in range(E):
        st = random_state(nx, ny, seed=9000 + e)
        g.write_state(*st, cloth=e)
        if e in oracles:
            oracles[e].write_state(*st)
    torn = 0
    for t in range(3):
        frames = bench.per_cloth_frames(t, E, nx, ny, K)
        for f in frames:
            f.tear_length = 9.5          # the jittered start state tears under the dragged cloths' mice
        g.step_each(frames)
        for e, o in oracles.items():
            o.step(frames[e])
            torn += o.torn_last
    assert g.last_schedule()[1] == "wave_frame_kernel<double,PER_CLOTH,PACK>"
    for e, o in oracles.items():
        assert_same_state(g.read_state(cloth=e), o.read_state(), f"packed ensemble with per-cloth frames, cloth {e} of {E}")
    assert torn > 0
