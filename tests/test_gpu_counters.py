"""clothgpu_get_counters: the frame kernels keep the counters themselves (warp-reduced, one atomic per warp that saw an
event) and the call is one small copy; a census pass over the flag bytes runs only after create / reset / write_state.
Every value must equal what a census of the read-back state gives, on every schedule, after every frame."""
import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import (FLAG_ACTIVE, FLAG_ALIVE_H, FLAG_ALIVE_V, FLAG_HIGHLIGHT, FLAG_PIN, MOUSE_CTRL, MOUSE_DRAGGING,
                           MOUSE_LEFT, MOUSE_RIGHT, SCHED_AUTO, SCHED_FUSED, SCHED_ROWSTREAM, SCHED_STREAMING, SCHED_WAVE)

from test_gpu_parity import random_state

pytestmark = pytest.mark.gpu


def census(nx, ny, fl):
    """What Cloth.Update's stick loop would visit (physics/cloth.go:97) and the particle flags, from the flag bytes."""
    f = fl.reshape(ny, nx)
    act = (f & FLAG_ACTIVE) != 0
    ah, av = (f & FLAG_ALIVE_H) != 0, (f & FLAG_ALIVE_V) != 0
    live = int((ah[:, :-1] & act[:, :-1] & act[:, 1:]).sum() + (av[:-1] & act[:-1] & act[1:]).sum())
    return {"live_sticks": live, "alive_sticks": int(ah.sum() + av.sum()), "pinned": int(((f & FLAG_PIN) != 0).sum()),
            "inactive": int((~act).sum()), "highlighted": int(((f & FLAG_HIGHLIGHT) != 0).sum())}


def frames_with_events(nx, ny, K):
    """idle, hole punch (right button), ctrl-click pin, drag with tearing, idle again — all near the cloth's centre."""
    cx, cy = 128.0 + 3.0 * nx, 164.0 + 3.0 * ny
    out = []
    for t in range(9):
        f = _abi.default_frame(K)
        f.win_w, f.win_h = 128 + 6 * nx + 60, 164 + 6 * ny + 40
        f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = cx + 7 * t, cy + 3 * t, cx + 7 * t - 6, cy + 3 * t + 4
        f.scroll_y = 40.0
        if t in (1, 2):
            f.buttons = MOUSE_RIGHT
        elif t == 3:
            f.buttons = MOUSE_LEFT | MOUSE_CTRL
            f.scroll_y = 30.0
        elif t in (4, 5, 6):
            f.buttons = MOUSE_LEFT | MOUSE_DRAGGING
            f.mouse_force = 3.0
            f.tear_length = 9.0
        elif t == 7:
            f.mouse_x = f.mouse_y = -1e4      # far away: nothing highlighted
        out.append(f)
    return out


@pytest.mark.parametrize("nx,ny,K,schedule", [(171, 36, 1, SCHED_ROWSTREAM), (300, 219, 1, SCHED_ROWSTREAM), (171, 36, 1, SCHED_FUSED),
                                              (130, 131, 3, SCHED_FUSED), (150, 70, 11, SCHED_FUSED), (300, 300, 4, SCHED_WAVE),
                                              (300, 300, 8, SCHED_WAVE), (129, 33, 2, SCHED_STREAMING), (171, 36, 1, SCHED_AUTO)])
def test_counters_equal_a_census_of_the_state_after_every_frame(nx, ny, K, schedule):
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K))
    g.set_schedule(schedule)
    assert {k: g.counters()[k] for k in census(nx, ny, g.read_state()[4])} == census(nx, ny, g.read_state()[4])
    st = random_state(nx, ny, seed=nx + 3 * K)
    g.write_state(*st)                                    # invalidates the census
    c = g.counters()
    assert {k: c[k] for k in census(nx, ny, st[4])} == census(nx, ny, st[4])
    torn = 0
    for t, f in enumerate(frames_with_events(nx, ny, K)):
        l0 = g.launch_count()
        g.step(f)
        step_launches = g.launch_count() - l0
        l0 = g.launch_count()
        c = g.counters()
        assert g.launch_count() == l0, "get_counters must not launch anything between frames"
        want = census(nx, ny, g.read_state()[4])
        assert {k: c[k] for k in want} == want, f"frame {t}: {c} vs {want}"
        assert c["frames"] == t + 1
        torn += c["torn_last"]
        if schedule != SCHED_STREAMING and K <= 8:
            assert step_launches == 1, "one frame kernel per frame"
    assert c["pinned"] > int(((st[4] & FLAG_PIN) != 0).sum()), "the ctrl-click pins something"
    assert c["inactive"] > int(((st[4] & FLAG_ACTIVE) == 0).sum()), "the right button punches a hole"
    assert torn > 0, "the drag tears sticks"
    fl0 = st[4].reshape(ny, nx).copy()
    fl0[:, -1] &= ~np.uint8(FLAG_ALIVE_H)                 # write_state drops sticks past the last column / row
    fl0[-1, :] &= ~np.uint8(FLAG_ALIVE_V)
    assert c["alive_sticks"] == int(((fl0 & FLAG_ALIVE_H) != 0).sum() + ((fl0 & FLAG_ALIVE_V) != 0).sum()) - torn


def test_counters_of_an_ensemble_sum_its_cloths():
    nx, ny, E, K = 100, 91, 5, 8
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K, ensemble=E))
    for e in range(E):
        g.write_state(*random_state(nx, ny, seed=40 + e), cloth=e)
    for t, f in enumerate(frames_with_events(nx, ny, K)):
        g.step(f)
        c = g.counters()
        want = {}
        for e in range(E):
            for k, v in census(nx, ny, g.read_state(cloth=e)[4]).items():
                want[k] = want.get(k, 0) + v
        assert {k: c[k] for k in want} == want, f"frame {t}"


def test_write_state_keeps_only_the_five_flag_bits():
    nx, ny = 128, 128
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=8, ensemble=4)
    a, b = clothgpu.Cloth(desc), clothgpu.Cloth(desc)
    a.set_schedule(SCHED_WAVE)                        # few cloths: AUTO would keep the tile kernel
    b.set_schedule(SCHED_STREAMING)
    for e in range(4):
        st = list(random_state(nx, ny, seed=e))
        dirty = st[4] | np.uint8(0xE0)                # stray high bits must not become sticks on any schedule
        a.write_state(*st[:4], dirty, cloth=e)
        b.write_state(*st[:4], st[4], cloth=e)
    f = _abi.default_frame(8)
    f.win_w, f.win_h = 2000, 2000
    for _ in range(3):
        a.step(f)
        b.step(f)
    assert a.last_schedule()[1].startswith("wave_frame_kernel<double,PACK"), a.last_schedule()
    for e in range(4):
        for p, q in zip(a.read_state(cloth=e), b.read_state(cloth=e)):
            assert np.array_equal(p, q)
        assert int(a.read_state(cloth=e)[4].max()) < 0x20


def test_last_schedule_reports_what_ran():
    g = clothgpu.Cloth(_abi.grid_desc(600, 600, max_iterations=8))
    with pytest.raises(clothgpu.ClothGpuError):
        g.last_schedule()
    g.step(_abi.default_frame(1))
    assert g.last_schedule() == (SCHED_ROWSTREAM, "stream1_frame_kernel<double>")
    f = _abi.default_frame(8)
    f.buttons = MOUSE_DRAGGING
    g.step(f)
    assert g.last_schedule() == (SCHED_WAVE, "wave_frame_kernel<double,DRAG>")
    g.step(_abi.default_frame(2))
    assert g.last_schedule() == (SCHED_FUSED, "fused_frame_kernel<double>")
    small = clothgpu.Cloth(_abi.grid_desc(20, 12, max_iterations=1))
    small.step(_abi.default_frame(1))
    assert small.last_schedule()[0] == SCHED_FUSED


def test_reset_of_attached_strips_is_a_strip_operation():
    """clothgpu_reset on attached strips hand-shakes like a frame: the neighbours' last stores into my ghost rows land
    before the init kernel rewrites them, and frames after the reset still equal the whole cloth."""
    from clothgpu.strips import StripPlan
    nx, ny, K = 150, 132, 1
    plan = StripPlan(ny, 2, ghost=2)
    whole = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K))
    strips = []
    for r in range(2):
        row0, rows = plan.rows(r)
        strips.append(clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K, strip_row0=row0, strip_rows=rows)))
    blobs = [s.strip_export() for s in strips]
    strips[0].strip_attach(None, blobs[1])
    strips[1].strip_attach(blobs[0], None)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 3000, 3000
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 7.0
    b1 = plan.bounds[1]
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 500.0, 164.0 + 6 * b1, 470.0, 150.0 + 6 * b1

    def compare(what):
        parts = [np.concatenate(p) for p in zip(*[s.read_state() for s in strips])]
        for n, a, b in zip(("x", "y", "px", "py", "flags"), parts, whole.read_state()):
            assert np.array_equal(a, b), f"{what}: plane {n}"

    for round_ in range(3):
        for t in range(7):
            whole.step(f)
            for s in strips:
                s.step(f)
        compare(f"round {round_} before reset")
        whole.reset(128 + round_, 164)                    # no sync in between: the reset queues behind the frames
        for s in strips:
            s.reset(128 + round_, 164)
        compare(f"round {round_} after reset")
        cs, cw = [s.counters() for s in strips], whole.counters()
        for k in ("live_sticks", "alive_sticks", "pinned", "inactive", "frames"):
            assert (cs[0][k] if k == "frames" else cs[0][k] + cs[1][k]) == cw[k], k
