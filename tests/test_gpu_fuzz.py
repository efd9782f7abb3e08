"""Differential fuzz: random shapes, sweep counts, ensembles, precisions, states and frames (random buttons, mouse inside
the cloth, small tear length) — every schedule that can run the case must leave the same bits as the one-launch-per-colour
streaming schedule, which the oracle pins on the fixed shapes of test_gpu_parity.py.  Seeds are fixed: a failure names its case."""
import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import (MOUSE_CTRL, MOUSE_DRAGGING, MOUSE_LEFT, MOUSE_RIGHT, SCHED_AUTO, SCHED_FUSED, SCHED_RESIDENT,
                           SCHED_ROWSTREAM, SCHED_STREAMING, SCHED_WAVE)

from test_gpu_parity import assert_same_state, random_state

pytestmark = pytest.mark.gpu


def random_case(seed):
    rng = np.random.default_rng(seed)
    kind = seed % 4
    if kind == 0:      # GUI-sized
        nx, ny = int(rng.integers(2, 200)), int(rng.integers(2, 60))
    elif kind == 1:    # wide and short / narrow and tall
        nx, ny = (int(rng.integers(300, 1400)), int(rng.integers(4, 40))) if seed % 8 == 1 else (int(rng.integers(2, 40)), int(rng.integers(300, 1500)))
    elif kind == 2:    # several windows / bands / chunks
        nx, ny = int(rng.integers(200, 900)), int(rng.integers(200, 700))
    else:              # ensembles of small cloths
        nx, ny = int(rng.integers(2, 129)), int(rng.integers(2, 140))
    ens = int(rng.integers(2, 7)) if kind == 3 else 1
    K = int(rng.integers(1, 10))
    prec = _abi.F32 if seed % 5 == 0 else _abi.F64
    return nx, ny, ens, K, prec, rng


def random_frames(rng, nx, ny, K, n):
    out = []
    for _ in range(n):
        f = _abi.default_frame(K)
        f.win_w, f.win_h = 128 + 6 * nx + int(rng.integers(-20, 60)), 164 + 6 * ny + int(rng.integers(-20, 40))   # may clip the cloth
        f.mouse_x, f.mouse_y = 128.0 + 6 * nx * rng.random(), 164.0 + 6 * ny * rng.random()
        f.mouse_px, f.mouse_py = f.mouse_x + rng.uniform(-40, 40), f.mouse_y + rng.uniform(-40, 40)
        f.buttons = int(rng.choice([0, MOUSE_LEFT | MOUSE_DRAGGING, MOUSE_RIGHT, MOUSE_LEFT | MOUSE_CTRL, MOUSE_LEFT | MOUSE_DRAGGING | MOUSE_CTRL]))
        f.mouse_force = float(rng.uniform(0, 12))
        f.scroll_y = float(rng.uniform(20, 130))
        f.tear_length = float(rng.choice([150.0, 9.0, 7.5]))
        f.win_off_x, f.win_off_y = float(rng.choice([0.0, 3.0])), float(rng.choice([0.0, -2.0]))
        out.append(f)
    return out


@pytest.mark.parametrize("seed", range(40))
def test_all_schedules_agree_on_random_cases(seed):
    nx, ny, ens, K, prec, rng = random_case(seed)
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW if seed % 3 else _abi.PIN_NONE, max_iterations=K, ensemble=ens, precision=prec)
    states = [random_state(nx, ny, seed=1000 * seed + e) for e in range(ens)]
    if prec == _abi.F32:
        states = [tuple(a.astype(np.float32).astype(np.float64) if a.dtype == np.float64 else a for a in st) for st in states]
    frames = random_frames(rng, nx, ny, K, 3)
    schedules = [SCHED_STREAMING, SCHED_AUTO, SCHED_FUSED]
    if K == 1:
        schedules.append(SCHED_ROWSTREAM)
    if 2 <= K <= 8:
        schedules.append(SCHED_WAVE)
    if ens == 1:
        schedules.append(SCHED_RESIDENT)          # falls back to AUTO when the cloth does not fit one CTA
    results, counters = {}, {}
    for sched in schedules:
        g = clothgpu.Cloth(desc)
        g.set_schedule(sched)
        for e, st in enumerate(states):
            g.write_state(*st, cloth=e)
        if sched == SCHED_RESIDENT:
            g.step_n(frames)
        else:
            for f in frames:
                g.step(f)
        results[sched] = [g.read_state(cloth=e) for e in range(ens)]
        c = g.counters()
        counters[sched] = {k: c[k] for k in ("live_sticks", "alive_sticks", "torn_last", "pinned", "inactive", "highlighted", "relaxations")}
        g.close()
    ref = results[SCHED_STREAMING]
    for sched in schedules[1:]:
        for e in range(ens):
            assert_same_state(results[sched][e], ref[e], f"seed {seed}: {nx}x{ny} x{ens} K={K} prec={prec} schedule {sched} cloth {e}")
        assert counters[sched] == counters[SCHED_STREAMING], f"seed {seed}: counters of schedule {sched}"
