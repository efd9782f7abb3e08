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


@pytest.mark.parametrize("seed", range(32))
def test_live_stick_export_on_random_cases(seed):
    """Random shapes (1 particle .. several hundred tiles), ensembles, precisions, two strips, and densities of dead sticks,
    holes and highlights from none to all: segments, highlight bytes, indices and outline quads against numpy /
    the copy path, through pageable buffers and through pinned ones (written by the kernel itself when small enough)."""
    import torch
    from test_gpu_parity import _exchange_halos_same_device, _expected_segments_fast
    from clothgpu._abi import FLAG_HIGHLIGHT
    rng = np.random.default_rng(7000 + seed)
    kind = seed % 4
    nx, ny = [(int(rng.integers(1, 40)), int(rng.integers(1, 40))), (int(rng.integers(1, 300)), int(rng.integers(1, 200))),
              (int(rng.integers(200, 1500)), int(rng.integers(100, 900))), (int(rng.integers(1, 9)), int(rng.integers(500, 3000)))][kind]
    ens = int(rng.integers(1, 4)) if kind < 2 else 1
    prec = _abi.F32 if seed % 5 == 0 else _abi.F64
    kill, holes, hot = [float(rng.choice([0.0, 0.02, 0.3, 1.0])) for _ in range(3)]
    e = int(rng.integers(0, ens))
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, ensemble=ens, precision=prec))
    for k in range(ens):
        st = list(random_state(nx, ny, seed=90 * seed + k, kill=kill, holes=holes))
        st[4][rng.random(st[4].size) < hot] |= FLAG_HIGHLIGHT
        if prec == _abi.F32:
            st[:4] = [a.astype(np.float32).astype(np.float64) for a in st[:4]]
        g.write_state(*st, cloth=k)
        if k == e:
            mine = st
    x, y, _, _, fl = g.read_state(cloth=e)
    want, want_hl = _expected_segments_fast(nx, ny, x, y, fl)
    got, got_hl, idx = g.read_segments(cloth=e, with_indices=True)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32)), (nx, ny, ens)
    assert np.array_equal(got_hl, want_hl)
    n = len(want)
    if n:
        d = idx[:, 1].astype(np.int64) - idx[:, 0]
        assert np.all((d == 1) | (d == nx))
    quads, qhl = g.read_quads(cloth=e, line_width=0.6)
    assert len(quads) == n and np.array_equal(qhl, want_hl)
    cap = max(1, 2 * nx * ny)
    pq, ph = torch.full((cap, 8), 5.0).pin_memory(), torch.full((cap,), 9, dtype=torch.uint8).pin_memory()
    assert g.read_quads_into(pq.data_ptr(), ph.data_ptr(), cap, cloth=e) == n
    assert np.array_equal(pq.numpy()[:n].view(np.uint32), quads.view(np.uint32)) and np.array_equal(ph.numpy()[:n], want_hl)
    ps = torch.full((cap, 4), 5.0).pin_memory()
    assert g.read_segments_into(ps.data_ptr(), ph.data_ptr(), 0, cap, cloth=e) == n
    assert np.array_equal(ps.numpy()[:n].view(np.uint32), want.view(np.uint32))
    if seed % 3 == 0 and ny > 3 and ens == 1:   # the same cloth as two strips: their lists concatenate to the whole one
        split = max(2, int(rng.integers(2, ny)) & ~1)   # strips start on an even row
        up = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=1, precision=prec, strip_row0=0, strip_rows=split))
        lo = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=1, precision=prec, strip_row0=split, strip_rows=ny - split))
        up.write_state(*[a.reshape(ny, nx)[:split].ravel() for a in mine])
        lo.write_state(*[a.reshape(ny, nx)[split:].ravel() for a in mine])
        _exchange_halos_same_device(up, lo)
        su, hu = up.read_segments()
        sl, hl = lo.read_segments()
        assert len(su) + len(sl) == n, (nx, ny, split)
        assert np.array_equal(np.concatenate([su, sl]).view(np.uint32), want.view(np.uint32)) and np.array_equal(np.concatenate([hu, hl]), want_hl)
