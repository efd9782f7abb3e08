"""GPU <-> oracle parity through the C ABI (bit-exact in binary64; the reference's own solver cannot run
here, see oracle/cloth_oracle.h "PARITY UNPINNED")."""
import ctypes as C

import os

import numpy as np
import pytest

import clothgpu
from clothgpu import _abi
from clothgpu._abi import (FLAG_ACTIVE, FLAG_ALIVE_H, FLAG_ALIVE_V, FLAG_HIGHLIGHT, FLAG_PIN, MOUSE_DRAGGING,
                           SCHED_FUSED, SCHED_ROWSTREAM, SCHED_STREAMING, SCHED_WAVE)
from traces import c0_trace

pytestmark = pytest.mark.gpu
NAMES = ("x", "y", "px", "py", "flags")


def assert_same_state(gpu_state, ora_state, what=""):
    for n, a, b in zip(NAMES, gpu_state, ora_state):
        if not np.array_equal(a, b):
            bad = np.flatnonzero(a != b)
            raise AssertionError(f"{what}: plane {n} differs at {bad.size} of {a.size} particles, first {bad[:5]}: "
                                 f"gpu {a[bad[:5]]} oracle {b[bad[:5]]}")


def random_state(nx, ny, seed, spacing=6.0, amp=3.0, origin=(128.0, 164.0), kill=0.02, pins=0.01, holes=0.01):
    """Jittered grid: a mix of stretched and compressed sticks, a few dead sticks, pins and holes."""
    rng = np.random.default_rng(seed)
    gx, gy = np.meshgrid(np.arange(nx) * spacing + origin[0], np.arange(ny) * spacing + origin[1])
    x = (gx + rng.uniform(-amp, amp, gx.shape)).ravel()
    y = (gy + rng.uniform(-amp, amp, gy.shape)).ravel()
    px = x + rng.uniform(-0.5, 0.5, x.shape)
    py = y + rng.uniform(-0.5, 0.5, y.shape)
    fl = np.full(nx * ny, FLAG_ACTIVE | FLAG_ALIVE_H | FLAG_ALIVE_V, np.uint8)
    fl2 = fl.reshape(ny, nx)
    fl2[:, -1] &= ~np.uint8(FLAG_ALIVE_H)
    fl2[-1, :] &= ~np.uint8(FLAG_ALIVE_V)
    fl[rng.random(fl.size) < kill] &= ~np.uint8(FLAG_ALIVE_H)
    fl[rng.random(fl.size) < kill] &= ~np.uint8(FLAG_ALIVE_V)
    fl[rng.random(fl.size) < pins] |= FLAG_PIN
    fl[rng.random(fl.size) < holes] &= ~np.uint8(FLAG_ACTIVE)
    return x, y, px, py, fl


def make_pair(oracle, desc, schedule):
    g = clothgpu.Cloth(desc)
    g.set_schedule(schedule)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    return g, o


@pytest.mark.parametrize("schedule", [SCHED_STREAMING, SCHED_FUSED, SCHED_ROWSTREAM])
def test_default_cloth_scripted_trace_bit_exact(oracle, schedule):
    """BASELINE configs[0]: the reference's default GUI cloth, 1000 frames, drag / hole / tear / ctrl-pin."""
    desc = _abi.default_desc()
    g, o = make_pair(oracle, desc, schedule)
    frames = c0_trace(1000)
    torn_total = 0
    for t, f in enumerate(frames):
        if t == 800:   # aim the ctrl-click at where a free particle is right now
            xs, ys = o.read_state()[:2]
            f.mouse_x, f.mouse_y = float(np.float32(xs[10 * 171 + 40])), float(np.float32(ys[10 * 171 + 40]))
        g.step(f)
        o.step(f)
        torn_total += o.torn_last
        if t % 50 == 49 or t in (100, 101, 450, 451, 600, 601, 640, 800, 801):
            assert_same_state(g.read_state(), o.read_state(), f"frame {t}")
            c = g.counters()
            assert c["torn_last"] == o.torn_last, t
            assert c["relaxations"] == o.relaxations, t
    c = g.counters()
    fl = o.read_state()[4]
    assert torn_total > 10, "the trace is meant to tear sticks"
    assert c["pinned"] == int(((fl & FLAG_PIN) != 0).sum()) > 11        # 11 Init pins + the ctrl-click(s)
    assert c["inactive"] == int(((fl & FLAG_ACTIVE) == 0).sum()) > 0
    assert c["alive_sticks"] == int(((fl & FLAG_ALIVE_H) != 0).sum() + ((fl & FLAG_ALIVE_V) != 0).sum())
    assert c["frames"] == 1000


SHAPES = [(11, 2), (2, 11), (12, 1), (1, 12), (64, 64), (128, 40), (129, 33), (130, 131), (171, 36), (257, 100),
          (300, 219)]


@pytest.mark.parametrize("nx,ny", SHAPES)
@pytest.mark.parametrize("K", [1, 2, 3, 8, 11])
def test_random_states_bit_exact(oracle, nx, ny, K):
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K)
    st = random_state(nx, ny, seed=nx * 1000 + ny * 7 + K)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20        # some particles hit the bounds clamp
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 9.5                                           # low threshold: many tears
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 128.0 + 3 * nx, 164.0 + 3 * ny, 120.0 + 3 * nx, 170.0 + 3 * ny
    results = []
    for schedule in (SCHED_STREAMING, SCHED_FUSED) + ((SCHED_ROWSTREAM,) if K == 1 else ()) + ((SCHED_WAVE,) if 2 <= K <= 8 else ()):
        g, o = make_pair(oracle, desc, schedule)
        g.write_state(*st)
        o.write_state(*st)
        for t in range(3):
            g.step(f)
            o.step(f)
            assert_same_state(g.read_state(), o.read_state(), f"{nx}x{ny} K={K} sched={schedule} frame {t}")
            c = g.counters()
            assert (c["torn_last"], c["relaxations"]) == (o.torn_last, o.relaxations)
        results.append(g.read_state())
        g.close()
    assert_same_state(results[0], results[1], "streaming vs fused")


def test_reset_restores_init_state(oracle):
    desc = _abi.default_desc()
    g, o = make_pair(oracle, desc, SCHED_FUSED)
    for f in c0_trace(130):
        g.step(f)
    g.reset(100, 90)
    o.reset(100, 90)
    assert_same_state(g.read_state(), o.read_state(), "after reset")
    assert g.counters()["frames"] == 0


def test_create_rejects_what_the_reference_panics_on():
    d = _abi.grid_desc(9, 9, pin_rule=_abi.PIN_REFERENCE)
    with pytest.raises(clothgpu.ClothGpuError) as e:
        clothgpu.Cloth(d)
    assert e.value.code == _abi.ERR_INVALID and "cloth.go:72" in str(e.value)


def test_masks_match_flag_bytes():
    desc = _abi.grid_desc(171, 36, pin_rule=_abi.PIN_REFERENCE)
    g = clothgpu.Cloth(desc)
    st = random_state(171, 36, 5)
    g.write_state(*st)
    fl = g.read_state()[4]
    m = g.read_masks()
    for name, bit in (("pin", FLAG_PIN), ("active", FLAG_ACTIVE), ("highlighted", FLAG_HIGHLIGHT),
                      ("alive_h", FLAG_ALIVE_H), ("alive_v", FLAG_ALIVE_V)):
        bits = np.unpackbits(m[name].view(np.uint8), bitorder="little")[: fl.size]
        assert np.array_equal(bits.astype(bool), (fl & bit) != 0), name


def _expected_segments(nx, ny, x, y, fl):
    """The render loop of physics/cloth.go:108-114 over the creation-ordered slice."""
    segs, hl = [], []
    act = (fl & FLAG_ACTIVE) != 0
    hi = (fl & FLAG_HIGHLIGHT) != 0
    for r in range(ny):
        for c in range(nx):
            p = r * nx + c
            for q, bit in ((p - nx, FLAG_ALIVE_V), (p - 1, FLAG_ALIVE_H)):
                if (bit == FLAG_ALIVE_V and r == 0) or (bit == FLAG_ALIVE_H and c == 0):
                    continue
                if (fl[q] & bit) and act[q] and act[p]:
                    segs.append((np.float32(x[q]), np.float32(y[q]), np.float32(x[p]), np.float32(y[p])))
                    hl.append(hi[q] and hi[p])
    return np.array(segs, np.float32).reshape(-1, 4), np.array(hl, np.uint8)


def test_live_stick_list_matches_reference_render_order(oracle):
    desc = _abi.default_desc()
    g, o = make_pair(oracle, desc, SCHED_FUSED)
    for t, f in enumerate(c0_trace(700)):
        g.step(f)
        o.step(f)
    x, y, px, py, fl = o.read_state()
    want, want_hl = _expected_segments(g.nx, g.ny, x, y, fl)
    got, got_hl, idx = g.read_segments(with_indices=True)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(got_hl, want_hl)
    assert g.counters()["live_sticks"] == len(want)
    assert np.all(idx[:, 1] > idx[:, 0])


def _expected_segments_fast(nx, ny, x, y, fl):
    """_expected_segments with numpy (creation order: by tail particle, the vertical stick before the horizontal one)."""
    x32, y32 = x.astype(np.float32).reshape(ny, nx), y.astype(np.float32).reshape(ny, nx)
    f2 = fl.reshape(ny, nx)
    act, hi = (f2 & FLAG_ACTIVE) != 0, (f2 & FLAG_HIGHLIGHT) != 0
    v = np.zeros((ny, nx), bool)
    h = np.zeros((ny, nx), bool)
    v[1:] = ((f2[:-1] & FLAG_ALIVE_V) != 0) & act[:-1] & act[1:]          # stick from the particle above
    h[:, 1:] = ((f2[:, :-1] & FLAG_ALIVE_H) != 0) & act[:, :-1] & act[:, 1:]  # ... from the particle to the left
    segs, hl = [], []
    order = np.zeros((ny, nx, 2), bool)
    order[..., 0], order[..., 1] = v, h
    rr, cc, kind = np.nonzero(order)      # row-major over (row, column, [vertical, horizontal]) = creation order
    hr, hc = rr - (kind == 0), cc - (kind == 1)
    segs = np.stack([x32[hr, hc], y32[hr, hc], x32[rr, cc], y32[rr, cc]], axis=1)
    hl = (hi[hr, hc] & hi[rr, cc]).astype(np.uint8)
    return segs.astype(np.float32).reshape(-1, 4), hl


@pytest.mark.parametrize("nx,ny,ens", [(171, 36, 1), (700, 90, 2), (1037, 1100, 1), (16, 300, 3), (2, 9, 1), (4100, 64, 1)])
def test_live_stick_list_on_random_states(nx, ny, ens):
    """The compaction kernels on shapes whose rows are shorter / longer than a warp's 512 particles, odd widths (row
    padding), several scatter blocks, an ensemble member other than 0: random dead sticks, holes and highlights."""
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, ensemble=ens)
    g = clothgpu.Cloth(desc)
    rng = np.random.default_rng(nx + ny)
    for e in range(ens):
        x, y, px, py, fl = random_state(nx, ny, seed=nx * 3 + e, kill=0.1, holes=0.05)
        fl[rng.random(fl.size) < 0.3] |= FLAG_HIGHLIGHT
        g.write_state(x, y, px, py, fl, cloth=e)
    e = ens - 1
    x, y, px, py, fl = g.read_state(cloth=e)
    want, want_hl = _expected_segments_fast(nx, ny, x, y, fl)
    got, got_hl, idx = g.read_segments(cloth=e, with_indices=True)
    assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
    assert np.array_equal(got_hl, want_hl)
    if len(idx):
        d = idx[:, 1].astype(np.int64) - idx[:, 0]
        assert np.all((d == 1) | (d == nx)) and np.all(np.diff(idx[:, 1].astype(np.int64)) >= 0)
    if nx <= 200:  # the slow reference loop agrees with the vectorised one
        w2, h2 = _expected_segments(nx, ny, x, y, fl)
        assert np.array_equal(w2, want) and np.array_equal(h2, want_hl)


@pytest.mark.parametrize("nx,ny", [(1, 1), (1, 50), (5000, 1), (64, 64), (128, 128), (128, 129), (4096, 4100)])
def test_live_stick_list_at_the_tile_boundaries_of_the_export(nx, ny):
    """The export works on tiles of 4096 padded particles: a single particle, one column (15 of 16 padded columns are
    empty), one row, exactly 1 and exactly 4 tiles (the largest cloth whose export blocks plan for themselves, ONE launch),
    5 tiles with a ragged last one (the smallest with a plan launch), and more than 4096 tiles (the last plan block scans
    the tile totals in more than one step).  Highlights only in a few places, then nowhere: most tiles write their
    highlight bytes as plain zeros, and a second export must not leave the first one's bytes behind."""
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE))
    x, y, px, py, fl = random_state(nx, ny, seed=nx + 7 * ny, kill=0.02, holes=0.01)
    rng = np.random.default_rng(nx * 31 + ny)
    fl2 = fl.copy()
    hot = rng.integers(0, fl.size, size=max(1, fl.size // 5000))   # a handful of highlighted spots ...
    for d in (0, 1, nx):                                           # ... each with its right and lower neighbour
        fl2[np.minimum(hot + d, fl.size - 1)] |= FLAG_HIGHLIGHT
    for flags in (fl2, fl & ~np.uint8(FLAG_HIGHLIGHT)):
        g.write_state(x, y, px, py, flags)
        st = g.read_state()
        want, want_hl = _expected_segments_fast(nx, ny, st[0], st[1], st[4])
        got, got_hl = g.read_segments()
        assert got.shape == want.shape and np.array_equal(got.view(np.uint32), want.view(np.uint32))
        assert np.array_equal(got_hl, want_hl)
        assert g.count_segments() == len(want)
    assert not want_hl.any()


def test_live_stick_lists_of_two_strips_concatenate_to_the_whole_cloth():
    """A strip lists the sticks whose TAIL it owns; the vertical sticks of its first row hang from the ghost row above."""
    nx, ny, split = 333, 80, 34
    whole = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=1)
    up = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=1, strip_row0=0, strip_rows=split)
    lo = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=1, strip_row0=split, strip_rows=ny - split)
    g, gu, gl = clothgpu.Cloth(whole), clothgpu.Cloth(up), clothgpu.Cloth(lo)
    st = random_state(nx, ny, seed=9, kill=0.1, holes=0.05)
    g.write_state(*st)
    gu.write_state(*[a.reshape(ny, nx)[:split].ravel() for a in st])
    gl.write_state(*[a.reshape(ny, nx)[split:].ravel() for a in st])
    _exchange_halos_same_device(gu, gl)
    sw, hw = g.read_segments()
    su, hu = gu.read_segments()
    sl, hl = gl.read_segments()
    assert len(su) + len(sl) == len(sw)
    assert np.array_equal(np.concatenate([su, sl]), sw) and np.array_equal(np.concatenate([hu, hl]), hw)


def test_outline_quads_match_addsegment_of_the_reference(oracle):
    """clothgpu_read_quads_f32 = addSegment / normal (physics/cloth.go:150-168) over the live-stick list: bit-exact float32
    against the oracle's restatement on the state the scripted trace leaves (torn, punched, highlighted)."""
    desc = _abi.default_desc()
    g, o = make_pair(oracle, desc, SCHED_ROWSTREAM)
    for t, f in enumerate(c0_trace(700)):
        g.step(f)
        o.step(f)
    segs, seg_hl = g.read_segments()
    quads, hl = g.read_quads(line_width=0.6)
    assert quads.shape == (len(segs), 8) and np.array_equal(hl, seg_hl)
    want = oracle.segment_quads(segs, 0.6)
    assert np.array_equal(quads.view(np.uint32), want.view(np.uint32))
    # a degenerate stick (both ends on the same spot) gets a zero normal, as normal() returns f32.Point{}
    x, y, px, py, fl = g.read_state()
    x[1], y[1] = x[0], y[0]
    g.write_state(x, y, px, py, fl)
    q2, _ = g.read_quads()
    s2, _ = g.read_segments()
    assert np.array_equal(q2.view(np.uint32), oracle.segment_quads(s2, 0.6).view(np.uint32))
    k = int(np.flatnonzero((s2[:, 0] == s2[:, 2]) & (s2[:, 1] == s2[:, 3]))[0])
    assert np.array_equal(q2[k], np.tile(s2[k, :2], 4))


@pytest.mark.parametrize("nx,ny", [(171, 36), (150, 100), (2, 9)])
def test_export_written_straight_into_pinned_host_memory(nx, ny):
    """GUI-sized lists: when the caller's buffers are pinned host memory the export kernel stores into them itself (no
    device buffer, no copies).  Same bytes as the copy path that pageable buffers take; buffers that do not qualify
    (misaligned, one of them pageable) fall back silently; a list longer than the capacity is an error, and nothing is
    written behind the capacity."""
    import torch
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE))
    x, y, px, py, fl = random_state(nx, ny, seed=nx * ny, kill=0.05, holes=0.02)
    fl[np.random.default_rng(nx).random(fl.size) < 0.2] |= FLAG_HIGHLIGHT
    g.write_state(x, y, px, py, fl)
    segs, hl, idx = g.read_segments(with_indices=True)      # pageable numpy buffers: the copy path
    quads, qhl = g.read_quads(line_width=0.6)
    n, cap = len(segs), 2 * nx * ny
    pin = lambda shape, dt: torch.full(shape, 77, dtype=dt).pin_memory()
    ps, ph, pi, pq = pin((cap + 2, 4), torch.float32), pin((cap + 32,), torch.uint8), pin((cap, 2), torch.int32), pin((cap + 2, 8), torch.float32)
    assert g.read_segments_into(ps.data_ptr(), ph.data_ptr(), pi.data_ptr(), cap) == n
    assert np.array_equal(ps.numpy()[:n].view(np.uint32), segs.view(np.uint32)) and np.array_equal(ph.numpy()[:n], hl)
    assert np.array_equal(pi.numpy()[:n].view(np.uint32), idx.view(np.uint32))
    assert (ps.numpy()[n:] == 77).all() and (ph.numpy()[n:] == 77).all()      # nothing behind the list
    assert g.read_quads_into(pq.data_ptr(), ph.data_ptr(), cap) == n
    assert np.array_equal(pq.numpy()[:n].view(np.uint32), quads.view(np.uint32)) and np.array_equal(ph.numpy()[:n], qhl)
    # without the highlight bytes / the indices
    ps.fill_(77)
    assert g.read_segments_into(ps.data_ptr(), 0, 0, cap) == n
    assert np.array_equal(ps.numpy()[:n].view(np.uint32), segs.view(np.uint32))
    # a destination that is not 32-byte aligned, a highlight buffer that is not 16-byte aligned: the copy path, same result
    ps.fill_(77), ph.fill_(77)
    assert g.read_segments_into(ps.data_ptr() + 16, ph.data_ptr() + 3, 0, cap) == n
    assert np.array_equal(ps.numpy().ravel()[4:4 + 4 * n].view(np.uint32), segs.ravel().view(np.uint32))
    assert np.array_equal(ph.numpy()[3:3 + n], hl)
    # capacity below the list's length
    if n > 40:
        ps.fill_(77), ph.fill_(77)
        small = n - 33
        with pytest.raises(clothgpu.ClothGpuError):
            g.read_segments_into(ps.data_ptr(), ph.data_ptr(), 0, small)
        assert (ps.numpy()[small:] == 77).all() and (ph.numpy()[small:] == 77).all()
        assert np.array_equal(ps.numpy()[:small].view(np.uint32), segs[:small].view(np.uint32))


def test_host_buffers_of_the_library_take_the_direct_export(oracle):
    """clothgpu_host_alloc / clothgpu_host_free (what the cgo shim stages its render lists in): the scripted trace's quads
    through such buffers frame by frame equal the copy path's, and the allocator rejects what it did not allocate."""
    desc = _abi.default_desc()
    g, o = make_pair(oracle, desc, _abi.SCHED_AUTO)
    cap = 2 * 171 * 36
    q, h = clothgpu.HostBuffer((cap, 8), np.float32), clothgpu.HostBuffer((cap,), np.uint8)
    assert q.ptr % 32 == 0 and h.ptr % 32 == 0
    for t, f in enumerate(c0_trace(460)):
        g.step(f)
        if t % 23 == 0 or t > 440:
            n = g.read_quads_into(q.ptr, h.ptr, cap)
            want, want_hl = g.read_quads()
            assert n == len(want) and np.array_equal(q.array[:n].view(np.uint32), want.view(np.uint32)) and np.array_equal(h.array[:n], want_hl)
    segs, _ = g.read_segments()
    assert np.array_equal(q.array[:n].view(np.uint32), oracle.segment_quads(segs, 0.6).view(np.uint32))
    q.close(), h.close()
    q.close()                                            # (idempotent)
    with pytest.raises(clothgpu.ClothGpuError):
        clothgpu._check(clothgpu.lib().clothgpu_host_free(np.zeros(8).ctypes.data))


@pytest.mark.parametrize("nx,ny", [(700, 90), (1037, 300)])
def test_outline_quads_of_a_cloth_of_many_tiles(oracle, nx, ny):
    """The QUADS instantiation of the export kernel behind a plan launch (more than 4 tiles): random dead sticks, holes and
    highlights, so both the direct-store and the ranked path run."""
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE))
    x, y, px, py, fl = random_state(nx, ny, seed=nx - ny, kill=0.05, holes=0.02)
    fl[np.random.default_rng(ny).random(fl.size) < 0.1] |= FLAG_HIGHLIGHT
    g.write_state(x, y, px, py, fl)
    segs, seg_hl = g.read_segments()
    quads, hl = g.read_quads(line_width=0.6)
    assert quads.shape == (len(segs), 8) and np.array_equal(hl, seg_hl)
    assert np.array_equal(quads.view(np.uint32), oracle.segment_quads(segs, 0.6).view(np.uint32))


@pytest.mark.parametrize("K,schedule", [(8, SCHED_FUSED), (8, SCHED_WAVE), (1, SCHED_FUSED), (1, SCHED_ROWSTREAM)])
def test_ensemble_cloths_are_independent(oracle, K, schedule):
    nx, ny, E = 128, 128, 5
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, ensemble=E, max_iterations=K)
    g = clothgpu.Cloth(desc)
    g.set_schedule(schedule)
    one = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW)
    oracles = [oracle.OracleCloth(one, oracle.COLOUR) for _ in range(E)]
    for e in range(E):
        st = random_state(nx, ny, seed=100 + e, pins=0.0)
        g.write_state(*st, cloth=e)
        oracles[e].write_state(*st)
    frames = []
    for e in range(E):
        f = _abi.default_frame(K)
        f.win_w, f.win_h = 2000, 2000
        f.buttons = MOUSE_DRAGGING if e % 2 else 0
        f.mouse_x, f.mouse_y = 128.0 + 40 * e, 300.0
        f.mouse_px, f.mouse_py = f.mouse_x - 3, f.mouse_y + 2
        f.tear_length = 10.0
        frames.append(f)
    for t in range(3):
        g.step_each(frames)
        for e in range(E):
            oracles[e].step(frames[e])
    for e in range(E):
        assert_same_state(g.read_state(cloth=e), oracles[e].read_state(), f"cloth {e}")
    # same frame for everybody
    g.step(frames[1])
    for e in range(E):
        oracles[e].step(frames[1])
        assert_same_state(g.read_state(cloth=e), oracles[e].read_state(), f"cloth {e} shared frame")


@pytest.mark.parametrize("nx,ny,E,K", [(128, 128, 6, 8), (100, 91, 7, 8), (64, 40, 9, 4), (17, 300, 3, 5), (128, 130, 2, 6)])
def test_packed_ensembles_on_the_wavefront_kernel(oracle, nx, ny, E, K):
    """Ensembles of narrow cloths run on the wavefront kernel as one tall virtual cloth: two cloths side by side in the
    256-column band, the pairs stacked (WavePlan::pack).  Odd member counts (an empty right half), odd heights (a padding
    row per cloth), widths that are not a multiple of 16, tears and holes: every member must equal the oracle run on that
    cloth alone, bit for bit, and the flag bytes of the cloth edges must survive the trip."""
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, ensemble=E, max_iterations=K)
    g = clothgpu.Cloth(desc)
    g.set_schedule(SCHED_WAVE)
    one = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K)
    oracles = [oracle.OracleCloth(one, oracle.COLOUR) for _ in range(E)]
    for e in range(E):
        st = random_state(nx, ny, seed=1000 + 17 * e + nx)
        g.write_state(*st, cloth=e)
        oracles[e].write_state(*st)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 9.5
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 128.0 + 3 * nx, 164.0 + 3 * ny, 120.0 + 3 * nx, 170.0 + 3 * ny
    relax = 0
    for t in range(3):
        g.step(f)
        for e in range(E):
            oracles[e].step(f)
    for e in range(E):
        assert_same_state(g.read_state(cloth=e), oracles[e].read_state(), f"{nx}x{ny} K={K} cloth {e} of {E}")
        relax += oracles[e].relaxations
    assert g.counters()["relaxations"] == relax


def _cudart():
    import glob
    import os
    for pat in ("/usr/local/cuda/lib64/libcudart.so*", "/usr/local/cuda/targets/x86_64-linux/lib/libcudart.so*"):
        hits = sorted(glob.glob(pat))
        if hits:
            return C.CDLL(hits[0])
    return C.CDLL("libcudart.so")


def _exchange_halos_same_device(upper, lower):
    """Device-to-device halo copy between two strip handles living on one GPU (test stand-in for the
    NVLink exchange): upper's bottom edge -> lower's top ghost, lower's top edge -> upper's bottom ghost."""
    rt = _cudart()
    rt.cudaMemcpy.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int]
    upper.sync()
    lower.sync()
    for plane in range(5):
        src, n = upper.halo_region(plane, 1)
        dst, m = lower.halo_region(plane, 2)
        assert n == m and n > 0
        assert rt.cudaMemcpy(dst, src, n, 3) == 0
        src, n = lower.halo_region(plane, 0)
        dst, m = upper.halo_region(plane, 3)
        assert n == m and n > 0
        assert rt.cudaMemcpy(dst, src, n, 3) == 0


@pytest.mark.parametrize("schedule,K", [(SCHED_STREAMING, 1), (SCHED_STREAMING, 3), (SCHED_FUSED, 1), (SCHED_FUSED, 3),
                                        (SCHED_ROWSTREAM, 1), (SCHED_WAVE, 3)])
def test_two_row_strips_equal_whole_cloth(oracle, schedule, K):
    """Strip decomposition is exact: 2 strips + ghost-row exchange == whole cloth, bit for bit."""
    nx, ny, split = 140, 120, 64
    whole = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K)
    up = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K, strip_row0=0, strip_rows=split)
    lo = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K, strip_row0=split, strip_rows=ny - split)
    g, gu, gl = clothgpu.Cloth(whole), clothgpu.Cloth(up), clothgpu.Cloth(lo)
    for h in (g, gu, gl):
        h.set_schedule(schedule)
    st = random_state(nx, ny, seed=77 + K, pins=0.0)
    g.write_state(*st)
    gu.write_state(*[a.reshape(ny, nx)[:split].ravel() for a in st])
    gl.write_state(*[a.reshape(ny, nx)[split:].ravel() for a in st])
    _exchange_halos_same_device(gu, gl)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 3000, 3000
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 10.0
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 500.0, 164.0 + 6 * split, 495.0, 160.0 + 6 * split
    for t in range(4):
        g.step(f)
        gu.step(f)
        gl.step(f)
        _exchange_halos_same_device(gu, gl)
        full = g.read_state()
        parts = [np.concatenate([a, b]) for a, b in zip(gu.read_state(), gl.read_state())]
        assert_same_state(parts, full, f"strips frame {t}")
    cu, cl, cw = gu.counters(), gl.counters(), g.counters()
    for k in ("live_sticks", "alive_sticks", "relaxations", "torn_last", "inactive"):
        assert cu[k] + cl[k] == cw[k], k


def test_strip_rejects_too_many_iterations():
    d = _abi.grid_desc(64, 64, max_iterations=2, strip_row0=0, strip_rows=32)
    g = clothgpu.Cloth(d)
    with pytest.raises(clothgpu.ClothGpuError):
        g.step(_abi.default_frame(3))


def test_full_size_1024_fused_equals_streaming_and_oracle(oracle):
    """BASELINE configs[1] at full size: 1024x1024, K=8, top row pinned.  The oracle (OpenMP) checks the
    first frames; fused == streaming is the size-independent property for the rest."""
    desc = _abi.grid_desc(1024, 1024, pin_rule=_abi.PIN_TOP_ROW, max_iterations=8)
    gf, gs = clothgpu.Cloth(desc), clothgpu.Cloth(desc)
    gf.set_schedule(SCHED_FUSED)
    gs.set_schedule(SCHED_STREAMING)
    gw = clothgpu.Cloth(desc)
    gw.set_schedule(SCHED_WAVE)
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    f = _abi.default_frame(8)
    f.win_w, f.win_h = 128 + 6 * 1024 + 128, 164 + 6 * 1024 * 4
    for t in range(6):
        gf.step(f)
        gs.step(f)
        gw.step(f)
        o.step(f, threads=8)
    assert_same_state(gf.read_state(), o.read_state(), "fused vs oracle")
    assert_same_state(gw.read_state(), o.read_state(), "wave vs oracle")
    for t in range(40):
        gf.step(f)
        gs.step(f)
        gw.step(f)
    assert_same_state(gf.read_state(), gs.read_state(), "fused vs streaming after 46 frames")
    assert_same_state(gw.read_state(), gs.read_state(), "wave vs streaming after 46 frames")
    assert gw.counters()["relaxations"] == 46 * 8 * (2 * 1024 * 1024 - 2048)
    assert gf.counters()["relaxations"] == 46 * 8 * (2 * 1024 * 1024 - 2048)


@pytest.mark.parametrize("K", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("nx,ny,ens", [(1300, 900, 1), (515, 2051, 2), (260, 4400, 1)])
def test_wavefront_schedule_equals_streaming_on_many_chunk_shapes(K, nx, ny, ens):
    """The wavefront kernel's worker sets run up to one step apart (set-local barriers chained by producer/consumer
    barriers, csrc/cloth_wave.cuh).  Shapes with several bands, several row chunks per band column and spans that run over
    the end of a band column, every sweep count (the stage split between the sets depends on it), random jitter with
    tears, pins and holes: the state after 12 frames must equal the one-launch-per-colour streaming schedule bit for bit.
    (GPU against GPU: two independent implementations; the oracle pins both on the smaller shapes above.)"""
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K, ensemble=ens)
    gw, gs = clothgpu.Cloth(desc), clothgpu.Cloth(desc)
    gw.set_schedule(SCHED_WAVE)
    gs.set_schedule(SCHED_STREAMING)
    for e in range(ens):
        st = random_state(nx, ny, seed=K * 31 + nx + e)
        gw.write_state(*st, cloth=e)
        gs.write_state(*st, cloth=e)
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 9.5
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 128.0 + 3 * nx, 164.0 + 3 * ny, 120.0 + 3 * nx, 170.0 + 3 * ny
    for t in range(12):
        gw.step(f)
        gs.step(f)
    for e in range(ens):
        assert_same_state(gw.read_state(cloth=e), gs.read_state(cloth=e), f"{nx}x{ny} K={K} cloth {e}: wave vs streaming")
    cw, cs = gw.counters(), gs.counters()
    for key in ("relaxations", "torn_last", "live_sticks", "alive_sticks", "pinned", "inactive"):
        assert cw[key] == cs[key], key
    gw.close()
    gs.close()


@pytest.mark.parametrize("K,schedule", [(1, SCHED_FUSED), (1, SCHED_ROWSTREAM), (3, SCHED_FUSED), (3, SCHED_WAVE),
                                        (8, SCHED_WAVE), (11, SCHED_FUSED)])
def test_attached_strips_exchange_inside_the_frame_kernel(oracle, K, schedule):
    """Three strips of one cloth, attached to each other: the fused kernel stores the edge rows into the neighbours'
    ghost rows itself (same-process attach = plain pointers; other processes use the IPC path, test_strips_multiproc).
    K = 11 splits the frame into two launches with an exchange after each.  Result == whole cloth, bit for bit."""
    from clothgpu.strips import StripPlan
    nx, ny = 150, 132
    ghost = 2 * K
    plan = StripPlan(ny, 3, ghost=min(ghost, 16))
    whole = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K))
    strips = []
    for r in range(3):
        row0, rows = plan.rows(r)
        strips.append(clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=K,
                                                    strip_row0=row0, strip_rows=rows)))
    blobs = [s.strip_export() for s in strips]
    for r, s in enumerate(strips):
        s.strip_attach(blobs[r - 1] if r > 0 else None, blobs[r + 1] if r < 2 else None)
        s.set_schedule(schedule)
    whole.set_schedule(schedule)
    st = random_state(nx, ny, seed=5 + K, pins=0.002)
    whole.write_state(*st)
    for r, s in enumerate(strips):
        row0, rows = plan.rows(r)
        s.write_state(*[a.reshape(ny, nx)[row0:row0 + rows].ravel() for a in st])
    for s in strips:
        s.strip_push()
    f = _abi.default_frame(K)
    f.win_w, f.win_h = 3000, 3000
    f.buttons = MOUSE_DRAGGING
    f.tear_length = 10.0
    b1 = plan.bounds[1]
    f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py = 500.0, 164.0 + 6 * b1, 495.0, 160.0 + 6 * b1
    for t in range(5):
        whole.step(f)
        for s in strips:
            s.step(f)
        parts = [np.concatenate(p) for p in zip(*[s.read_state() for s in strips])]
        assert_same_state(parts, whole.read_state(), f"attached strips frame {t}")
    cw = whole.counters()
    cs = [s.counters() for s in strips]
    for k in ("live_sticks", "alive_sticks", "relaxations", "torn_last", "inactive"):
        assert sum(c[k] for c in cs) == cw[k], k
    with pytest.raises(clothgpu.ClothGpuError):
        strips[0].set_schedule(SCHED_STREAMING)
        strips[0].step(f)


def _run_strip_worker(nproc, args, timeout=300):
    import socket
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={nproc}",
           "--master-addr", "127.0.0.1", "--master-port", str(port), os.path.join(here, "strip_worker.py")] + args
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=timeout)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "STRIPS-OK" in r.stdout, r.stdout[-3000:]


@pytest.mark.parametrize("K,schedule", [(1, "fused"), (1, "rowstream"), (8, "fused"), (8, "wave")])
def test_strips_multiproc_one_gpu_ipc(K, schedule):
    """Two processes share cuda:0: the neighbour's memory is mapped through cudaIpcOpenMemHandle (the path real
    multi-GPU runs take), rendezvous over gloo.  Rank 0 compares the gathered strips with the whole cloth."""
    _run_strip_worker(2, ["--backend", "gloo", "--same-device", "--nx", "200", "--ny", "160", "--iterations", str(K),
                          "--frames", "6", "--schedule", schedule])


def test_strips_multi_gpu_nccl():
    """One process per GPU over NCCL (rendezvous) + NVLink peer stores; needs >= 2 GPUs."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs")
    _run_strip_worker(min(n, 8), ["--backend", "nccl", "--nx", "1024", "--ny", "1024", "--iterations", "8", "--frames", "6"])


@pytest.mark.parametrize("nx,ny,E,K", [(128, 128, 6, 8), (100, 91, 7, 8), (64, 40, 9, 4), (17, 300, 3, 5), (128, 130, 5, 6)])
def test_packed_ensembles_with_one_frame_per_cloth(oracle, nx, ny, E, K):
    """BASELINE configs[3] as SURVEY §8(d) defines it: every cloth of the ensemble has its own mouse.  clothgpu_step_each on
    the packed wavefront path: the entry warps use the frame of the cloth a row belongs to, the sweeps learn per stick
    whether its cloth is being dragged (block state bits 12..15).  Neighbouring cloths in the band (side by side and
    stacked) get different buttons on purpose: a drag next to an idle cloth next to a hole punch."""
    desc = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, ensemble=E, max_iterations=K)
    g = clothgpu.Cloth(desc)
    g.set_schedule(SCHED_WAVE)
    one = _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, max_iterations=K)
    oracles = [oracle.OracleCloth(one, oracle.COLOUR) for _ in range(E)]
    for e in range(E):
        st = random_state(nx, ny, seed=2000 + 13 * e + ny)
        g.write_state(*st, cloth=e)
        oracles[e].write_state(*st)

    def frames_at(t):
        out = []
        for e in range(E):
            f = _abi.default_frame(K)
            f.win_w, f.win_h = 128 + 6 * nx + 50, 164 + 6 * ny + 20
            f.tear_length = 9.5
            f.mouse_x, f.mouse_y = 128.0 + 6 * nx * ((e * 37 + t * 11) % 17) / 16.0, 164.0 + 6 * ny * ((e * 5 + t * 3) % 7) / 6.0
            f.mouse_px, f.mouse_py = f.mouse_x - 8.0 - e, f.mouse_y + 5.0
            kind = (e + t) % 4
            if kind == 0:
                f.buttons = _abi.MOUSE_LEFT | MOUSE_DRAGGING
                f.mouse_force = 1.0 + e
            elif kind == 1:
                f.buttons = _abi.MOUSE_RIGHT
                f.scroll_y = 35.0 + e
            elif kind == 2:
                f.buttons = _abi.MOUSE_LEFT | _abi.MOUSE_CTRL
            f.slider[1] = 200.0 + 10 * e          # gravity differs per cloth too
            out.append(f)
        return out

    torn = 0
    for t in range(4):
        fr = frames_at(t)
        g.step_each(fr)
        for e in range(E):
            oracles[e].step(fr[e])
            torn += oracles[e].torn_last
    assert g.last_schedule()[1] == "wave_frame_kernel<double,PER_CLOTH,PACK>"
    relax = 0
    for e in range(E):
        assert_same_state(g.read_state(cloth=e), oracles[e].read_state(), f"{nx}x{ny} K={K} cloth {e} of {E}, per-cloth frames")
        relax += oracles[e].relaxations
    assert torn > 0, "dragged cloths tear, the others must not"
    c = g.counters()
    assert c["relaxations"] == relax
    bad = frames_at(0)
    bad[1].relax_gain = 0.3
    with pytest.raises(clothgpu.ClothGpuError):
        g.step_each(bad)
