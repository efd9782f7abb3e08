"""clothgpu_save_state / clothgpu_load_state: snapshot -> new handle -> resume == the uninterrupted run, bit for bit
(SURVEY.md §8(f) 3), and the file is the one clothgpu/statefile.py and the Go harness read and write."""
import numpy as np
import pytest

import clothgpu
from clothgpu import _abi, statefile
from clothgpu._abi import SCHED_FUSED, SCHED_ROWSTREAM
from traces import c0_trace

from test_gpu_parity import assert_same_state, random_state

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("schedule", [SCHED_FUSED, SCHED_ROWSTREAM])
def test_save_load_resume_equals_the_uninterrupted_run(tmp_path, oracle, schedule):
    frames = c0_trace(700)
    a = clothgpu.Cloth(_abi.default_desc())
    a.set_schedule(schedule)
    for f in frames[:620]:                       # past the ctrl-pin, the drag, the hole punch, into the tearing zig-zag
        a.step(f)
    path = tmp_path / "c0_620.clstate"
    a.save_state(path)
    b = clothgpu.Cloth(_abi.default_desc())      # a NEW handle
    b.set_schedule(schedule)
    b.load_state(path)
    assert b.counters()["frames"] == 620
    for k in ("live_sticks", "alive_sticks", "pinned", "inactive", "highlighted"):
        assert a.counters()[k] == b.counters()[k], k
    for f in frames[620:670]:                    # 50 more frames on both
        a.step(f)
        b.step(f)
    assert_same_state(b.read_state(), a.read_state(), "resumed vs uninterrupted")
    # the file is what the numpy reader sees, and what the oracle would write
    hdr, cloths = statefile.read(path)
    assert (hdr["nx"], hdr["ny"], hdr["frames"], hdr["ensemble"]) == (171, 36, 620, 1)
    o = oracle.OracleCloth(_abi.default_desc(), oracle.COLOUR)
    for f in frames[:620]:
        o.step(f)
    assert_same_state(cloths[0], o.read_state(), "snapshot vs oracle")


def test_snapshot_of_an_ensemble_and_rejections(tmp_path):
    nx, ny, E = 70, 33, 3
    g = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, ensemble=E, max_iterations=2))
    states = [random_state(nx, ny, seed=e) for e in range(E)]
    for e, st in enumerate(states):
        g.write_state(*st, cloth=e)
    g.step(_abi.default_frame(2))
    p = tmp_path / "ens.clstate"
    g.save_state(p)
    h = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_NONE, ensemble=E, max_iterations=2))
    h.load_state(p)
    for e in range(E):
        assert_same_state(h.read_state(cloth=e), g.read_state(cloth=e), f"cloth {e}")
    other = clothgpu.Cloth(_abi.grid_desc(nx + 1, ny, pin_rule=_abi.PIN_NONE, ensemble=E))
    with pytest.raises(clothgpu.ClothGpuError, match="holds"):
        other.load_state(p)
    raw = bytearray(open(p, "rb").read())
    raw[200] ^= 1
    open(p, "wb").write(raw)
    with pytest.raises(clothgpu.ClothGpuError, match="checksum"):
        h.load_state(p)
    with pytest.raises(clothgpu.ClothGpuError):
        h.load_state(tmp_path / "missing.clstate")
    # written by the numpy writer, read by the library
    statefile.write(p, nx, ny, states, frames=7)
    h.load_state(p)
    assert h.counters()["frames"] == 7
    x, y, px, py, fl = h.read_state(cloth=2)
    assert np.array_equal(x, states[2][0]) and np.array_equal(py, states[2][3])


def test_export_respects_the_callers_capacity():
    """clothgpu_read_quads_f32 / read_segments_f32 copy small exports back speculatively (count and data in one
    synchronisation): never beyond the caller's capacity, and a too small capacity is an error, not an overrun."""
    import ctypes as C
    g = clothgpu.Cloth(_abi.default_desc())
    for f in c0_trace(120):
        g.step(f)
    quads, hl = g.read_quads()
    n = len(quads)
    assert n == g.counters()["live_sticks"] > 12000
    guard = 64
    for cap, ok in ((n, True), (n + 3, True), (n - 5, False), (0, False)):
        qbuf = np.full((cap + guard) * 8, np.float32(-7.0), np.float32)
        hbuf = np.full(cap + guard, 0xAB, np.uint8)
        if ok:
            got = g.read_quads_into(qbuf.ctypes.data, hbuf.ctypes.data, cap)
            assert got == n
            assert np.array_equal(qbuf[:n * 8].reshape(-1, 8), quads) and np.array_equal(hbuf[:n], hl)
        else:
            with pytest.raises(clothgpu.ClothGpuError) as e:
                g.read_quads_into(qbuf.ctypes.data, hbuf.ctypes.data, cap)
            assert e.value.code == _abi.ERR_CAPACITY
        assert np.all(qbuf[cap * 8:] == np.float32(-7.0)) and np.all(hbuf[cap:] == 0xAB), "wrote past the caller's capacity"
    assert g.count_segments() == n
