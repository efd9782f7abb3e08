"""On-disk formats, CPU side: the JSON-lines mouse trace (recorder / replayer) and the CLSTATE1
snapshot file.  The GPU round trip of the snapshot is tests/test_gpu_formats.py."""
import ctypes as C
import os

import numpy as np
import pytest

from clothgpu import _abi, statefile
from clothgpu.mouse import EventAdapter
from clothgpu.trace import RecordingMouse, TraceRecorder, load_trace, replay, save_trace
from traces import c0_trace

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def frame_bytes(f):
    return bytes(C.string_at(C.byref(f), C.sizeof(_abi.Frame)))


def test_committed_trace_replays_to_the_scripted_frames_bit_for_bit():
    """tests/golden/c0_trace.jsonl (make_golden.py --dump-trace) is the scripted trace as a file: replaying it through
    nothing but the Mouse setters gives the very frames the generator builds through the event adapter."""
    want = c0_trace(1000)
    got = list(replay(load_trace(os.path.join(GOLDEN, "c0_trace.jsonl"))))
    assert len(got) == len(want) == 1000
    for t, (a, b) in enumerate(zip(got, want)):
        assert frame_bytes(a) == frame_bytes(b), f"frame {t}"


def test_recorder_round_trip_with_several_moves_per_frame(tmp_path):
    """Several UpdatePosition calls in one frame (only the last two positions matter), a held-still drag
    (px,py keep the last delta, §3.4a), a wall-clock force (§3.4b), scrolling, window offsets."""
    rec = TraceRecorder(RecordingMouse())
    ad = EventAdapter(rec.mouse)
    frames = []

    def end(**kw):
        f = _abi.default_frame(kw.get("iterations", 1))
        if "win" in kw:
            f.win_w, f.win_h = kw["win"]
        if "win_off" in kw:
            f.win_off_x, f.win_off_y = kw["win_off"]
        ad.mouse.fill(f)
        frames.append(f)
        rec.end_frame(**kw)

    ad.move(10.25, 20.5)
    ad.move(11.1, 21.3)
    ad.move(12.7, 22.9)
    end()
    ad.press(12.7, 22.9, EventAdapter.PRIMARY)
    end(win=(1500, 900))
    ad.begin_frame(0.123456789)
    ad.drag(40.0, 41.0, EventAdapter.PRIMARY)
    ad.drag(47.3, 39.9, EventAdapter.PRIMARY)
    end(win=(1500, 900), win_off=(3.0, -2.0))
    ad.begin_frame(0.2)
    end()                                  # held still: no move, the previous delta stays
    ad.scroll_by(7.5)
    end(iterations=3)
    ad.release()
    ad.press(5.0, 6.0, EventAdapter.SECONDARY, ctrl=True)
    end()
    path = tmp_path / "t.jsonl"
    rec.save(path)
    assert len(rec.records[0]["moves"]) == 3 and rec.records[3]["moves"] == []
    got = list(replay(load_trace(path)))
    assert [frame_bytes(a) for a in got] == [frame_bytes(b) for b in frames]
    save_trace(tmp_path / "t2.jsonl", load_trace(path))
    assert open(path).read() == open(tmp_path / "t2.jsonl").read()


def test_state_file_round_trip_and_damage_detection(tmp_path, oracle):
    desc = _abi.default_desc()
    o = oracle.OracleCloth(desc, oracle.COLOUR)
    for f in c0_trace(130):
        o.step(f)
    st = o.read_state()
    p = tmp_path / "s.clstate"
    statefile.write(p, 171, 36, [st], frames=130)
    hdr, cloths = statefile.read(p)
    assert hdr == dict(nx=171, ny=36, row0=0, rows=36, ensemble=1, spacing=6, pos=(128, 164), frames=130)
    for a, b in zip(cloths[0], st):
        assert a.dtype == b.dtype and np.array_equal(a, b)
    assert os.path.getsize(p) == 64 + 32 * 6156 + (6156 + 7) // 8 * 8 + 8
    raw = bytearray(open(p, "rb").read())
    raw[1000] ^= 0x10
    open(p, "wb").write(raw)
    with pytest.raises(ValueError, match="checksum"):
        statefile.read(p)
    open(p, "wb").write(raw[:-9])
    with pytest.raises(ValueError, match="truncated"):
        statefile.read(p)
