"""Recorder / replayer of mouse traces, one JSON object per frame (JSON lines) — SURVEY.md §3.4:

    {"moves": [[x, y], ...], "left": b, "right": b, "dragging": b, "ctrl": b, "force": f64, "scrollY": f32,
     "win": [w, h], "win_off": [ox, oy], "dt": f64, "iterations": K}                      (the last four are optional)

`moves` are the arguments of every Mouse.UpdatePosition call of the frame, in order (physics/mouse.go:21-27: each call
shifts x,y into px,py, so a held-still drag keeps its last delta and only the last two positions of a frame matter);
the four bools, `force` and `scrollY` are the Mouse's fields when (*Cloth).Update reads them.  In the GUI `force` is
wall-clock driven (main.go:174-177), which is why a deterministic trace carries it explicitly.  Floats are written with
repr(), so a trace replays bit for bit.

Replaying needs nothing but the exported setters of physics.Mouse, so the same file drives
  * this package (-> clothgpu_frame for the GPU library and the tests' CPU checker), and
  * the unmodified Go reference (host/go/harness/main.go).
Pure bookkeeping: no physics here."""
import json

from . import _abi
from .mouse import Mouse


class RecordingMouse(Mouse):
    """physics.Mouse that remembers the UpdatePosition calls of the current frame."""

    def __init__(self):
        super().__init__()
        self.moves = []

    def UpdatePosition(self, x, y):
        super().UpdatePosition(x, y)
        self.moves.append([float(x), float(y)])


class TraceRecorder:
    """Attach to an EventAdapter built on a RecordingMouse; call end_frame() once per frame, after the frame's events."""

    def __init__(self, mouse):
        if not isinstance(mouse, RecordingMouse):
            raise TypeError("TraceRecorder needs a RecordingMouse")
        self.mouse, self.records = mouse, []

    def end_frame(self, win=None, win_off=None, dt=None, iterations=None):
        m = self.mouse
        rec = {"moves": m.moves, "left": m.leftDown, "right": m.rightDown, "dragging": m.isDragging, "ctrl": m.ctrlDown,
               "force": m.force, "scrollY": m.scrollY}
        if win is not None:
            rec["win"] = [int(win[0]), int(win[1])]
        if win_off is not None:
            rec["win_off"] = [float(win_off[0]), float(win_off[1])]
        if dt is not None:
            rec["dt"] = float(dt)
        if iterations is not None:
            rec["iterations"] = int(iterations)
        m.moves = []
        self.records.append(rec)
        return rec

    def save(self, path):
        save_trace(path, self.records)


def save_trace(path, records):
    with open(path, "w") as f:
        for r in records:
            f.write(json.dumps(r, separators=(",", ":")) + "\n")


def load_trace(path):
    with open(path) as f:
        return [json.loads(line) for line in f if line.strip()]


def replay(records, mouse=None, scroll0=50.0, max_scroll=120.0):
    """Yield one clothgpu_frame per record, applying it to `mouse` through the setters of physics/mouse.go only."""
    m = mouse or Mouse()
    m.SetScrollY(scroll0)         # consts.DefaultFocusArea, main.go:84
    m.SetMaxScrollY(max_scroll)   # consts.MaxFocusArea, main.go:85
    for r in records:
        for x, y in r["moves"]:
            m.UpdatePosition(x, y)
        m.SetLeftButton() if r["left"] else m.ReleaseLeftButton()
        m.SetRightButton() if r["right"] else m.ReleaseRightButton()
        m.SetDragging(r["dragging"])
        m.SetCtrlDown(r["ctrl"])
        m.SetForce(r["force"])
        m.SetScrollY(r["scrollY"])
        f = _abi.default_frame(r.get("iterations", 1))
        if "win" in r:
            f.win_w, f.win_h = r["win"]
        if "win_off" in r:
            f.win_off_x, f.win_off_y = r["win_off"]
        if "dt" in r:
            f.dt = r["dt"]
        m.fill(f)
        yield f
