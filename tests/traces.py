"""Deterministic scripted mouse traces: a list of clothgpu_frame per frame, produced by
replaying pointer events through the mirror of the reference's event handling (clothgpu.mouse)."""
from clothgpu import _abi
from clothgpu.mouse import EventAdapter

P, S = EventAdapter.PRIMARY, EventAdapter.SECONDARY


def c0_trace(n_frames=1000, iterations=1, scale=1.0, win=(1280, 820), pin_click=(128 + 6 * 40, 164 + 6 * 10), recorder=None):
    """frame 2 ctrl-click on particle (40,10) while the cloth is still (almost) at its rest grid, so the click really
    pins it; frames 3-99 idle; 100-399 slow left-drag sweep with the force ramp; 400-449 released;
    450-549 right-button hold moving along a line (hole punch); 550-599 scroll the focus area to 120;
    600-799 fast zig-zag left-drag (tears); 800 ctrl-click on a particle; then idle."""
    ad = EventAdapter(recorder.mouse if recorder is not None else None)
    frames = []
    press_frame = None
    for t in range(n_frames):
        if press_frame is not None:
            ad.begin_frame((t - press_frame) / 60.0)
        if t == 2:
            ad.press(pin_click[0], pin_click[1], P, ctrl=True)
            press_frame = t
        elif t == 3:
            ad.release()
            press_frame = None
        elif t == 100:
            ad.press(300 * scale, 250 * scale, P)
            press_frame = t
        elif 100 < t < 400:
            k = t - 100
            ad.drag((300 + 700 * k / 299.0) * scale, (250 + 50 * k / 299.0) * scale, P)
        elif t == 400:
            ad.release()
            press_frame = None
        elif t == 450:
            ad.press(200 * scale, 300 * scale, S)
            press_frame = t
        elif 450 < t < 550:
            ad.move((200 + 8 * (t - 450)) * scale, 300 * scale, S)
        elif t == 550:
            ad.release()
            press_frame = None
        elif 550 < t < 600:
            ad.scroll_by(2.0)
        elif t == 600:
            ad.press(400 * scale, 200 * scale, P)
            press_frame = t
        elif 600 < t < 800:
            k = t - 600
            zig = 60 if (k // 4) % 2 == 0 else -60
            ad.drag((400 + 25 * (k % 16) + 2 * k) * scale, (260 + zig + 0.4 * k) * scale, P)
        elif t == 800:
            ad.release()
            press_frame = None
            ad.press(pin_click[0], pin_click[1], P, ctrl=True)
            press_frame = t
        elif t == 801:
            ad.release()
            press_frame = None
        f = _abi.default_frame(iterations)
        f.win_w, f.win_h = win
        ad.mouse.fill(f)
        if recorder is not None:   # clothgpu.trace.TraceRecorder: the frame as a JSON-lines record (replayable by the Go harness)
            recorder.end_frame(win=win, dt=f.dt, iterations=iterations)
        frames.append(f)
    return frames
