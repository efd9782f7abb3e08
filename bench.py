#!/usr/bin/env python
"""bench.py — throughput of the per-frame cloth update (stick relaxations/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME]

One "step" = one frame: the particle loop + `iterations` colour-ordered relaxation sweeps with tearing
(physics/cloth.go:92-100 of the reference) over the rank's share of the workload.

Default workload `c4t` — the north_star target: ONE 16384 x 16384 cloth (268 M particles, the largest single-GPU
configuration of BASELINE.json), ONE sweep per frame (the reference's own semantics), a left-button drag that tears
sticks in EVERY timed frame (the DRAG instantiation of the frame kernel), binary64.  N = 1 steps the whole cloth; N > 1
cuts it into row strips, one per GPU, whose edge rows travel inside the frame kernel (NVLink peer stores + in-kernel
epoch handshake): strong scaling, and before timing the N strips are checked bit for bit against the whole cloth.
The same JSON line carries a `workloads` object with the other BASELINE configurations (see SUBS below).

Timing: W warm-up steps, then blocks of EXACTLY K steps, each bracketed by barrier + synchronize and timed with CUDA
events on the launching stream, max over ranks; the block is repeated until >= 0.5 s have been timed and the MEDIAN block
is reported (`timing` says how many).  Prints ONE JSON line.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "cloth-physics_b200", "python"), os.path.join(ROOT, "oracle"), os.path.join(ROOT, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)

BYTES_PER_PARTICLE_F64 = 65.1   # x,y,px,py read+write (64 B) + 9 flag bits
BYTES_PER_PARTICLE_F32 = 33.1
FP64_INSTR_PER_RELAX = 32.0     # binary64-pipe thread instructions per stretched stick (31 + the rest-length test), DESIGN.md §3.1
FP64_PIPE_PEAK = 1.8e13         # thread-instructions/s measured by tools/pipe_peaks.cu on B200 (profiles/README.md)
MIN_TIMED_S = 0.5
MAX_BLOCKS = 160

WORKLOADS = {
    # name: (nx, ny, sweeps, cloths per GPU [0 = one cloth in row strips over the ranks], pin rule, trace, description)
    "c4t": (16384, 16384, 1, 0, "top_row", "tear",
            "north_star target / BASELINE configs[4] geometry: one 16384x16384 cloth, 1 sweep/frame (reference semantics), "
            "drag that tears sticks in every timed frame, f64; whole cloth on 1 GPU, row strips with in-kernel halo exchange on N"),
    "c4": (16384, 16384, 8, 0, "top_row", "idle", "BASELINE configs[4]: one 16384x16384 cloth in row strips over the GPUs, 8 sweeps/frame, f64"),
    "c4k1": (16384, 16384, 1, 0, "top_row", "idle", "one 16384x16384 cloth in row strips over the GPUs, 1 sweep/frame, idle mouse, f64"),
    "c4s": (4096, 4096, 8, 0, "top_row", "idle", "one 4096x4096 cloth in row strips over the GPUs, 8 sweeps/frame, f64"),
    "c4st": (4096, 4096, 1, 0, "top_row", "tear", "one 4096x4096 cloth in row strips over the GPUs, 1 sweep/frame, tearing drag, f64"),
    "c1": (1024, 1024, 8, 8, "top_row", "idle", "BASELINE configs[1]: 1024x1024 cloth, 8 sweeps/frame, top row pinned, f64 (8 cloths per GPU: 277 MB > L2)"),
    "c1k1": (1024, 1024, 1, 8, "top_row", "idle", "1024x1024 cloth, 1 sweep/frame (reference semantics), top row pinned, f64"),
    "c0": (171, 36, 1, 1, "reference", "c0", "BASELINE configs[0]: reference default GUI cloth 171x36, 1 sweep/frame, scripted 1000-frame drag/tear/hole trace"),
    "c2": (4096, 4096, 8, 1, "top_row", "idle", "BASELINE configs[2] geometry: 4096x4096 cloth, 8 sweeps/frame, f64"),
    "c2k1": (4096, 4096, 1, 1, "top_row", "idle", "4096x4096 cloth, 1 sweep/frame (reference semantics), f64"),
    "c2t": (4096, 4096, 1, 1, "top_row", "c0x24", "BASELINE configs[2]: 4096x4096 cloth, scripted drag/tear/hole-punch trace, live-stick compaction every frame, 1 sweep/frame, f64"),
    "c2t8": (4096, 4096, 8, 1, "top_row", "c0x24", "BASELINE configs[2]: 4096x4096 cloth, scripted drag/tear/hole-punch trace, live-stick compaction every frame, 8 sweeps/frame, f64"),
    "c3": (128, 128, 8, 2048, "top_row", "per_cloth", "BASELINE configs[3] shard: 2048 independent 128x128 cloths per GPU, 8 sweeps/frame, per-cloth mouse (phase = cloth id)"),
    "c3u": (128, 128, 8, 2048, "top_row", "idle", "2048 independent 128x128 cloths per GPU, 8 sweeps/frame, one idle frame for all"),
}
# what the default run reports besides the headline (name -> which ranks take part)
SUBS = ("c4", "c4t@f32", "c1", "c2t", "c3", "c0")     # "@f32": the optional binary32 mode of the same workload


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def offline_ncu(workload):
    """DRAM bytes per launch of the dominant kernel from the committed `ncu --set full` capture of the same kernel and
    shape (profiles/ncu_facts.json names the file).  ncu cannot run inside a bench, so this is NOT measured in this run."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_facts.json")) as f:
            return json.load(f).get(workload) or {}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed regions."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.samples, self._stop, self._t = index, [], threading.Event(), None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                      "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                parts = [p.strip() for p in out.strip().split(",")]
                if len(parts) >= 6:
                    self.samples.append(parts)
            except Exception:
                pass
            self._stop.wait(0.05)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit())
        mx = [float(s[1]) for s in self.samples if s[1].replace(".", "").isdigit()]
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        reasons = sorted({n for s in self.samples for n, v in zip(names, s[2:6]) if v.lower().startswith("active")})
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(self.samples)}


# ---------------------------------------------------------------------------------------------------------------------
# frames
def window_for(nx, ny):
    return 128 + 6 * nx + 128, 164 + 6 * ny * 4     # large enough for the cloth to hang freely


def tear_sweep_trace(n_frames, nx, ny, K, period=600):
    """Left-button drag from near the top-left to near the bottom-right corner of the cloth (in `period` frames, then
    back, and so on), zig-zagging: the mouse pushes the particles within the drag radius by the clamped mouse delta times
    the drag force (particle.go:108-125), their sticks stretch past 150 px and tear while the drag lasts
    (constraint.go:35-39).  Every frame after the press is a dragging frame; one traverse crosses every strip boundary.
    Events go through the mirror of the reference's event handling (clothgpu.mouse.EventAdapter, main.go:260-309); the
    force ramp is pinned at 3 s (drag force 15).  Frame t does not depend on n_frames."""
    from clothgpu import _abi
    from clothgpu.mouse import EventAdapter
    ad, P = EventAdapter(), EventAdapter.PRIMARY
    w, h = window_for(nx, ny)
    x0, y0, x1, y1 = 128 + 0.15 * 6 * nx, 164 + 0.04 * 6 * ny, 128 + 0.85 * 6 * nx, 164 + 0.96 * 6 * ny
    frames = []
    for t in range(n_frames):
        a = (t % (2 * period)) / float(period)
        a = a if a <= 1.0 else 2.0 - a
        zig = 45.0 if (t // 3) % 2 == 0 else -45.0
        x, y = x0 + a * (x1 - x0) + zig, y0 + a * (y1 - y0) + 0.5 * zig
        if t == 0:
            ad.press(x, y, P)
        else:
            ad.begin_frame(3.0)
            ad.drag(x, y, P)
        f = _abi.default_frame(K)
        f.win_w, f.win_h = w, h
        ad.mouse.fill(f)
        frames.append(f)
    return frames


def per_cloth_frames(step, cloths, nx, ny, K, first_id=0):
    """BASELINE configs[3]: every cloth has its own mouse, phase = cloth id (`seed = cloth_id`): one cloth in four drags
    (and tears) along its own zig-zag, one in four hovers with the focus area, the rest are idle."""
    from clothgpu import _abi
    w, h = window_for(nx, ny)
    out = []
    for e in range(cloths):
        cid = first_id + e
        f = _abi.default_frame(K)
        f.win_w, f.win_h = w, h
        ph = (cid * 37 + step * 5) % 97 / 96.0
        f.mouse_x, f.mouse_y = 128 + 6 * nx * ph, 164 + 6 * ny * ((cid * 11 + step * 3) % 53 / 52.0)
        f.mouse_px, f.mouse_py = f.mouse_x - 9.0, f.mouse_y + 4.0
        if cid % 4 == 0:
            f.buttons = _abi.MOUSE_LEFT | _abi.MOUSE_DRAGGING
            f.mouse_force = 2.0 + (cid % 7)
        elif cid % 4 == 1:
            f.scroll_y = 60.0
        else:
            f.mouse_x = f.mouse_y = f.mouse_px = f.mouse_py = 0.0
        out.append(f)
    return out


def make_frames(wl, n, nx, ny, K):
    """n frames of the workload's trace (idle workloads repeat one frame; the scripted c0 trace scaled to 4096^2 is the
    window [380, 800) of it and wraps around)."""
    from clothgpu import _abi
    trace = WORKLOADS[wl][5]
    if trace == "tear":
        return tear_sweep_trace(n, nx, ny, K)
    if trace == "c0":
        from traces import c0_trace
        return c0_trace(n, iterations=K)
    if trace == "c0x24":
        # the c0 trace scaled x24 to the 4096^2 cloth: frames 380.. cover the end of the slow drag, the
        # release, the right-button hole punch with a moving focus area, the scroll and the fast zig-zag
        from traces import c0_trace
        win = c0_trace(800, iterations=K, scale=24.0 * nx / 4096.0, win=window_for(nx, ny))[380:]
        return [win[i % len(win)] for i in range(n)]
    f = _abi.default_frame(K)
    f.win_w, f.win_h = window_for(nx, ny)
    return [f] * n


# ---------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle's restatement of the reference's CPU path (TEST INFRASTRUCTURE; only this leg and --impl reference
# execute it).  Sequential Gauss-Seidel sweep in creation order incl. the slice-removal artefact, AoS + pointers as in the
# Go program, 1 thread = the reference's single physics goroutine (main.go:87-100).
def cpu_shape(wl):
    nx, ny, K, cloths, pin, trace, _ = WORKLOADS[wl]
    if nx * ny > 2048 * 2048:
        nx = ny = 2048      # bounded sample: per-stick cost does not depend on the size once the cloth is out of cache
    return nx, ny


def cpu_run(wl, seconds_target=None, steps=None, warmup=1):
    import pyoracle
    from clothgpu import _abi
    nx0, ny0, K, cloths, pin, trace, desc_txt = WORKLOADS[wl]
    nx, ny = cpu_shape(wl)
    if wl == "c0":
        d = _abi.default_desc()
    else:
        d = _abi.grid_desc(nx, ny, pin_rule={"top_row": _abi.PIN_TOP_ROW, "reference": _abi.PIN_REFERENCE}[pin], max_iterations=K)
    o = pyoracle.OracleCloth(d, pyoracle.SEQ_FAITHFUL)
    n_max = (steps if steps is not None else 2000) + warmup
    frames = make_frames(wl if trace != "per_cloth" else "c3u", n_max, nx, ny, K)
    for i in range(warmup):
        o.step(frames[i], 1)
    r0, n, t0 = o.relaxations, 0, time.perf_counter()
    while True:
        o.step(frames[warmup + n], 1)
        n += 1
        dt = time.perf_counter() - t0
        if (steps is not None and n >= steps) or (steps is None and (dt >= seconds_target or n >= 2000)):
            break
    relax = o.relaxations - r0
    return {"value": relax / dt, "unit": "stick relaxations/s", "cores": 1, "kind": "port",
            "sample": f"{n} frames of one {nx}x{ny} cloth ({'the whole workload cloth' if (nx, ny) == (nx0, ny0) else 'a bounded sample of the workload cloth'}), "
                      f"{K} sweep(s)/frame, trace '{trace}', sequential-faithful C restatement of physics/cloth.go:92-100 on one host core "
                      f"(the Go reference cannot be built here: no toolchain), {dt:.1f} s",
            "frames_per_s": n / dt, "ms_per_frame": 1e3 * dt / n, "seconds": dt, "frames": n}


def config_for(wl):
    """`config` of the JSON line: the workload and nothing else, so that both arms (--impl clothgpu / reference) print the
    SAME object for the same workload; how an arm ran it (schedule, sharding, the CPU arm's sample) goes into `run`."""
    nx, ny, K, cloths, pin, trace, desc_txt = WORKLOADS[wl]
    return {"workload": desc_txt, "name": wl, "nx": nx, "ny": ny, "iterations": K, "pin": pin, "trace": trace,
            "cloths": "one cloth (row strips over the GPUs when N > 1)" if cloths == 0 else f"{cloths} independent cloths per GPU",
            "l2": "inputs larger than L2: every frame re-streams the cloth's state from HBM" if nx * ny * max(1, cloths) * 33 > 126e6
                  else "state fits L2"}


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path (its C restatement, oracle/cloth_oracle.h) on the
    host cores, same metric and config; each step = one frame of a bounded sample of the workload."""
    if rank != 0:
        return
    wl = args.workload
    nx0, ny0, K, cloths, pin, trace, desc_txt = WORKLOADS[wl]
    r = cpu_run(wl, steps=args.steps, warmup=args.warmup)
    nx, ny = cpu_shape(wl)
    line = {
        "impl": "reference", "metric": "stick_relaxations_per_sec", "value": r["value"], "unit": "stick relaxations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": r["ms_per_frame"],
        "higher_is_better": True, "scaling": "strong" if cloths == 0 else "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_for(wl),
        "run": {"sample": f"each step = one frame of ONE {nx}x{ny} cloth on one host core (the reference's physics is one goroutine, "
                          f"main.go:87-100, and a sequential Gauss-Seidel sweep does not thread)"},
        "cpu_baseline": {k: r[k] for k in ("value", "unit", "cores", "kind", "sample")},
        "e2e": {"value": r["value"], "unit": "stick relaxations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_sec": r["frames_per_s"],
    }
    print(json.dumps(line), flush=True)


# ---------------------------------------------------------------------------------------------------------------------
class Ctx:
    """Process-wide state of a GPU run: torch.distributed plumbing and helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get("RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        self.args = args

    def start(self):
        torch, dist = self.torch, self.dist
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; clothgpu has no CPU fallback (use --impl reference for the CPU arm)")
        torch.cuda.set_device(self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))
        self.stream = torch.cuda.Stream()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def reduce(self, v, op):
        if self.world == 1:
            return v
        t = self.torch.tensor([v], dtype=self.torch.float64, device="cuda")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def timed(self, fn):
        """fn() issues work on self.stream; returns its device time in ms, max over ranks, barrier + sync on both sides."""
        torch = self.torch
        self.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(self.stream):
            e0.record(self.stream)
            fn()
            e1.record(self.stream)
        e1.synchronize()
        self.barrier()
        return self.reduce(e0.elapsed_time(e1), "MAX")


class Arm:
    """One workload on this rank's GPU: the handle(s), its frames, how a block of steps is issued."""

    def __init__(self, ctx, wl, precision="f64", schedule="auto"):
        import clothgpu
        from clothgpu import _abi
        self.ctx, self.wl, self.abi = ctx, wl, _abi
        self.nx, self.ny, self.K, cloths, self.pin, self.trace, self.desc_txt = WORKLOADS[wl]
        self.strips = cloths == 0
        self.world = ctx.world
        self.precision = _abi.F64 if precision == "f64" else _abi.F32
        self.bpp = BYTES_PER_PARTICLE_F64 if precision == "f64" else BYTES_PER_PARTICLE_F32
        self.holder = None
        self.cloth = None
        rule = {"top_row": _abi.PIN_TOP_ROW, "reference": _abi.PIN_REFERENCE}[self.pin]
        if self.strips:
            self.cloths = 1
            if self.world > 1:
                from clothgpu.strips import StripCloth
                self.holder = StripCloth(self.nx, self.ny, max_iterations=self.K, precision=self.precision, device=ctx.local)
                self.cloth = self.holder.cloth
            else:
                self.cloth = clothgpu.Cloth(_abi.grid_desc(self.nx, self.ny, pin_rule=rule, precision=self.precision, device=ctx.local,
                                                           max_iterations=self.K))
        else:
            self.cloths = cloths
            if wl == "c0":
                d = _abi.default_desc()
                d.device, d.precision = ctx.local, self.precision
            else:
                d = _abi.grid_desc(self.nx, self.ny, pin_rule=rule, precision=self.precision, ensemble=cloths, device=ctx.local, max_iterations=self.K)
            self.cloth = clothgpu.Cloth(d)
        self.cloth.set_schedule({"auto": _abi.SCHED_AUTO, "fused": _abi.SCHED_FUSED, "rowstream": _abi.SCHED_ROWSTREAM,
                                 "wave": _abi.SCHED_WAVE, "streaming": _abi.SCHED_STREAMING}[schedule])
        self.cloth.set_stream(ctx.stream.cuda_stream)
        self.particles = self.cloth.nx * self.cloth.rows * self.cloths      # on this GPU (a strip owns `rows` of the ny rows)
        self.compact = self.trace == "c0x24"                               # BASELINE configs[2]: live-stick compaction every frame
        self._frames = None
        self._arrays = {}

    # frames [a, b) of the workload's trace (a finite trace wraps around)
    def frames(self, a, b):
        import clothgpu
        if self.trace == "per_cloth":
            raise RuntimeError("per-cloth workloads build their frame blocks in issue()")
        if self.trace == "idle":
            a, b = 0, b - a
        if self._frames is None or (self.trace == "tear" and len(self._frames) < b):
            n = {"idle": 4096, "tear": max(b, 4096), "c0": 1000, "c0x24": 420}[self.trace]
            self._frames = make_frames(self.wl, n, self.nx, self.ny, self.K)
            self._arrays = {}
        L = len(self._frames)
        key = (a % L, b - a)
        if key not in self._arrays:
            self._arrays[key] = clothgpu.frame_array([self._frames[i % L] for i in range(a, b)])
        return self._arrays[key]

    def prepare(self, total_frames):
        if self.trace == "per_cloth":
            import clothgpu
            first = self.ctx.rank * self.cloths if self.world > 1 else 0
            self._pc = [clothgpu.frame_array(per_cloth_frames(s, self.cloths, self.nx, self.ny, self.K, first)) for s in range(8)]
        else:
            self.frames(0, min(total_frames, 4096))

    def issue(self, a, b):
        """Frames [a, b) of the trace, back to back on the stream, state resident in HBM."""
        if self.trace == "per_cloth":
            for t in range(a, b):
                self.cloth.step_each(self._pc[t % len(self._pc)])
            return
        arr = self.frames(a, b)
        if not self.compact:
            self.cloth.step_n(arr)
            return
        for i in range(len(arr)):            # tear mask update + compaction of the live-stick list, all on the device
            self.cloth.step(arr[i])
            self.cloth.compact_segments()

    def counters(self):
        return self.cloth.counters()

    def close(self):
        if self.holder is not None:
            self.holder.close()
        elif self.cloth is not None:
            self.cloth.close()
        self.cloth = self.holder = None


def measure(ctx, arm, steps, warmup, clock=None):
    """W warm-up steps, then blocks of exactly `steps` steps until MIN_TIMED_S of device time; the median block."""
    arm.prepare(warmup + MAX_BLOCKS * steps)
    arm.issue(0, warmup)
    arm.cloth.sync()
    blocks = []
    at = warmup
    l0 = arm.cloth.launch_count()
    while True:
        r0 = arm.counters()["relaxations"]
        a, b = at, at + steps
        ms = ctx.timed(lambda: arm.issue(a, b))
        relax = ctx.reduce(float(arm.counters()["relaxations"] - r0), "SUM")
        blocks.append((ms, relax))
        at += steps
        if (sum(m for m, _ in blocks) >= 1e3 * MIN_TIMED_S and len(blocks) >= 3) or len(blocks) >= MAX_BLOCKS:
            break
    launches = arm.cloth.launch_count() - l0
    med = sorted(blocks)[len(blocks) // 2]
    ms, relax = med
    res = {"value": relax / (ms * 1e-3), "ms_per_step": ms / steps, "relax_per_step": relax / steps,
           "timing": {"blocks": len(blocks), "steps_per_block": steps, "timed_s": sum(m for m, _ in blocks) * 1e-3,
                      "block_ms_min": min(m for m, _ in blocks), "block_ms_median": ms, "block_ms_max": max(m for m, _ in blocks)},
           "gpu_launches": int(launches), "frames_used": at}
    return res


def roofline_of(arm, ms_per_launch, kernel, offline_key=None):
    peak, peak_src = peaks()
    alg = arm.bpp * arm.particles
    achieved = alg / (ms_per_launch * 1e-3) / 1e9
    off = offline_ncu(offline_key) if offline_key else {}
    return {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
            "traffic": off.get("dram_bytes_per_launch"),
            "traffic_is": ("offline: ncu --set full capture of this kernel on this shape, " + off["source"] + " — not measured in this run")
                          if off.get("dram_bytes_per_launch") else None,
            "peak_source": peak_src, "kernel": kernel, "algorithmic_bytes_per_launch": alg,
            "note": "achieved = compulsory bytes (65.1 B/particle/frame in f64, independent of the sweep count) / launch time of the frame kernel"}


def frame_kernel_time(ctx, arm, steps, used):
    """ms per frame-kernel launch.  Workloads that also compact the live-stick list every frame time the same kind of
    frames again without the compaction launches."""
    if not arm.compact:
        return None
    keep, arm.compact = arm.compact, False
    ms = [ctx.timed(lambda: arm.issue(used + i * steps, used + (i + 1) * steps)) for i in range(3)]
    arm.compact = keep
    return sorted(ms)[1] / steps


def e2e_counters(ctx, arm, steps, start):
    """End to end through the C ABI with HOST buffers: per step the host frame struct goes in (128 B) and the counters
    come back (64 B, pinned staging inside the library, synchronises) — the call pattern of a headless host loop."""
    state = {"last": None, "live": 0}

    def block(a):
        fr = arm.frames(start + a, start + a + steps) if arm.trace != "per_cloth" else None

        def go():
            for i in range(steps):
                if arm.trace == "per_cloth":
                    arm.cloth.step_each(arm._pc[(a + i) % len(arm._pc)])
                else:
                    arm.cloth.step(fr[i])
                if arm.compact:
                    state["live"] = arm.cloth.count_segments()
                state["last"] = arm.cloth.counters()
        return go
    blocks, a = [], 0
    while True:
        r0 = arm.counters()["relaxations"]
        ms = ctx.timed(block(a))
        relax = ctx.reduce(float(state["last"]["relaxations"] - r0), "SUM")
        blocks.append((ms, relax))
        a += steps
        if (sum(m for m, _ in blocks) >= 1e3 * MIN_TIMED_S and len(blocks) >= 3) or len(blocks) >= MAX_BLOCKS:
            break
    ms, relax = sorted(blocks)[len(blocks) // 2]
    frame_bytes = C.sizeof(arm.abi.Frame) * (arm.cloths if arm.trace == "per_cloth" else 1)
    return {"value": relax / (ms * 1e-3), "unit": "stick relaxations/s", "h2d_bytes_per_step": frame_bytes,
            "d2h_bytes_per_step": C.sizeof(arm.abi.Counters) + (8 if arm.compact else 0), "steps": steps, "blocks": len(blocks),
            "ms_per_step": ms / steps,
            "what": "clothgpu_step(host frame) + " + ("live-stick compaction + count (8 B D2H) + " if arm.compact else "") +
                    "clothgpu_get_counters (64 B, synchronises) per frame"}, state


def e2e_render(ctx, arm, steps, start):
    """What the cgo shim's Update does per frame (host/go/physics/clothgpu_cgo.go, physics/cloth.go:85-138): clothgpu_step
    with the host frame, then clothgpu_read_quads_f32 — live-stick compaction, addSegment / normal on the device, and the
    outline quads + highlight bytes copied into (pinned) host memory.  One cloth, rank 0."""
    torch = ctx.torch
    cap = 2 * arm.cloth.nx * arm.cloth.rows
    quads = torch.empty((cap, 8), dtype=torch.float32, pin_memory=True)
    hl = torch.empty((cap,), dtype=torch.uint8, pin_memory=True)
    state = {"n": 0}

    def block(a):
        fr = arm.frames(start + a, start + a + steps)

        def go():
            for i in range(steps):
                arm.cloth.step(fr[i])
                state["n"] = arm.cloth.read_quads_into(quads.data_ptr(), hl.data_ptr(), cap)
        return go
    blocks, a = [], 0
    while True:
        ms = ctx.timed(block(a))
        blocks.append(ms)
        a += steps
        if (sum(blocks) >= 1e3 * MIN_TIMED_S and len(blocks) >= 3) or len(blocks) >= MAX_BLOCKS:
            break
    ms = sorted(blocks)[len(blocks) // 2]
    return {"ms_per_frame": ms / steps, "frames_per_s": steps / (ms * 1e-3), "h2d_bytes_per_step": C.sizeof(arm.abi.Frame),
            "d2h_bytes_per_step": int(state["n"]) * 33 + 8, "live_sticks": int(state["n"]), "blocks": len(blocks),
            "what": "clothgpu_step(host frame) + clothgpu_read_quads_f32 into pinned host memory (32 B of outline quad + 1 highlight byte per live stick) per frame; lists of at most 2 MB are stored there by the export kernel itself, longer ones copied"}


def strip_parity(ctx):
    """N > 1: a 2048 x 2048 cloth, 20 frames of the tearing drag (it crosses every strip boundary), as N row strips on the
    N GPUs and as the whole cloth on rank 0; all five planes must agree bit for bit."""
    import numpy as np

    import clothgpu
    from clothgpu import _abi
    from clothgpu.strips import StripCloth
    nx = ny = 2048
    frames = tear_sweep_trace(20, nx, ny, 1, period=19)   # one traverse: crosses every strip boundary
    s = StripCloth(nx, ny, max_iterations=1, device=ctx.local)
    whole = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, max_iterations=1, device=ctx.local)) if ctx.rank == 0 else None
    s.step_n(frames)
    got = s.gather_state()
    cs = s.counters()
    verdict = "bit-exact"
    if ctx.rank == 0:
        whole.step_n(frames)
        ref = whole.read_state()
        cw = whole.counters()
        for name, a, b in zip(("x", "y", "px", "py", "flags"), got, ref):
            if not np.array_equal(a, b):
                bad = np.flatnonzero(a != b)
                verdict = f"MISMATCH: plane {name} differs at {bad.size} particles, first rows {np.unique(bad // nx)[:6].tolist()}"
                break
        if verdict == "bit-exact":
            for k in ("relaxations", "alive_sticks", "live_sticks", "inactive", "pinned"):
                if cs[k] != cw[k]:
                    verdict = f"MISMATCH: counter {k}: strips {cs[k]} != whole {cw[k]}"
        torn = cw["alive_sticks"]
        whole.close()
    s.close()
    ok = ctx.reduce(1.0 if verdict == "bit-exact" else 0.0, "MIN")
    if ok < 1.0:
        if ctx.rank == 0:
            print(json.dumps({"error": "strip parity check failed", "strip_parity": verdict}), flush=True)
        raise SystemExit(3)
    return {"strip_parity": "bit-exact", "strip_parity_what": f"2048x2048 cloth, 20 tearing-drag frames, {ctx.world} row strips over NVLink vs the whole cloth on rank 0: "
                                                               f"x, y, px, py, flags and the counters identical" + (f" ({2 * nx * ny - nx - ny - torn} sticks torn)" if ctx.rank == 0 else "")}


def sub_workload(ctx, name, steps, warmup):
    """One entry of the `workloads` object: value / ms_per_step / roofline.frac / kernel (+ what is specific to it).
    In a multi-rank run only the workloads that shard (strips, ensembles) are run, on every rank."""
    name, _, prec = name.partition("@")
    prec = prec or "f64"
    nx, ny, K, cloths, pin, trace, desc_txt = WORKLOADS[name]
    arm = Arm(ctx, name, precision=prec)
    out = {"workload": desc_txt + (" [optional binary32 mode: state and arithmetic in f32]" if prec == "f32" else ""), "n_gpus": arm.world,
           "dtype": prec}
    if name == "c0":
        # the whole 1000-frame scripted trace from the rest state per block; Reset (cloth.go:142-148) between blocks
        blocks = []
        while True:
            arm.cloth.reset(128, 164)
            arm.cloth.sync()
            r0 = arm.counters()["relaxations"]
            ms = ctx.timed(lambda: arm.issue(0, 1000))
            blocks.append((ms, float(arm.counters()["relaxations"] - r0)))
            if (sum(b[0] for b in blocks) >= 1e3 * MIN_TIMED_S and len(blocks) >= 3) or len(blocks) >= MAX_BLOCKS:
                break
        ms, relax = sorted(blocks)[len(blocks) // 2]
        m = {"value": relax / (ms * 1e-3), "ms_per_step": ms / 1000, "frames_used": 0,
             "timing": {"blocks": len(blocks), "steps_per_block": 1000, "timed_s": sum(b[0] for b in blocks) * 1e-3}}
    else:
        m = measure(ctx, arm, steps, warmup)
    used = m["frames_used"]
    out.update({"value": m["value"], "unit": "stick relaxations/s", "ms_per_step": m["ms_per_step"], "timing": m["timing"]})
    out["frames_per_sec"] = (arm.cloths * (1 if arm.strips else arm.world)) / (m["ms_per_step"] * 1e-3)
    kernel = arm.cloth.last_schedule()[1]
    kms = frame_kernel_time(ctx, arm, steps, used) if arm.compact else m["ms_per_step"]
    launches_per_frame = max(1, -(-K // 8)) if kernel.startswith("fused") else 1
    rf = roofline_of(arm, kms / launches_per_frame, kernel, name if prec == "f64" else None)
    out["kernel"] = kernel
    out["roofline"] = {k: rf[k] for k in ("bound", "achieved", "peak", "unit", "frac", "traffic", "traffic_is", "algorithmic_bytes_per_launch")}
    if prec == "f32":
        out["note"] = ("binary32 mode: same kernel, 8-byte column pairs per lane, its own branch-free sqrt / division expansion (proven exact "
                       "against sqrt.rn.f32 / div.rn.f32 on every operand by clothgpu_selftest_stickmath_f32); as many instructions per "
                       "particle as binary64 for half the bytes, so it is bound by issue slots rather than by HBM (33.1 B/particle)")
    if arm.compact:
        out["frame_kernel_ms"] = kms
        out["compaction_ms"] = m["ms_per_step"] - kms
    if K > 1:
        out["fp64_pipe"] = {"frac": (m["value"] / arm.world) * FP64_INSTR_PER_RELAX / FP64_PIPE_PEAK,
                            "what": f"relaxations/s per GPU x {FP64_INSTR_PER_RELAX:.0f} binary64 instructions per stretched stick / {FP64_PIPE_PEAK:.1e} "
                                    f"thread-instructions/s (tools/pipe_peaks.cu): the first roof for K >= 3, shown beside the HBM fraction"}
    if name == "c0":
        # the same trace as ONE launch of the resident CTA (CLOTHGPU_SCHED_RESIDENT, csrc/cloth_small.cuh)
        arm.cloth.set_schedule(arm.abi.SCHED_RESIDENT)
        ms_res = []
        for _ in range(5):
            arm.cloth.reset(128, 164)
            arm.cloth.sync()
            ms_res.append(ctx.timed(lambda: arm.issue(0, 1000)))
        arm.cloth.set_schedule(arm.abi.SCHED_AUTO)
        out["resident_single_launch_ms_per_step"] = sorted(ms_res)[2] / 1000
    if name in ("c0", "c2t"):
        if name == "c0":
            arm.cloth.reset(128, 164)
        out["e2e_render"] = e2e_render(ctx, arm, 50 if name == "c0" else 3, 100 if name == "c0" else used)
    if name == "c0":
        cpu = cpu_run("c0", steps=1000, warmup=0)
        out["cpu"] = {k: cpu[k] for k in ("value", "frames_per_s", "ms_per_frame", "cores", "kind", "sample")}
        out["side_by_side_ms_per_frame"] = {"cpu_port_physics_only": cpu["ms_per_frame"], "gpu_device": m["ms_per_step"],
                                            "gpu_device_resident_single_launch": out["resident_single_launch_ms_per_step"],
                                            "gpu_e2e_render": out["e2e_render"]["ms_per_frame"]}
    arm.close()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="clothgpu", choices=["clothgpu", "reference"])
    ap.add_argument("--workload", default="c4t", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--schedule", default="auto", choices=["auto", "fused", "rowstream", "wave", "streaming"],
                    help="auto = what clothgpu_step picks itself")
    ap.add_argument("--iterations", type=int, default=0, help="override the workload's sweeps per frame (tuning)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-subs", action="store_true", help="headline workload only (no `workloads` object)")
    ap.add_argument("--subs", default=",".join(SUBS), help="comma-separated sub-workloads of the default run")
    args = ap.parse_args()
    if args.impl == "clothgpu":
        args.warmup = max(args.warmup, 3)
    if args.iterations > 0:
        w_ = list(WORKLOADS[args.workload])
        w_[2] = args.iterations
        w_[6] += f" [sweeps overridden: {args.iterations}]"
        WORKLOADS[args.workload] = tuple(w_)

    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return

    wl = args.workload
    # The CPU leg runs first, before the process group exists: the other ranks then wait in the rendezvous (asleep)
    # instead of spinning in an NCCL barrier on their GPUs.
    cpu = None
    if rank == 0 and not args.no_cpu:
        cpu = cpu_run(wl, seconds_target=args.cpu_seconds)
        cpu = {k: cpu[k] for k in ("value", "unit", "cores", "kind", "sample", "frames_per_s")}

    ctx = Ctx(args)
    ctx.start()
    from clothgpu import _abi
    nx, ny, K, cloths, pin, trace, desc_txt = WORKLOADS[wl]
    extra = {}
    if cloths == 0 and ctx.world > 1:
        extra.update(strip_parity(ctx))

    arm = Arm(ctx, wl, precision=args.precision, schedule=args.schedule)
    with ClockSampler(ctx.local) as clk:
        m = measure(ctx, arm, args.steps, args.warmup)
        e2e, st = e2e_counters(ctx, arm, args.steps, m["frames_used"])
    sched, kernel = arm.cloth.last_schedule()
    kms = frame_kernel_time(ctx, arm, args.steps, m["frames_used"]) or m["ms_per_step"]
    launches_per_frame = max(1, -(-K // 8)) if kernel.startswith("fused") else 1
    roofline = roofline_of(arm, kms / launches_per_frame, kernel, wl if args.precision == "f64" and args.schedule == "auto" else None)
    if K > 1:
        roofline["fp64_pipe_frac"] = (m["value"] / ctx.world) * FP64_INSTR_PER_RELAX / FP64_PIPE_PEAK
    last = st["last"] or {}
    if arm.trace != "idle":
        extra.update({"alive_sticks_end": ctx.reduce(float(last.get("alive_sticks", 0)), "SUM"),
                      "torn_last_frame": ctx.reduce(float(last.get("torn_last", 0)), "SUM"),
                      "sticks_at_init": 2 * nx * ny - nx - ny})
    line = {
        "metric": "stick_relaxations_per_sec", "value": m["value"], "unit": "stick relaxations/s", "n_gpus": ctx.world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": m["ms_per_step"], "higher_is_better": True,
        "scaling": "strong" if arm.strips else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": config_for(wl),
        "run": {"cloths_per_gpu": (1.0 / ctx.world) if arm.strips else arm.cloths, "schedule": args.schedule,
                   "l2": f"{arm.particles * arm.bpp / 2 / 1e6:.0f} MB of state per GPU, re-streamed every frame",
                   "sharding": ("row strips, ghost rows 2 x sweeps deep; the frame kernel stores its edge rows into the neighbours' memory (NVLink peer "
                                "stores) and hand-shakes per edge chunk (ld.acquire.sys / st.release.sys epoch flags): 1 launch per frame, no NCCL on the data path")
                               if arm.strips else "independent cloths per GPU, no data-path collective"},
        "timing": m["timing"],
        "frames_per_sec": (1 if arm.strips else ctx.world * arm.cloths) / (m["ms_per_step"] * 1e-3),
        "gpu_launches": m["gpu_launches"],
        "e2e": e2e, "roofline": roofline,
    }
    line.update(extra)
    arm.close()

    if wl == "c4t" and not args.no_subs:
        subs = {}
        for name in [s for s in args.subs.split(",") if s]:
            if ctx.world > 1 and name not in ("c4", "c3", "c4t@f32"):
                continue        # single-GPU configurations are reported by the N = 1 run
            subs[name] = sub_workload(ctx, name, args.steps, args.warmup)
        line["workloads"] = subs

    if ctx.rank == 0:
        line["clocks"] = clk.summary()
        if cpu is not None:
            line["cpu_baseline"] = cpu
        print(json.dumps(line), flush=True)
    if ctx.world > 1:
        ctx.dist.barrier()
        ctx.dist.destroy_process_group()


if __name__ == "__main__":
    main()
