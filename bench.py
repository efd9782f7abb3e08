#!/usr/bin/env python
"""bench.py — throughput of the per-frame cloth update (stick relaxations/s) on B200.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload c1|c0|c2]

One "step" = one frame (particle loop + `iterations` colour-ordered relaxation sweeps with tearing) over the
rank's batch of cloths.  Default workload = BASELINE.json configs[1]: 1024x1024 particles, 8 sweeps/frame, top
row pinned, binary64 — held as a batch of 8 independent cloths per GPU so the working set (277 MB) exceeds the
126 MB L2 and every frame streams its state from HBM.  Prints ONE JSON line (see README/DESIGN.md).
"""
import argparse
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

WORKLOADS = {
    # name: (nx, ny, iterations, cloths per GPU, pin rule, description)
    "c1": (1024, 1024, 8, 8, "top_row", "BASELINE configs[1]: 1024x1024 cloth, 8 sweeps/frame, top row pinned, f64"),
    "c1k1": (1024, 1024, 1, 8, "top_row", "1024x1024 cloth, 1 sweep/frame (reference semantics), top row pinned, f64"),
    "c0": (171, 36, 1, 1, "reference", "BASELINE configs[0]: reference default GUI cloth 171x36, 1 sweep/frame"),
    "c2": (4096, 4096, 8, 1, "top_row", "BASELINE configs[2] geometry: 4096x4096 cloth, 8 sweeps/frame, f64"),
    "c2k1": (4096, 4096, 1, 1, "top_row", "4096x4096 cloth, 1 sweep/frame (reference semantics), f64"),
    "c3": (128, 128, 8, 2048, "top_row", "BASELINE configs[3] shard: 2048 independent 128x128 cloths per GPU"),
    # BASELINE configs[2]: scripted drag / tear / hole-punch trace, live-stick compaction after every frame
    "c2t": (4096, 4096, 1, 1, "top_row", "BASELINE configs[2]: 4096x4096 cloth, scripted drag/tear/hole-punch trace, live-stick compaction every frame, 1 sweep/frame, f64"),
    "c2t8": (4096, 4096, 8, 1, "top_row", "BASELINE configs[2]: 4096x4096 cloth, scripted drag/tear/hole-punch trace, live-stick compaction every frame, 8 sweeps/frame, f64"),
    # one cloth cut into row strips over the ranks (strong scaling; cloths-per-GPU column = 0 marks the mode)
    "c4": (16384, 16384, 8, 0, "top_row", "BASELINE configs[4]: one 16384x16384 cloth in row strips over the GPUs, 8 sweeps/frame, f64"),
    "c4k1": (16384, 16384, 1, 0, "top_row", "one 16384x16384 cloth in row strips over the GPUs, 1 sweep/frame (reference semantics), f64"),
    "c4s": (4096, 4096, 8, 0, "top_row", "one 4096x4096 cloth in row strips over the GPUs, 8 sweeps/frame, f64"),
}


def peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_facts(workload):
    """Per-launch DRAM traffic and f64-pipe utilisation of the dominant kernel, from the committed ncu --set full captures
    (profiles/ncu_facts.json; the capture file is named there).  Not measured live: ncu cannot run inside a bench."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_facts.json")) as f:
            return json.load(f).get(workload) or {}
    except Exception:
        return {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region."""
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
            self._stop.wait(0.1)

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


def make_desc(_abi, wl, device, precision):
    nx, ny, K, cloths, pin, _ = WORKLOADS[wl]
    rule = {"top_row": _abi.PIN_TOP_ROW, "reference": _abi.PIN_REFERENCE}[pin]
    if wl == "c0":
        d = _abi.default_desc()
        d.device = device
        d.precision = precision
        return d
    return _abi.grid_desc(nx, ny, pin_rule=rule, precision=precision, ensemble=cloths, device=device, max_iterations=K)


def make_frame(_abi, wl):
    nx, ny, K, *_ = WORKLOADS[wl]
    f = _abi.default_frame(K)
    if wl != "c0":
        f.win_w, f.win_h = 128 + 6 * nx + 128, 164 + 6 * ny * 4     # window large enough to hang freely
    return f


def cpu_desc(_abi, wl, precision):
    """Descriptor of the CPU arm's sample: ONE cloth of the workload; cloths beyond 2048x2048 are sampled as a
    2048x2048 cloth of the same spacing/pin rule (per-stick cost does not depend on the size once out of cache)."""
    nx, ny, K, cloths, pin, _ = WORKLOADS[wl]
    if nx * ny > 2048 * 2048:
        nx = ny = 2048
        return _abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, precision=precision, max_iterations=K), nx, ny
    d = make_desc(_abi, wl, 0, precision)
    d.ensemble = 1
    return d, nx, ny


def cpu_sample(wl, precision, seconds_target, threads=1):
    """Time the oracle's restatement of the reference CPU path (sequential Gauss-Seidel sweep, AoS +
    pointer layout as in the Go program, 1 thread = the reference's single physics goroutine) on ONE
    cloth of the workload for a bounded number of frames."""
    import pyoracle
    from clothgpu import _abi
    nx, ny, K, *_ = WORKLOADS[wl]
    d, nx, ny = cpu_desc(_abi, wl, precision)
    o = pyoracle.OracleCloth(d, pyoracle.SEQ_FAITHFUL)
    f = make_frame(_abi, wl)
    o.step(f, threads)                      # warm the caches / page in
    r0, n, t0 = o.relaxations, 0, time.perf_counter()
    while True:
        o.step(f, threads)
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_target or n >= 2000:
            break
    relax = o.relaxations - r0
    return {"value": relax / dt, "unit": "stick relaxations/s", "cores": threads, "kind": "port",
            "sample": f"{n} frames of one {nx}x{ny} cloth, {K} sweeps/frame, sequential-faithful oracle "
                      f"(C restatement of physics/cloth.go:92-100; the Go reference cannot be built here), {dt:.1f} s",
            "frames_per_s": n / dt}


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (its C restatement, see
    oracle/cloth_oracle.h) on the host cores, same metric and config."""
    if rank != 0:
        return
    import pyoracle
    from clothgpu import _abi
    wl = args.workload
    nx, ny, K, cloths, pin, desc_txt = WORKLOADS[wl]
    d, nx, ny = cpu_desc(_abi, wl, _abi.F64)
    o = pyoracle.OracleCloth(d, pyoracle.SEQ_FAITHFUL)
    f = make_frame(_abi, wl)
    for _ in range(args.warmup):
        o.step(f, 1)
    r0, t0 = o.relaxations, time.perf_counter()
    for _ in range(args.steps):
        o.step(f, 1)
    dt = time.perf_counter() - t0
    relax = o.relaxations - r0
    val = relax / dt
    line = {
        "impl": "reference", "metric": "stick_relaxations_per_sec", "value": val, "unit": "stick relaxations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": desc_txt, "nx": nx, "ny": ny, "iterations": K, "pin": pin,
                   "sample": f"each step = one frame of ONE {nx}x{ny} cloth on the CPU (the GPU arm steps {cloths} per GPU)"},
        "cpu_baseline": {"value": val, "unit": "stick relaxations/s", "cores": 1, "kind": "port",
                         "sample": f"{args.steps} frames of one {nx}x{ny} cloth, sequential-faithful C restatement of the "
                                   f"reference's single-goroutine CPU path (Go toolchain absent: oracle/_ref cannot be built)"},
        "e2e": {"value": val, "unit": "stick relaxations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "frames_per_sec": args.steps / dt,
    }
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="clothgpu", choices=["clothgpu", "reference"])
    ap.add_argument("--workload", default="c1", choices=sorted(WORKLOADS))
    ap.add_argument("--precision", default="f64", choices=["f64", "f32"])
    ap.add_argument("--schedule", default="auto", choices=["auto", "fused", "rowstream", "wave", "streaming"],
                    help="auto = what clothgpu_step picks itself: the row-streaming kernel for one-sweep frames, the tile kernel otherwise")
    ap.add_argument("--iterations", type=int, default=0, help="override the workload's sweeps per frame (tuning)")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--extras", action="store_true", help="also time the K=1 and L2-resident variants")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "clothgpu" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist

    import clothgpu
    from clothgpu import _abi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device; clothgpu has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def sum_over_ranks(v):
        if world == 1:
            return v
        t = torch.tensor([v], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        return float(t.item())

    wl = args.workload
    if args.iterations > 0:
        w_ = list(WORKLOADS[wl])
        w_[2] = args.iterations
        w_[5] += f" [sweeps overridden: {args.iterations}]"
        WORKLOADS[wl] = tuple(w_)
    nx, ny, K, cloths, pin, desc_txt = WORKLOADS[wl]
    precision = _abi.F64 if args.precision == "f64" else _abi.F32
    strips = cloths == 0
    if strips:
        # one cloth, row strips: rank r owns rows [bounds[r], bounds[r+1]); edge rows travel inside the fused kernel
        from clothgpu.strips import StripCloth
        if world > 1:
            holder = StripCloth(nx, ny, max_iterations=K, precision=precision, device=local)
            cloth = holder.cloth
        else:
            cloth = clothgpu.Cloth(_abi.grid_desc(nx, ny, pin_rule=_abi.PIN_TOP_ROW, precision=precision, device=local,
                                                  max_iterations=K))
        cloths = 1
    else:
        cloth = clothgpu.Cloth(make_desc(_abi, wl, local, precision))
    cloth.set_schedule({"auto": _abi.SCHED_AUTO, "fused": _abi.SCHED_FUSED, "rowstream": _abi.SCHED_ROWSTREAM,
                        "wave": _abi.SCHED_WAVE, "streaming": _abi.SCHED_STREAMING}[args.schedule])
    stream = torch.cuda.Stream()
    cloth.set_stream(stream.cuda_stream)
    frame = make_frame(_abi, wl)
    traced = wl.startswith("c2t")
    if traced:
        # the c0 trace scaled x24 to the 4096^2 cloth: frames 380.. cover the end of the slow drag, the
        # release, the right-button hole punch with a moving focus area, the scroll and (for long runs) the fast zig-zag
        from traces import c0_trace
        tr = c0_trace(380 + args.warmup + args.steps, iterations=K, scale=24.0, win=(frame.win_w, frame.win_h))[380:]
        frames_w = clothgpu.frame_array(tr[:args.warmup])
        frames_t = clothgpu.frame_array(tr[args.warmup:])
    else:
        frames_w = clothgpu.frame_array([frame] * args.warmup)
        frames_t = clothgpu.frame_array([frame] * args.steps)

    def run_frames(arr):
        if not traced:
            cloth.step_n(arr)
            return
        for i in range(len(arr)):            # tear mask update + compaction of the live-stick list, all on the device
            cloth.step(arr[i])
            cloth.compact_segments()

    # ---- device-resident throughput: K steps back to back, state already in HBM --------------------------
    run_frames(frames_w)
    cloth.sync()
    c0 = cloth.counters()
    l0 = cloth.launch_count()
    barrier()
    with ClockSampler(local) as clk:
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        with torch.cuda.stream(stream):
            ev0.record(stream)
            run_frames(frames_t)
            ev1.record(stream)
        ev1.synchronize()
        barrier()
        ms_local = ev0.elapsed_time(ev1)
        # keep sampling for a moment if the region was very short
        if ms_local < 400:
            time.sleep(0.4)
    launches = cloth.launch_count() - l0
    c1 = cloth.counters()
    relax_local = c1["relaxations"] - c0["relaxations"]
    ms = max_over_ranks(ms_local)
    relax = sum_over_ranks(relax_local)
    value = relax / (ms * 1e-3)

    # ---- end to end through the C ABI with host buffers: per step, the host frame struct goes in and the
    # ---- counters come back (synchronous D2H); the GUI-facing call pattern of (*Cloth).Update ---------------
    n_e2e = max(10, min(args.steps, 100))
    barrier()
    with torch.cuda.stream(stream):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0 = cloth.counters()["relaxations"]
        e0.record(stream)
        for i in range(n_e2e):
            cloth.step(frames_t[i % len(frames_t)] if traced else frame)
            if traced:
                live_n = cloth.count_segments()      # compaction + the live-stick count back on the host
            last = cloth.counters()
        e1.record(stream)
    e1.synchronize()
    barrier()
    e2e_ms = max_over_ranks(e0.elapsed_time(e1))
    e2e_relax = sum_over_ranks(last["relaxations"] - r0)
    import ctypes as C
    e2e = {"value": e2e_relax / (e2e_ms * 1e-3), "unit": "stick relaxations/s",
           "h2d_bytes_per_step": C.sizeof(_abi.Frame), "d2h_bytes_per_step": C.sizeof(_abi.Counters),
           "steps": n_e2e, "ms_per_step": e2e_ms / n_e2e,
           "what": "clothgpu_step(host frame) + " + ("live-stick compaction + count (8 B D2H) + " if traced else "") +
                   "clothgpu_get_counters (sync D2H) per frame"}
    if traced:
        line_extra = {"live_sticks_end": live_n, "alive_sticks_end": last["alive_sticks"], "inactive_end": last["inactive"]}
    else:
        line_extra = {}

    # ---- roofline of the dominant kernel (fused_frame_kernel: one launch per step) -----------------------------
    peak, peak_src = peaks()
    bpp = BYTES_PER_PARTICLE_F64 if args.precision == "f64" else BYTES_PER_PARTICLE_F32
    particles = cloth.nx * cloth.rows * cloths      # per GPU (a strip owns `rows` of the ny rows)
    launches_per_step = launches / args.steps
    fused_per_step = max(1, -(-K // 8)) if args.schedule != "streaming" else 1      # handshake kernels are not frame kernels
    if traced:
        # the frame kernel alone: time the same frames again without the compaction launches
        cloth.reset(128, 164)
        cloth.step_n(frames_w)
        with torch.cuda.stream(stream):
            k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            k0.record(stream)
            cloth.step_n(frames_t)
            k1.record(stream)
        k1.synchronize()
        ms_local = k0.elapsed_time(k1)
    kernel_ms = ms_local / args.steps / fused_per_step
    alg_bytes = bpp * particles
    achieved = alg_bytes / (kernel_ms * 1e-3) / 1e9
    suffix = "<double>" if args.precision == "f64" else "<float>"
    kernel = ("stream1_frame_kernel" if (K == 1 and (args.schedule == "rowstream" or (args.schedule == "auto" and nx * ny * cloths >= 256 * 256))) else
              "wave_frame_kernel" if (args.schedule == "wave" and 2 <= K <= 8) or (args.schedule == "auto" and 4 <= K <= 8 and ((nx >= 192 and cloth.rows >= 256) or (not strips and cloths >= 2 and nx <= 128 and cloths * cloth.rows >= 256 * 148))) else
              "fused_frame_kernel" if args.schedule != "streaming" else "relax_pass_kernel") + suffix
    facts = ncu_facts(wl if args.precision == "f64" and args.schedule == "auto" else None)
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": facts.get("dram_bytes_per_launch"), "peak_source": peak_src, "kernel": kernel,
                "algorithmic_bytes_per_launch": alg_bytes,
                "fp64_pipe_active_pct": facts.get("fp64_pipe_active_pct"),
                "ncu_source": facts.get("source"),
                "note": "achieved = compulsory bytes (65.1 B/particle/frame in f64, independent of the sweep count) / launch time. "
                        "With 8 binary64 sweeps per frame the kernel is bound by the f64 pipe (fp64_pipe_active_pct, from the ncu "
                        "capture named in ncu_source), with 1 sweep by HBM; see DESIGN.md §3"}

    line = {
        "metric": "stick_relaxations_per_sec", "value": value, "unit": "stick relaxations/s", "n_gpus": world,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if strips else "weak", "vs_baseline": None, "dtype": args.precision, "data": "synthetic",
        "config": {"workload": desc_txt, "nx": nx, "ny": ny, "iterations": K, "pin": pin,
                   "cloths_per_gpu": (1.0 / world) if strips else cloths, "schedule": args.schedule,
                   "l2": f"inputs larger than L2: {particles * bpp / 2 / 1e6:.0f} MB of state per GPU, every frame re-streams it",
                   "sharding": ("row strips, 2K ghost rows, edge rows stored into the neighbours' memory by the frame kernel "
                                "(NVLink peer stores + epoch flags), no NCCL on the data path") if strips
                               else "independent cloths per GPU, no data-path collective"},
        "frames_per_sec": (1 if strips else world * cloths) * args.steps / (ms * 1e-3),
        "gpu_launches": int(launches),
        "e2e": e2e, "roofline": roofline,
    }
    line.update(line_extra)

    if rank == 0:
        line["clocks"] = clk.summary()
        if not args.no_cpu and world >= 1:
            line["cpu_baseline"] = cpu_sample(wl, precision, args.cpu_seconds)
        print(json.dumps(line), flush=True)
    if strips and world > 1:
        holder.close()
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
