/*
 * clothgpu.h — C ABI of libclothgpu.so, the B200 (sm_100a) drop-in for the physics half of the
 * reference's per-frame cloth update.
 *
 * The reference (esimov/cloth-physics) is pure Go and has no FFI boundary today; its GUI calls the Go
 * methods of package `physics` directly.  This header is the boundary a cgo shim for that package binds
 * (INTEGRATION.md shows the shim).  Every entry point cites the reference interface it replaces as
 * `file:line` relative to the reference tree.
 *
 * Conventions
 *   - plain pointers and sizes only; no C++/torch types cross this boundary.
 *   - every function returns 0 (CLOTHGPU_OK) or a negative clothgpu_status; nothing aborts the process.
 *     `clothgpu_last_error()` returns a thread-local description of the last failure.
 *   - a handle is used by one host thread at a time (the reference runs its physics on one goroutine,
 *     main.go:87-100).  Work is queued on the handle's CUDA stream; `clothgpu_step*` are asynchronous,
 *     the `clothgpu_read_*` / `clothgpu_counters` calls synchronise.
 *   - there is NO CPU fallback: if no CUDA device is usable, `clothgpu_create` fails.
 */
#ifndef CLOTHGPU_H
#define CLOTHGPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define CLOTHGPU_API __attribute__((visibility("default")))
#else
#define CLOTHGPU_API
#endif

#define CLOTHGPU_VERSION_MAJOR 0
#define CLOTHGPU_VERSION_MINOR 1

typedef struct clothgpu clothgpu; /* opaque; owns all device memory of one cloth (or ensemble / strip) */

typedef enum clothgpu_status {
    CLOTHGPU_OK = 0,
    CLOTHGPU_ERR_INVALID = -1, /* bad argument / descriptor                                   */
    CLOTHGPU_ERR_CUDA = -2,    /* a CUDA runtime call failed; see clothgpu_last_error()        */
    CLOTHGPU_ERR_NODEVICE = -3,/* no usable sm_100 device: there is no CPU fallback            */
    CLOTHGPU_ERR_NOMEM = -4,
    CLOTHGPU_ERR_STATE = -5,   /* call not valid in the handle's current state                 */
    CLOTHGPU_ERR_CAPACITY = -6 /* caller buffer too small                                      */
} clothgpu_status;

/* Which particles Cloth.Init pins. */
typedef enum clothgpu_pin_rule {
    CLOTHGPU_PIN_REFERENCE = 0, /* row 0, x % (clothX/10) == 0            (physics/cloth.go:72-75)  */
    CLOTHGPU_PIN_TOP_ROW = 1,   /* whole row 0 (BASELINE.json configs[1])                           */
    CLOTHGPU_PIN_NONE = 2
} clothgpu_pin_rule;

typedef enum clothgpu_precision {
    CLOTHGPU_F64 = 0, /* parity mode: IEEE binary64, no FMA contraction (matches Go/amd64)           */
    CLOTHGPU_F32 = 1  /* optional throughput mode; state and arithmetic in binary32                  */
} clothgpu_precision;

/* Kernel schedule used by clothgpu_step.  All schedules produce bit-identical results. */
typedef enum clothgpu_schedule {
    CLOTHGPU_SCHED_AUTO = 0,      /* ROWSTREAM for one-sweep frames of cloths large enough to feed it, else FUSED */
    CLOTHGPU_SCHED_STREAMING = 1, /* one Verlet launch + one launch per colour pass (4*K), in place   */
    CLOTHGPU_SCHED_FUSED = 2,     /* one launch per frame: shared-memory tiles with 2*K halos        */
    CLOTHGPU_SCHED_ROWSTREAM = 3, /* one launch per frame, iterations == 1 only: every warp streams a 64-column window
                                     down the rows in registers (frames with more sweeps fall back to FUSED) */
    CLOTHGPU_SCHED_WAVE = 4,      /* one launch per frame, 2 <= iterations <= 8: rows stream through a shared-memory ring,
                                     the sweeps pipelined along the rows (other sweep counts fall back to FUSED) */
    CLOTHGPU_SCHED_RESIDENT = 5   /* GUI-sized cloths (one cloth of up to ~7000 particles in binary64): ONE resident CTA keeps
                                     the whole cloth in shared memory and steps all n frames of a clothgpu_step_n call in ONE
                                     launch (no launch latency between frames, but one SM's arithmetic: measured 14 us per
                                     frame on the default cloth against 8 us for a multi-CTA launch per frame, so AUTO does not
                                     pick it).  Larger cloths and ensembles fall back to AUTO. */
} clothgpu_schedule;

/* Bits of the per-particle flag byte (also the layout returned by clothgpu_read_flags). */
enum {
    CLOTHGPU_FLAG_PIN = 1u << 0,       /* particle.pinX        physics/particle.go:23               */
    CLOTHGPU_FLAG_ACTIVE = 1u << 1,    /* particle.isActive    physics/particle.go:24               */
    CLOTHGPU_FLAG_HIGHLIGHT = 1u << 2, /* particle.highlighted physics/particle.go:25               */
    CLOTHGPU_FLAG_ALIVE_H = 1u << 3,   /* stick (x,y)-(x+1,y) is still in Cloth.constraints         */
    CLOTHGPU_FLAG_ALIVE_V = 1u << 4    /* stick (x,y)-(x,y+1) is still in Cloth.constraints         */
};

/* Bits of clothgpu_frame.buttons — the four bools of physics.Mouse (physics/mouse.go:15-18). */
enum {
    CLOTHGPU_MOUSE_LEFT = 1u << 0,
    CLOTHGPU_MOUSE_RIGHT = 1u << 1,
    CLOTHGPU_MOUSE_DRAGGING = 1u << 2,
    CLOTHGPU_MOUSE_CTRL = 1u << 3
};

/* Index of clothgpu_frame.slider[] = gui.HudSliderType (gui/hud.go:32-38).  Note the reference's enum
 * and its slider table are mis-aligned (gui/hud.go:86-92): index 2 carries the slider titled
 * "Cloth friction" (default 30) and index 3 the one titled "Cloth stiffness" (default 0.98). */
enum {
    CLOTHGPU_SLIDER_DRAG_FORCE = 0,
    CLOTHGPU_SLIDER_GRAVITY = 1,
    CLOTHGPU_SLIDER_STIFFNESS = 2,
    CLOTHGPU_SLIDER_FRICTION = 3,
    CLOTHGPU_SLIDER_TEAR_DISTANCE = 4
};

/* NewCloth(width, height, spacing, col) + Init(posX, posY, hud)      physics/cloth.go:31-38, 42-81 */
typedef struct clothgpu_desc {
    int32_t width;      /* Cloth.Width  (pixels); particles per row    nx = width /spacing + 1      */
    int32_t height;     /* Cloth.Height (pixels); particle rows        ny = height/spacing + 1      */
    int32_t spacing;    /* Cloth.spacing; every stick's rest length is (double)spacing              */
    int32_t pos_x;      /* Init posX: x of the top-left particle                                    */
    int32_t pos_y;      /* Init posY                                                                */
    int32_t pin_rule;   /* clothgpu_pin_rule                                                        */
    int32_t precision;  /* clothgpu_precision                                                       */
    int32_t device;     /* CUDA device ordinal                                                      */
    int32_t ensemble;   /* number of independent identical cloths held by this handle (>= 1)        */
    /* Row-strip decomposition of ONE cloth over several handles (one per GPU).  The handle owns the
     * global rows [strip_row0, strip_row0 + strip_rows) and keeps `ghost rows` copies of its
     * neighbours' edge rows; strip_rows == 0 means "the whole cloth" (no decomposition). */
    int32_t strip_row0;
    int32_t strip_rows;
    int32_t max_iterations; /* largest clothgpu_frame.iterations this handle will see (sizes ghosts) */
    int32_t reserved[4];
} clothgpu_desc;

/* Everything (*Cloth).Update reads besides the cloth itself: the Mouse, the five HUD sliders, the
 * window size and resize offset, dt.   physics/cloth.go:85, physics/particle.go:75-181,
 * physics/constraint.go:25-54, physics/mouse.go:9-19, gui/hud.go:86-97 */
typedef struct clothgpu_frame {
    float slider[5];       /* hud.Sliders[i].Widget.Value (float32, widened to float64 when used)   */
    float drag_force_max;  /* hud.Sliders[HudSliderDragForce].Max          physics/particle.go:97   */
    double mouse_x, mouse_y;   /* Mouse.x, Mouse.y                                                  */
    double mouse_px, mouse_py; /* Mouse.px, Mouse.py (position before the last UpdatePosition)      */
    double mouse_force;        /* Mouse.force                                                       */
    float scroll_y;            /* Mouse.scrollY (unit.Dp = float32): focus-area radius              */
    uint32_t buttons;          /* CLOTHGPU_MOUSE_* bits                                             */
    int32_t win_w, win_h;      /* gtx.Constraints.Max.X / .Y               physics/particle.go:102  */
    double win_off_x, win_off_y; /* hud.WinOffsetX / WinOffsetY            physics/particle.go:89-90 */
    double dt;                 /* consts.Delta = 0.022                     consts/consts.go:9       */
    /* Extensions; the defaults reproduce the reference's literals. */
    int32_t iterations;        /* relaxation sweeps per frame; the reference does exactly 1
                                  (physics/cloth.go:96-100)                                         */
    int32_t reserved0;
    double tear_length;        /* 150: a stick longer than this tears while dragging
                                  (physics/constraint.go:36)                                        */
    double relax_gain;         /* 0.35                                     physics/constraint.go:42 */
} clothgpu_frame;

typedef struct clothgpu_counters {
    uint64_t frames;        /* frames stepped since create/reset                                    */
    uint64_t live_sticks;   /* sticks that Cloth.Update's loop would relax: in the slice and both
                               endpoints active (physics/cloth.go:97), after the last frame        */
    uint64_t alive_sticks;  /* sticks still in Cloth.constraints (not torn)                         */
    uint64_t torn_last;     /* sticks torn by the last frame                                        */
    uint64_t pinned;        /* particles with pinX                                                  */
    uint64_t inactive;      /* particles with !isActive                                             */
    uint64_t highlighted;   /* particles highlighted by the last frame                              */
    uint64_t relaxations;   /* stick visits (constraint.Update calls) done since create/reset       */
} clothgpu_counters;

typedef struct clothgpu_info {
    int32_t nx, ny;          /* particles per row / rows of the WHOLE cloth                         */
    int32_t row0, rows;      /* rows owned by this handle (strip); rows == ny when not decomposed   */
    int32_t ghost;           /* ghost rows kept above and below the owned rows                      */
    int32_t pitch;           /* elements between consecutive rows of the state planes               */
    int32_t ensemble;
    int32_t precision;
    int64_t particles;       /* owned particles per cloth = nx * rows                               */
    int64_t sticks;          /* sticks of the whole cloth at Init = 2*nx*ny - nx - ny               */
    int32_t device;
    int32_t sm_count;
} clothgpu_info;

/* Fill `f` with the reference's defaults: HUD slider defaults (gui/hud.go:86-92), dt = consts.Delta,
 * window consts.WindowSizeX x WindowSizeY, scroll_y = consts.DefaultFocusArea, mouse idle at (0,0),
 * iterations = 1, tear_length = 150, relax_gain = 0.35. */
CLOTHGPU_API int clothgpu_frame_default(clothgpu_frame* f);

/* Fill `d` with the reference's default GUI cloth: 1024 x 211 px, spacing 6, origin (128,164),
 * reference pin rule (main.go:149-166). */
CLOTHGPU_API int clothgpu_desc_default(clothgpu_desc* d);

/* NewCloth + Init                                                   physics/cloth.go:31-38, 42-81 */
CLOTHGPU_API int clothgpu_create(const clothgpu_desc* desc, clothgpu** out);

/* (*Cloth).Reset(startX, startY, hud): back to the rest grid, all sticks alive, Init pins only.
 *                                                                      physics/cloth.go:142-148   */
CLOTHGPU_API int clothgpu_reset(clothgpu* c, int32_t pos_x, int32_t pos_y);

/* Explicit lifetime (Go side: runtime.SetFinalizer + Close). */
CLOTHGPU_API void clothgpu_destroy(clothgpu* c);

/* The physics half of (*Cloth).Update: all particle updates, then `iterations` colour-ordered sweeps
 * over all live sticks with tearing.      physics/cloth.go:92-100, particle.go:75-181,
 * constraint.go:25-64.  Asynchronous on the handle's stream. */
CLOTHGPU_API int clothgpu_step(clothgpu* c, const clothgpu_frame* frame);

/* n consecutive frames (headless loop); frames[i] is frame i: one frame-kernel launch per frame, or — on request, for
 * cloths that fit the shared memory of one SM — all n frames in a single launch (CLOTHGPU_SCHED_RESIDENT). */
CLOTHGPU_API int clothgpu_step_n(clothgpu* c, const clothgpu_frame* frames, int32_t n);

/* Ensemble handles only: frames[e] drives cloth e (e < ensemble) for one frame.  The frames may differ in everything the
 * particle loop reads (mouse, buttons, sliders, window, dt); `iterations`, `tear_length` and `relax_gain` must be the
 * same for all cloths (CLOTHGPU_ERR_INVALID otherwise). */
CLOTHGPU_API int clothgpu_step_each(clothgpu* c, const clothgpu_frame* frames, int32_t count);

/* Block until everything queued on the handle has finished. */
CLOTHGPU_API int clothgpu_sync(clothgpu* c);

CLOTHGPU_API int clothgpu_get_info(clothgpu* c, clothgpu_info* info);
CLOTHGPU_API int clothgpu_set_schedule(clothgpu* c, int32_t schedule);

/* Copy the owned particles of cloth `cloth` to host, row-major, nx per row, no padding.  Any pointer
 * may be NULL.  In F32 handles values are widened to double.  `flags` gets one CLOTHGPU_FLAG_* byte
 * per particle.  This is particle{x,y,px,py,pinX,isActive,highlighted} (physics/particle.go:16-27)
 * plus membership of the two sticks the particle heads in Cloth.constraints. */
CLOTHGPU_API int clothgpu_read_state(clothgpu* c, int32_t cloth, double* x, double* y, double* px, double* py,
                        uint8_t* flags);

/* Inverse of clothgpu_read_state (restore a snapshot / seed a test state).  NULL planes are kept.
 * Coordinates must be finite (|v| <= 1e30), else CLOTHGPU_ERR_INVALID; -0.0 is stored as +0.0 (the update never
 * produces -0.0 itself).  clothgpu_step rejects non-finite frame values the same way — the Go code would
 * propagate NaN instead. */
CLOTHGPU_API int clothgpu_write_state(clothgpu* c, int32_t cloth, const double* x, const double* y,
                         const double* px, const double* py, const uint8_t* flags);

/* On-disk snapshot of everything the handle owns (no reference counterpart: the reference's only state transition
 * besides Update is Reset, physics/cloth.go:142-148).  One little-endian file:
 *     char     magic[8] = "CLSTATE1"
 *     uint32   header_bytes = 64, version = 1
 *     int32    nx, ny, row0, rows        particles per row / rows of the whole cloth; the owned rows stored here
 *     int32    ensemble, spacing, pos_x, pos_y
 *     uint64   frames                    frames stepped since create / reset
 *     uint64   reserved = 0
 *     per cloth: double x[rows*nx], y[], px[], py[]; uint8 flags[rows*nx] (CLOTHGPU_FLAG_* bits), zero-padded to 8 bytes
 *     uint64   FNV-1a 64 of every byte before it
 * i.e. particle{x,y,px,py,pinX,isActive,highlighted} (physics/particle.go:16-27) plus stick membership, in the row-major
 * order of Cloth.particles.  F32 handles store their values widened to double.  clothgpu_save_state synchronises;
 * clothgpu_load_state needs a handle of the same geometry (nx, ny, owned rows, ensemble), validates the checksum and the
 * values as clothgpu_write_state does, and restores the frame count; attached strips call clothgpu_strip_push after it.
 * host/go/physics/dump_harness.go writes the same file from the unmodified reference. */
CLOTHGPU_API int clothgpu_save_state(clothgpu* c, const char* path);
CLOTHGPU_API int clothgpu_load_state(clothgpu* c, const char* path);

/* Bit-planes of the flag byte, 1 bit per particle, LSB first, ceil(particles/32) words each; any
 * pointer may be NULL. */
CLOTHGPU_API int clothgpu_read_masks(clothgpu* c, int32_t cloth, uint32_t* pin, uint32_t* active,
                        uint32_t* highlighted, uint32_t* alive_h, uint32_t* alive_v);

/* The live-stick list the reference's render loops walk (physics/cloth.go:108-114, 126-134):
 * for every stick still in the slice whose endpoints are both active, in the order Cloth.Init
 * created them (physics/cloth.go:61-70: by tail particle in row-major order, the vertical stick
 * from the particle above before the horizontal stick from the particle to the left), the
 * float32 endpoints {x1,y1,x2,y2} and a byte that is 1 when both endpoints are highlighted.
 * `cap` = capacity in sticks; *n = sticks written.  `sticks_idx` (optional) gets {i1,i2} particle
 * indices. */
CLOTHGPU_API int clothgpu_read_segments_f32(clothgpu* c, int32_t cloth, float* xyxy, uint8_t* highlighted,
                               uint32_t* sticks_idx, size_t cap, size_t* n);

/* The outline geometry the reference's render loops hand to Gio (physics/cloth.go:108-114 -> addSegment :150-157 ->
 * normal :160-168): for every live stick, in the order of clothgpu_read_segments_f32, the four float32 corners
 * {a+n, b+n, b-n, a-n} (8 floats: MoveTo, LineTo, LineTo, LineTo) where n is the stick's normal scaled to `line_width`
 * (the reference's lineWidth is 0.6, physics/cloth.go:16; a stick shorter than 1e-5 gets n = 0 as in normal()).
 * `highlighted` as in clothgpu_read_segments_f32 (the second render loop, :126-134, draws that subset).  `cap` = capacity
 * in sticks; *n = live sticks; with both output pointers NULL only the count is returned. */
CLOTHGPU_API int clothgpu_read_quads_f32(clothgpu* c, int32_t cloth, float line_width, float* quads,
                            uint8_t* highlighted, size_t cap, size_t* n);

/* Page-locked host memory for the two calls above (32-byte aligned; no reference counterpart — Go slices are pageable).
 * When `xyxy` / `quads`, `highlighted` and `sticks_idx` all lie in such memory (this allocator's, cudaHostAlloc'ed or
 * cudaHostRegister'ed) and the list is at most 2 MB, the export kernel stores the list there itself and the call is one
 * launch and one synchronisation — no device buffer, no copies.  Any other destination works too, through copies.
 * clothgpu_host_free(NULL) is a no-op. */
CLOTHGPU_API int clothgpu_host_alloc(size_t bytes, void** p);
CLOTHGPU_API int clothgpu_host_free(void* p);

/* Only the device half of the above: build the compacted live-stick list of cloth `cloth` in device memory,
 * asynchronously on the handle's stream (the next clothgpu_read_segments_f32 re-runs it; with all three output
 * pointers NULL that call returns just the count).  Lets a renderer that consumes the list on the GPU, or a benchmark
 * of the tear + compaction path (BASELINE.json configs[2]), avoid the host copy. */
CLOTHGPU_API int clothgpu_compact_segments(clothgpu* c, int32_t cloth);

/* Counters after the last queued frame (synchronises). */
CLOTHGPU_API int clothgpu_get_counters(clothgpu* c, clothgpu_counters* out);

/* Strip handles: device pointers/sizes for the halo exchange.
 * plane 0..3 = x,y,px,py, 4 = flags.  `which`: 0 = my top edge rows (to send up), 1 = my bottom
 * edge rows (to send down), 2 = my top ghost rows (to receive), 3 = my bottom ghost rows. */
CLOTHGPU_API int clothgpu_halo_region(clothgpu* c, int32_t plane, int32_t which, void** dev_ptr, size_t* bytes);

/* ---- one cloth over several GPUs: row strips, one process (or handle) per GPU -------------------------------
 * No reference counterpart (the reference is single-threaded, main.go:87-100).  Each strip handle is created with
 * desc.strip_row0/strip_rows; it exports a small blob describing its device memory (clothgpu_strip_export), the
 * host program passes the blobs of the upper / lower neighbour to clothgpu_strip_attach (any transport: the Python
 * harness uses torch.distributed all_gather), and from then on clothgpu_step's fused kernel stores the edge rows a
 * neighbour keeps as ghosts straight into the neighbour's memory (NVLink peer stores) and hand-shakes with
 * stream-ordered epoch flags — no host round trip, no separate exchange pass.  All strips must issue the same
 * sequence of clothgpu_step / clothgpu_reset / clothgpu_strip_push calls (they index each other's double buffers).
 * Blobs from the same process are attached by plain pointer, from other processes through cudaIpcOpenMemHandle. */
CLOTHGPU_API int clothgpu_strip_export(clothgpu* c, void* blob, size_t cap, size_t* len); /* blob==NULL: size query */
CLOTHGPU_API int clothgpu_strip_attach(clothgpu* c, const void* upper_blob, const void* lower_blob); /* NULL = cloth edge */
/* After clothgpu_write_state on attached strips: copy my edge rows into the neighbours' ghost rows. */
CLOTHGPU_API int clothgpu_strip_push(clothgpu* c);

/* Use an existing CUDA stream (cudaStream_t passed as void*) for all work of this handle. */
CLOTHGPU_API int clothgpu_set_stream(clothgpu* c, void* cuda_stream);

/* Milliseconds of device time spent in the kernels of the last clothgpu_step / step_n call
 * (CUDA events on the handle's stream); synchronises. */
CLOTHGPU_API int clothgpu_last_step_ms(clothgpu* c, float* ms);

/* Number of kernel launches issued by this handle since create. */
CLOTHGPU_API int clothgpu_launch_count(clothgpu* c, uint64_t* launches);

/* The schedule the last clothgpu_step / step_n / step_each actually ran (AUTO and the fall-backs resolved): a
 * clothgpu_schedule value, and the name of the frame kernel (static string, e.g. "stream1_frame_kernel<double,DRAG>").
 * Either pointer may be NULL.  CLOTHGPU_ERR_STATE before the first step. */
CLOTHGPU_API int clothgpu_last_schedule(clothgpu* c, int32_t* schedule, const char** kernel_name);

/* Device self-test of the binary64 stick arithmetic (test hook; no reference counterpart).  Every fast kernel derives
 * sqrt(s), (L - d)/d and L/d of physics/constraint.go:28,41-42 from one reciprocal-square-root seed; this runs that
 * expansion against the device's own IEEE sqrt.rn.f64 / div.rn.f64 on `samples` pseudo-random stretched sticks (rest
 * length 1..64, squared length up to 2^61, random mantissas; seed selects the sequence) and on every squared length of
 * `s_list` (n_list entries, may be NULL) paired with every rest length 1..64 that leaves it stretched.  The report
 * counts the compared pairs and the mismatches (0 expected) and holds the first mismatching operands. */
typedef struct clothgpu_stickmath_report {
    uint64_t tested;
    uint64_t mismatches;
    double first_s, first_L;
    double got_dist, want_dist, got_mul, want_mul;
} clothgpu_stickmath_report;
CLOTHGPU_API int clothgpu_selftest_stickmath(int32_t device, uint64_t seed, uint64_t samples, const double* s_list,
                                size_t n_list, clothgpu_stickmath_report* report);
/* The binary32 mode's expansion, EXHAUSTIVELY: every binary32 squared length in [1, 2^63) with every rest length 1..64
 * that leaves the stick stretched (2.9e10 pairs, about a second) against sqrt.rn.f32 / div.rn.f32. */
CLOTHGPU_API int clothgpu_selftest_stickmath_f32(int32_t device, clothgpu_stickmath_report* report);

CLOTHGPU_API const char* clothgpu_last_error(void);
CLOTHGPU_API const char* clothgpu_version(void);

#ifdef __cplusplus
}
#endif
#endif /* CLOTHGPU_H */
