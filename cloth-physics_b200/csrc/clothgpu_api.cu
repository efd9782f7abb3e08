// clothgpu_api.cu — implementation of include/clothgpu.h: handle management, per-frame scheduling of
// the sm_100a kernels, read-back.  No CPU compute path exists here: every physics result comes from
// the kernels in cloth_kernels.cuh / cloth_fused.cuh.
#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include <unistd.h>

#include "cloth_fused.cuh"
#include "cloth_kernels.cuh"
#include "cloth_selftest.cuh"
#include "cloth_small.cuh"
#include "cloth_stream1.cuh"
#include "cloth_wave.cuh"
#include "frame_fold.h"

using namespace clothgpu_impl;

namespace {

thread_local std::string g_last_error;

int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_last_error = buf;
    return code;
}

#define CU_TRY(expr)                                                                         \
    do {                                                                                     \
        cudaError_t _e = (expr);                                                             \
        if (_e != cudaSuccess)                                                               \
            return fail(CLOTHGPU_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                        __FILE__, __LINE__);                                                 \
    } while (0)

constexpr double kMaxMagnitude = 1e30;  // |frame values| and |coordinates| accepted from the host
constexpr int kMaxSweepsPerLaunch = 8;  // halo 16; larger K is split into several launches

}  // namespace

struct clothgpu {
    clothgpu_desc desc{};
    Geom g{};
    int ny = 0;
    int ghost = 0;
    int precision = CLOTHGPU_F64;
    int device = 0;
    int sm_count = 0;
    int schedule = CLOTHGPU_SCHED_AUTO;
    size_t elem = 8;
    size_t plane_elems = 0;  // per plane, all cloths
    // double-buffered planes: positions+flags flip every launch, previous positions every frame
    void* X[2] = {nullptr, nullptr};
    void* Y[2] = {nullptr, nullptr};
    void* PX[2] = {nullptr, nullptr};
    void* PY[2] = {nullptr, nullptr};
    uint8_t* FL[2] = {nullptr, nullptr};
    int cur_pos = 0, cur_prev = 0;
    DevCounters* ctr_base = nullptr;      // two slots, accumulated by the frame kernels (cloth_types.cuh): frame f uses slot f & 1
    DevCounters* ctr = nullptr;           // the slot of the frame being issued / issued last
    int ctr_slot = 0;                     // slot the NEXT frame will use
    unsigned long long* census = nullptr; // 5 values, written by count_kernel
    // Pinned staging for clothgpu_get_counters: [0] a DevCounters, then the 5 census values.
    unsigned char* ctr_host = nullptr;
    // The census (a full pass over the flag bytes) runs only when the flags changed behind the kernels' back — create,
    // reset, write_state; from then on the frame kernels keep the deltas (DevCounters) and get_counters is one 56-byte copy.
    bool census_valid = false;
    struct {
        unsigned long long live = 0, alive = 0, pinned = 0, inactive = 0, highlighted = 0;
        unsigned long long torn_total = 0;  // DevCounters::torn_total when the census was taken
        uint64_t frames = 0;                // c->frames when the census was taken
    } base;
    int last_sched = -1;                  // what the last step resolved to (clothgpu_last_schedule)
    const char* last_kernel = nullptr;
    void* frames_dev = nullptr;        // per-cloth FrameU array (ensemble input)
    void* frames_host = nullptr;       // pinned staging for the above
    cudaEvent_t frames_copied = nullptr;
    cudaStream_t stream = nullptr;
    bool own_stream = true;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    bool timed = false;
    uint64_t frames = 0;
    uint64_t launches = 0;
    int forced_tile_rows = 0;
    int fused_variant = -1;  // -1: chosen per launch (choose_variant); >= 0: forced, see launch_fused
    int sweeps_per_launch = kMaxSweepsPerLaunch;
    int k1_stream = 1;       // 1: frames with one sweep take the register-streaming kernel (cloth_stream1.cuh)
    int s1_chunk_rows = 0;   // 0: chosen per launch
    int wave_chunk_rows = 0; // 0: chosen per launch
    int wave_auto = 1;       // AUTO may pick the wavefront kernel
    // AUTO does NOT take the resident kernel by itself: measured on the default 171 x 36 cloth it needs 14.3 us per frame
    // (one SM's binary64 pipe: ~4 us of arithmetic per frame at best) against 8.2 us for one multi-CTA launch per frame.
    int small_auto = 0;
    int wave_pack = 1;       // ensembles of narrow cloths run on the wavefront kernel as one tall virtual cloth
    int s1_ring = 2;         // shared-memory prefetch ring depth of stream1_frame_kernel (0: register prefetch)
    int s1_warps_per_sm = 4 * S1_RING_CTAS;  // resident warps of stream1_frame_kernel per SM (its launch bounds)
    // All state planes and the strip epoch slots live in ONE device allocation, so a neighbouring strip (another
    // process, another GPU) maps everything with a single cudaIpcMemHandle.
    void* slab = nullptr;
    size_t slab_bytes = 0;
    size_t off_plane[10] = {0};  // X0 X1 Y0 Y1 PX0 PX1 PY0 PY1 FL0 FL1
    size_t off_slots = 0;
    unsigned int* slots = nullptr;  // [0] written by my upper neighbour, [1] by my lower neighbour
    unsigned long long* ticket = nullptr;      // [2] next to the slots: units stored towards the upper / lower neighbour
    unsigned long long ticket_goal[2] = {0, 0};  // cumulative quota of all operations issued so far
    struct Link {
        bool on = false;
        bool ipc = false;      // base came from cudaIpcOpenMemHandle (must be closed)
        char* base = nullptr;  // the neighbour's slab, mapped here
        size_t off_plane[10] = {0};
        size_t off_slots = 0;
        long long cloth_stride = 0;
        int held_row0 = 0, held_rows = 0;
        int row_lo = 0, row_hi = 0;  // my owned rows the neighbour keeps as ghosts
    } link[2];                       // 0 = upper neighbour, 1 = lower neighbour
    unsigned int epoch = 0;          // operations (fused launches, pushes) done since attach
    // scratch for segment export
    void* small_frames_dev = nullptr;   // folded FrameU per frame of a resident multi-frame launch (cloth_small.cuh)
    void* small_frames_host = nullptr;  // pinned staging for the above
    size_t small_frames_cap = 0;
    bool direct_export = true;                 // small exports into pinned caller memory: written by the kernel, no copies (CLOTHGPU_DIRECT_EXPORT=0: off)
    SegPlan* seg_plan = nullptr;               // the export's plan: one record per tile of kSegBlock particles ...
    unsigned int* seg_starts = nullptr;        // ... and where every tile's sticks start in the list
    size_t seg_plan_cap = 0;
    void* seg_out = nullptr;                   // per possible stick: 32 B (quads, or endpoints + indices) + 1 highlight byte
    size_t seg_out_cap = 0;
    unsigned long long* seg_total = nullptr;   // [0] sticks listed by the last export, [1] (as unsigned int) the plan kernel's count of finished blocks
};

namespace {

template <typename T>
Planes<T> planes(clothgpu* c, int pos, int prev) {
    Planes<T> p;
    p.x = (T*)c->X[pos];
    p.y = (T*)c->Y[pos];
    p.px = (T*)c->PX[prev];
    p.py = (T*)c->PY[prev];
    p.fl = c->FL[pos];
    return p;
}

// Tile-row choice for the fused kernel: the largest tile that fits shared memory minimises halo
// overhead, but the number of tiles should fill whole waves of SMs; cost ~ waves * tile rows.
// CTAs the variant is built to co-host per SM (its __launch_bounds__ minimum)
int variant_ctas(int variant) { return variant == 2 ? 2 : variant == 5 ? 3 : 1; }

// Which instantiation runs a launch of K sweeps.  Few sweeps: the launch is bound by HBM latency/bandwidth, so small
// tiles with 3 CTAs per SM overlap one tile's load/store with another's arithmetic.  Many sweeps: the halo (2K per
// side) must be amortised over the largest tile shared memory holds, one CTA per SM with as many warps as the
// register file allows.  CLOTHGPU_FUSED_VARIANT overrides (tuning/tests).
int choose_variant(const clothgpu* c, int K) {
    if (c->fused_variant >= 0) return c->fused_variant;
    return K <= 2 ? 5 : 4;
}

template <typename T>
FusedPlan make_plan(const clothgpu* c, int K, int do_verlet, int variant) {
    const Geom& g = c->g;
    FusedPlan pl{};
    pl.K = K;
    pl.h = 2 * K;
    pl.do_verlet = do_verlet;
    const int h = pl.h, TW = kFusedTW;
    // columns
    if (g.nx <= TW) {
        pl.tiles_x = 1;
        pl.first_w = TW;
        pl.mid_w = TW - 2 * h;
    } else {
        pl.first_w = TW - h;
        pl.mid_w = TW - 2 * h;
        pl.tiles_x = 1 + (g.nx - pl.first_w + pl.mid_w - 1) / pl.mid_w;
    }
    // rows
    const int above = g.own_row0 - g.held_row0;
    const int below = (g.held_row0 + g.held_rows) - (g.own_row0 + g.own_rows);
    pl.top_margin = std::min(h, above);
    const int bottom_margin = std::min(h, below);
    // shared memory per CTA when `ctas` CTAs are to co-reside on an SM (228 KB per SM, 1 KB reserved per CTA)
    const int ctas = variant_ctas(variant);
    const size_t smem_max = ctas == 1 ? 227 * 1024 : (size_t)(228 * 1024) / ctas - 1024;
    int th_max = 2;
    while (fused_smem_bytes<T>(th_max + 2) <= smem_max) th_max += 2;
    const int th_min = 2 * h + 2;
    if (th_max < th_min) {  // this variant's tiles cannot hold the halo of K sweeps
        pl.TH = 0;
        return pl;
    }
    auto tiles_y_for = [&](int TH, int* first_h) {
        if (pl.top_margin + g.own_rows + bottom_margin <= TH) {
            *first_h = g.own_rows;
            return 1;
        }
        *first_h = TH - pl.top_margin - h;
        const int mid = TH - 2 * h;
        return 1 + (g.own_rows - *first_h + mid - 1) / mid;
    };
    int best_th = th_max;
    if (c->forced_tile_rows >= th_min && c->forced_tile_rows <= th_max) {
        best_th = c->forced_tile_rows & ~1;
    } else {
        double best_cost = 1e300;
        for (int TH = th_max; TH >= th_min; TH -= 2) {
            int fh;
            const int ty = tiles_y_for(TH, &fh);
            const long long tiles = (long long)ty * pl.tiles_x * g.ensemble;
            const long long slots = (long long)c->sm_count * ctas;
            const long long waves = (tiles + slots - 1) / slots;
            // a tile costs ~ its rows (+ a little fixed overhead); small tiles may co-reside 2 per SM
            const double cost = (double)waves * (TH + 6);
            if (cost < best_cost - 1e-9) {
                best_cost = cost;
                best_th = TH;
            }
            if (pl.top_margin + g.own_rows + bottom_margin > TH && TH - 2 * h < 2) break;
        }
    }
    pl.TH = best_th;
    pl.tiles_y = tiles_y_for(best_th, &pl.first_h);
    pl.mid_h = best_th - 2 * h;
    return pl;
}

// The neighbour's buffers with the given indices, as the target of this operation's edge-row stores.
template <typename T>
PeerOut<T> peer_out(const clothgpu* c, int side, int pos, int prev) {
    PeerOut<T> p{};
    const clothgpu::Link& l = c->link[side];
    if (!l.on) return p;  // row_lo == row_hi == 0: no row matches
    p.x = (T*)(l.base + l.off_plane[0 + pos]);
    p.y = (T*)(l.base + l.off_plane[2 + pos]);
    p.px = (T*)(l.base + l.off_plane[4 + prev]);
    p.py = (T*)(l.base + l.off_plane[6 + prev]);
    p.fl = (uint8_t*)(l.base + l.off_plane[8 + pos]);
    p.cloth_stride = l.cloth_stride;
    p.held_row0 = l.held_row0;
    p.row_lo = l.row_lo;
    p.row_hi = l.row_hi;
    return p;
}

bool attached(const clothgpu* c) { return c->link[0].on || c->link[1].on; }

// The in-kernel handshake of one frame-kernel launch (cloth_types.cuh, StripSync).  `columns` = how many pieces of
// work share one row of the cloth (column windows / bands / tile columns, times the ensemble): a neighbour's ghost
// copy of my rows [row_lo, row_hi) is complete when rows x columns units have been reported.  Counts as one strip
// operation (the epoch advances).
StripSync make_sync(clothgpu* c, long long columns) {
    StripSync s{};
    if (!attached(c)) return s;
    s.my_slots = c->slots;
    s.ticket = c->ticket;
    s.need = c->epoch;
    for (int side = 0; side < 2; ++side) {
        const clothgpu::Link& l = c->link[side];
        if (!l.on) continue;
        // I am the LOWER neighbour of my upper neighbour: its slot 1; and the UPPER neighbour of my lower one: its slot 0
        s.peer_slot[side] = (unsigned int*)(l.base + l.off_slots) + (side == 0 ? 1 : 0);
        c->ticket_goal[side] += (unsigned long long)(l.row_hi - l.row_lo) * (unsigned long long)columns;
        s.goal[side] = c->ticket_goal[side];
    }
    c->epoch++;
    return s;
}

// Host-driven strip operations (push, reset): operation k may start once both neighbours finished operation k-1
// (cloth_kernels.cuh).  The frame kernels do this themselves, piece by piece (make_sync).
int strip_begin_op(clothgpu* c) {
    if (!attached(c)) return CLOTHGPU_OK;
    strip_wait_kernel<<<1, 1, 0, c->stream>>>(c->slots, c->link[0].on, c->link[1].on, c->epoch);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CLOTHGPU_OK;
}

int strip_end_op(clothgpu* c) {
    if (!attached(c)) return CLOTHGPU_OK;
    c->epoch++;
    // I am the LOWER neighbour of my upper neighbour: its slot 1; and the UPPER neighbour of my lower one: its slot 0
    unsigned int* up = c->link[0].on ? (unsigned int*)(c->link[0].base + c->link[0].off_slots) + 1 : nullptr;
    unsigned int* dn = c->link[1].on ? (unsigned int*)(c->link[1].base + c->link[1].off_slots) + 0 : nullptr;
    strip_signal_kernel<<<1, 1, 0, c->stream>>>(up, dn, c->epoch);
    CU_TRY(cudaGetLastError());
    c->launches++;
    return CLOTHGPU_OK;
}

template <typename T>
int launch_fused(clothgpu* c, const FrameU<T>& u, const FrameU<T>* per_cloth, int K, int do_verlet) {
    // The frame block travels by value in the kernel parameters; per-cloth blocks (ensembles) through
    // device memory.
    int variant = choose_variant(c, K);
    FusedPlan pl = make_plan<T>(c, K, do_verlet, variant);
    if (pl.TH == 0) {
        variant = 4;
        pl = make_plan<T>(c, K, do_verlet, variant);
    }
    const size_t smem = fused_smem_bytes<T>(pl.TH);
    const int nxt = c->cur_pos ^ 1;
    Planes<T> in = planes<T>(c, c->cur_pos, c->cur_prev);
    const int nprev = do_verlet ? (c->cur_prev ^ 1) : c->cur_prev;
    Planes<T> out = planes<T>(c, nxt, nprev);
    // strips: the neighbours run the same operation sequence, so their buffer indices equal mine
    const PeerOut<T> pup = peer_out<T>(c, 0, nxt, nprev), pdn = peer_out<T>(c, 1, nxt, nprev);
    const StripSync ss = make_sync(c, (long long)pl.tiles_x * c->g.ensemble);
    dim3 grid(pl.tiles_x, pl.tiles_y, c->g.ensemble);
    auto go = [&](auto kernel, int threads) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e != cudaSuccess) return e;
        kernel<<<grid, threads, smem, c->stream>>>(in, out, c->g, pl, u, per_cloth, c->ctr, pup, pdn, ss);
        return cudaSuccess;
    };
    // PER_CLOTH: every cloth has its own frame block (buttons differ), so the tear test stays in.  Otherwise the
    // DRAG=false instantiation drops it for frames without a drag (the common case).
    const bool pc = per_cloth != nullptr;
    const bool drag = pc || (u.buttons & CLOTHGPU_MOUSE_DRAGGING);
    auto pick = [&](auto nt, auto blk, auto ctas) -> cudaError_t {
        constexpr int NT = decltype(nt)::value, BLK = decltype(blk)::value, CT = decltype(ctas)::value;
        if (pc) return go(fused_frame_kernel<T, NT, BLK, true, true, CT>, NT);
        return drag ? go(fused_frame_kernel<T, NT, BLK, false, true, CT>, NT)
                    : go(fused_frame_kernel<T, NT, BLK, false, false, CT>, NT);
    };
    using std::integral_constant;
    switch (variant) {
        case 1:
            CU_TRY(pick(integral_constant<int, 512>{}, integral_constant<int, 2>{}, integral_constant<int, 1>{}));
            break;
        case 2:
            CU_TRY(pick(integral_constant<int, 256>{}, integral_constant<int, 1>{}, integral_constant<int, 2>{}));
            break;
        case 5:
            CU_TRY(pick(integral_constant<int, 256>{}, integral_constant<int, 1>{}, integral_constant<int, 3>{}));
            break;
        case 3:
            CU_TRY(pick(integral_constant<int, 640>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{}));
            break;
        case 4:
            CU_TRY(pick(integral_constant<int, 768>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{}));
            break;
        default:
            CU_TRY(pick(integral_constant<int, 512>{}, integral_constant<int, 1>{}, integral_constant<int, 1>{}));
            break;
    }
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->cur_pos = nxt;
    if (do_verlet) c->cur_prev ^= 1;
    c->last_kernel = std::is_same<T, double>::value ? (pc ? "fused_frame_kernel<double,PER_CLOTH>" : drag ? "fused_frame_kernel<double,DRAG>" : "fused_frame_kernel<double>")
                                                    : (pc ? "fused_frame_kernel<float,PER_CLOTH>" : drag ? "fused_frame_kernel<float,DRAG>" : "fused_frame_kernel<float>");
    return CLOTHGPU_OK;
}

// One sweep per frame (the reference's own semantics): the register-streaming kernel, one pass over HBM.
template <typename T>
int launch_stream1(clothgpu* c, const FrameU<T>& u, const FrameU<T>* per_cloth) {
    const Geom& g = c->g;
    Stream1Plan pl{};
    pl.windows = g.nx <= s1_window_cols<T>() ? 1 : (g.nx - s1_window_cols<T>() + s1_window_step<T>() - 1) / s1_window_step<T>() + 1;
    // Row chunks.  Warps are independent and equally loaded, so the launch runs in "waves" of `slots` resident warps
    // (20 per SM with the shared-memory prefetch ring): long chunks amortise the 4 halo rows, but the task count should be just under a
    // whole number of waves.  With many waves the last, partial one no longer matters and 128-row chunks are used.
    const long long slots = (long long)c->sm_count * c->s1_warps_per_sm;
    const long long per_chunk = (long long)pl.windows * g.ensemble;  // tasks added by one more chunk
    int R = c->s1_chunk_rows;
    if (R < 2) {
        R = 128;
        const double waves = (double)(per_chunk * ((g.own_rows + R - 1) / R)) / (double)slots;
        if (waves < 6.0) {
            auto rows_for = [&](long long w) {
                const long long chunks = std::max<long long>(1, w * slots / per_chunk);
                const int r = (int)((g.own_rows + chunks - 1) / chunks);
                return std::max(8, (r + 1) & ~1);
            };
            long long w = std::max<long long>(1, (long long)(waves + 0.5));
            // A launch of ONE wave starts and ends with every warp in the same phase (measured: 16384 x 2048 runs 4.6 %
            // faster in three waves of 64-row chunks than in one wave of 206-row chunks); shorter chunks pay 4 halo rows
            // each, so several waves are only worth it while the chunks stay long enough.
            if (w == 1) {
                if (rows_for(3) >= 64) w = 3;
                else if (rows_for(2) >= 96) w = 2;
            }
            R = rows_for(w);
        }
    }
    pl.chunk_rows = R;
    pl.chunks = (g.own_rows + R - 1) / R;
    pl.tasks = (long long)pl.windows * pl.chunks * g.ensemble;
    const int nxt = c->cur_pos ^ 1, nprev = c->cur_prev ^ 1;
    Planes<T> in = planes<T>(c, c->cur_pos, c->cur_prev);
    Planes<T> out = planes<T>(c, nxt, nprev);
    const PeerOut<T> pup = peer_out<T>(c, 0, nxt, nprev), pdn = peer_out<T>(c, 1, nxt, nprev);
    const StripSync ss = make_sync(c, (long long)pl.windows * g.ensemble);
    const unsigned blocks = (unsigned)((pl.tasks + kS1Threads / 32 - 1) / (kS1Threads / 32));
    const bool pc = per_cloth != nullptr;
    const bool drag = pc || (u.buttons & CLOTHGPU_MOUSE_DRAGGING);
    auto go = [&](auto kernel, size_t smem) -> cudaError_t {
        if (smem > 48 * 1024) {
            cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (e != cudaSuccess) return e;
        }
        kernel<<<blocks, kS1Threads, smem, c->stream>>>(in, out, g, pl, u, per_cloth, c->ctr, pup, pdn, ss);
        return cudaSuccess;
    };
    auto pick = [&](auto ring) -> cudaError_t {
        constexpr int RING = decltype(ring)::value;
        const size_t smem = (size_t)RING * kS1Threads * 8 * 16;  // per lane and stage: 2 rows x 4 planes x 16 bytes
        if (pc) return go(stream1_frame_kernel<T, true, true, RING>, smem);
        return drag ? go(stream1_frame_kernel<T, false, true, RING>, smem) : go(stream1_frame_kernel<T, false, false, RING>, smem);
    };
    switch (c->s1_ring) {
        case 2: CU_TRY(pick(std::integral_constant<int, 2>{})); break;
        case 3: CU_TRY(pick(std::integral_constant<int, 3>{})); break;
        default: CU_TRY(pick(std::integral_constant<int, 0>{})); break;
    }
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->cur_pos = nxt;
    c->cur_prev = nprev;
    c->last_kernel = std::is_same<T, double>::value ? (pc ? "stream1_frame_kernel<double,PER_CLOTH>" : drag ? "stream1_frame_kernel<double,DRAG>" : "stream1_frame_kernel<double>")
                                                    : (pc ? "stream1_frame_kernel<float,PER_CLOTH>" : drag ? "stream1_frame_kernel<float,DRAG>" : "stream1_frame_kernel<float>");
    return CLOTHGPU_OK;
}

// 2..8 sweeps per frame: the wavefront kernel (cloth_wave.cuh).
// Ensembles of cloths of at most 128 columns run on the wavefront kernel as ONE tall virtual cloth: two cloths side by
// side in the 256-column band, the pairs stacked on top of each other (cloth_wave.cuh, WavePlan::pack).
static bool wave_packable(const clothgpu* c, bool per_cloth_frames) {
    const Geom& g = c->g;
    (void)per_cloth_frames;  // (one frame block per cloth is fine: cloth_wave.cuh, kPerClothPack)
    return c->wave_pack && c->desc.strip_rows <= 0 && g.ensemble >= 2 && g.nx <= kWaveTW / 2 && g.own_rows == g.held_rows &&
           (long long)((g.ensemble + 1) / 2) * ((g.own_rows + 1) & ~1) < (1ll << 30);
}

template <typename T>
int launch_wave(clothgpu* c, const FrameU<T>& u, const FrameU<T>* per_cloth) {
    Geom g = c->g;
    WavePlan pl{};
    if (wave_packable(c, per_cloth != nullptr)) {
        pl.pack = 1;
        pl.nx = g.nx, pl.ny = g.own_rows, pl.rows_v = (g.own_rows + 1) & ~1, pl.ensemble = g.ensemble;
        pl.pitch = g.pitch, pl.cloth_stride = g.cloth_stride;
        // the virtual cloth the kernel streams (its pitch / cloth_stride are not used: the entry warps map rows themselves)
        g.nx = kWaveTW;
        g.own_rows = g.held_rows = ((pl.ensemble + 1) / 2) * pl.rows_v;
        g.own_row0 = g.held_row0 = 0;
        g.ensemble = 1;
    }
    pl.K = u.iterations;
    pl.h = 2 * pl.K;
    pl.RB = 4 * pl.K + 4 + 2 * kWaveSlack;
    if (g.nx <= kWaveTW) {
        pl.bands = 1;
        pl.first_w = kWaveTW;
        pl.mid_w = kWaveTW - 2 * pl.h;
    } else {
        pl.first_w = kWaveTW - pl.h;
        pl.mid_w = kWaveTW - 2 * pl.h;
        pl.bands = 1 + (g.nx - pl.first_w + pl.mid_w - 1) / pl.mid_w;
    }
    // Work split: the row units of all band columns laid end to end, cut into equal spans, one per CTA, a whole number
    // of CTAs per SM (one CTA is resident per SM).  A span pays 2h halo rows (trapezoid-wise) and ~2K pipeline steps of
    // ramp per chunk it touches, so spans should be long; several waves only when the spans stay long anyway.
    pl.columns = pl.bands * g.ensemble;
    const long long rows_e = (g.own_rows + 1) & ~1;
    const long long total = (long long)pl.columns * rows_e;
    long long span = c->wave_chunk_rows;
    if (span < 2) {
        long long ctas = c->sm_count;
        // (packed ensembles have no row halo to amortise and a costlier entry side: one wave of long spans measured
        //  best, 1.96 vs 2.05 ms on 2048 cloths of 128^2)
        while (total / ctas > (pl.pack ? 4096 : 640)) ctas += c->sm_count;
        span = (total + ctas - 1) / ctas;
        span = std::max<long long>(2 * pl.h, (span + 1) & ~1ll);
    }
    pl.span = (int)std::min<long long>(span, rows_e * (long long)pl.columns);
    const unsigned n_ctas = (unsigned)((total + pl.span - 1) / pl.span);
    const size_t smem = wave_smem_bytes<T>(pl.RB);
    const int nxt = c->cur_pos ^ 1, nprev = c->cur_prev ^ 1;
    Planes<T> in = planes<T>(c, c->cur_pos, c->cur_prev);
    Planes<T> out = planes<T>(c, nxt, nprev);
    const PeerOut<T> pup = peer_out<T>(c, 0, nxt, nprev), pdn = peer_out<T>(c, 1, nxt, nprev);
    const StripSync ss = make_sync(c, pl.columns);
    dim3 grid(n_ctas);
    const bool pc = per_cloth != nullptr;
    const bool drag = pc || (u.buttons & CLOTHGPU_MOUSE_DRAGGING);
    auto go = [&](auto kernel) -> cudaError_t {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(227 * 1024));
        if (e != cudaSuccess) return e;
        kernel<<<grid, kWaveThreads, smem, c->stream>>>(in, out, g, pl, u, per_cloth, c->ctr, pup, pdn, ss);
        return cudaSuccess;
    };
    if (pc && pl.pack)
        CU_TRY(go(wave_frame_kernel<T, true, true, true>));
    else if (pc)
        CU_TRY(go(wave_frame_kernel<T, true, true>));
    else if (pl.pack)
        CU_TRY(drag ? go(wave_frame_kernel<T, false, true, true>) : go(wave_frame_kernel<T, false, false, true>));
    else if (drag)
        CU_TRY(go(wave_frame_kernel<T, false, true>));
    else
        CU_TRY(go(wave_frame_kernel<T, false, false>));
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->cur_pos = nxt;
    c->cur_prev = nprev;
    c->last_kernel = std::is_same<T, double>::value
                         ? (pc ? (pl.pack ? "wave_frame_kernel<double,PER_CLOTH,PACK>" : "wave_frame_kernel<double,PER_CLOTH>") : pl.pack ? (drag ? "wave_frame_kernel<double,DRAG,PACK>" : "wave_frame_kernel<double,PACK>")
                                                                                 : drag ? "wave_frame_kernel<double,DRAG>" : "wave_frame_kernel<double>")
                         : (pc ? (pl.pack ? "wave_frame_kernel<float,PER_CLOTH,PACK>" : "wave_frame_kernel<float,PER_CLOTH>") : pl.pack ? (drag ? "wave_frame_kernel<float,DRAG,PACK>" : "wave_frame_kernel<float,PACK>")
                                                                                : drag ? "wave_frame_kernel<float,DRAG>" : "wave_frame_kernel<float>");
    return CLOTHGPU_OK;
}

template <typename T>
int launch_streaming(clothgpu* c, const FrameU<T>& u, const FrameU<T>* per_cloth) {
    const Geom& g = c->g;
    Planes<T> P = planes<T>(c, c->cur_pos, c->cur_prev);
    const FrameU<T>* fr = per_cloth;
    const long long np = (long long)g.held_rows * g.pitch;
    dim3 gp((unsigned)((np + kStreamThreads - 1) / kStreamThreads), g.ensemble);
    const long long nh = (long long)g.held_rows * std::max(1, g.nx / 2);
    const long long nv = (long long)((g.held_rows + 1) / 2) * g.nx;
    dim3 gh((unsigned)((nh + kStreamThreads - 1) / kStreamThreads), g.ensemble);
    dim3 gv((unsigned)((nv + kStreamThreads - 1) / kStreamThreads), g.ensemble);
    auto run = [&](auto pcflag) {
        constexpr bool PC = decltype(pcflag)::value;
        verlet_kernel<T, PC><<<gp, kStreamThreads, 0, c->stream>>>(P, g, u, fr, c->ctr);
        c->launches++;
        for (int it = 0; it < u.iterations; ++it) {
            const int last = it + 1 == u.iterations;
            relax_pass_kernel<T, 0, PC><<<gh, kStreamThreads, 0, c->stream>>>(P, g, u, fr, c->ctr, last);
            relax_pass_kernel<T, 1, PC><<<gh, kStreamThreads, 0, c->stream>>>(P, g, u, fr, c->ctr, last);
            relax_pass_kernel<T, 2, PC><<<gv, kStreamThreads, 0, c->stream>>>(P, g, u, fr, c->ctr, last);
            relax_pass_kernel<T, 3, PC><<<gv, kStreamThreads, 0, c->stream>>>(P, g, u, fr, c->ctr, last);
            c->launches += 4;
        }
    };
    if (fr)
        run(std::true_type{});
    else
        run(std::false_type{});
    CU_TRY(cudaGetLastError());
    c->last_kernel = std::is_same<T, double>::value ? "verlet_kernel<double> + relax_pass_kernel<double> x 4K" : "verlet_kernel<float> + relax_pass_kernel<float> x 4K";
    return CLOTHGPU_OK;
}

}  // namespace

// torn_last bookkeeping: frame f accumulates into ctr->torn_last; it is zeroed (stream ordered)
// before the frame's first launch.
namespace {

// A cloth small enough to live in the shared memory of one CTA (cloth_small.cuh): one cloth, not a strip.
bool small_ok(const clothgpu* c) {
    const Geom& g = c->g;
    const long long n = (long long)g.nx * g.own_rows;
    const size_t bytes = c->precision == CLOTHGPU_F32 ? small_smem_bytes<float>(n) : small_smem_bytes<double>(n);
    return g.ensemble == 1 && c->desc.strip_rows <= 0 && n >= 1 && bytes + 1024 <= (size_t)227 * 1024;
}

int validate_frame(const clothgpu_frame& f, int i) {
    const double vals[] = {f.slider[0], f.slider[1], f.slider[2], f.slider[3], f.slider[4], f.drag_force_max,
                           f.mouse_x, f.mouse_y, f.mouse_px, f.mouse_py, f.mouse_force, f.scroll_y,
                           f.win_off_x, f.win_off_y, f.dt, f.tear_length, f.relax_gain};
    for (double v : vals)
        if (!(std::fabs(v) <= kMaxMagnitude))
            return fail(CLOTHGPU_ERR_INVALID, "frame %d holds a non-finite or absurd value (%g)", i, v);
    return CLOTHGPU_OK;
}

// n frames in ONE launch of one resident CTA.
template <typename T>
int launch_small(clothgpu* c, const clothgpu_frame* frames, int n) {
    if (n <= 0) return CLOTHGPU_OK;
    for (int i = 0; i < n; ++i)
        if (int rc = validate_frame(frames[i], i)) return rc;
    if (c->small_frames_cap < (size_t)n) {
        CU_TRY(cudaStreamSynchronize(c->stream));  // an earlier launch may still read the old buffers
        cudaFree(c->small_frames_dev);
        cudaFreeHost(c->small_frames_host);
        c->small_frames_dev = c->small_frames_host = nullptr;
        c->small_frames_cap = 0;
        const size_t cap = std::max<size_t>(64, (size_t)n);
        CU_TRY(cudaMalloc(&c->small_frames_dev, cap * sizeof(FrameU<double>)));
        CU_TRY(cudaMallocHost(&c->small_frames_host, cap * sizeof(FrameU<double>)));
        c->small_frames_cap = cap;
    }
    CU_TRY(cudaEventSynchronize(c->frames_copied));  // the previous upload has left the pinned staging buffer
    FrameU<T>* st = (FrameU<T>*)c->small_frames_host;
    for (int i = 0; i < n; ++i) st[i] = fold_frame<T>(frames[i], c->desc.spacing);
    CU_TRY(cudaMemcpyAsync(c->small_frames_dev, st, sizeof(FrameU<T>) * (size_t)n, cudaMemcpyHostToDevice, c->stream));
    CU_TRY(cudaEventRecord(c->frames_copied, c->stream));
    const long long np = (long long)c->g.nx * c->g.own_rows;
    const size_t smem = small_smem_bytes<T>(np);
    CU_TRY(cudaFuncSetAttribute(small_frames_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(226 * 1024)));
    // frame i of the call uses counter slot (ctr_slot + i) & 1; the kernel reports through the slot of the last one
    DevCounters* last = c->ctr_base + ((c->ctr_slot + n - 1) & 1);
    small_frames_kernel<T><<<1, kSmallThreads, smem, c->stream>>>(planes<T>(c, c->cur_pos, c->cur_prev), c->g,
                                                                  (const FrameU<T>*)c->small_frames_dev, n, last);
    CU_TRY(cudaGetLastError());
    c->launches++;
    c->ctr = last;
    c->ctr_slot = (c->ctr_slot + n) & 1;
    c->frames += (uint64_t)n;
    c->last_sched = CLOTHGPU_SCHED_RESIDENT;
    c->last_kernel = std::is_same<T, double>::value ? "small_frames_kernel<double>" : "small_frames_kernel<float>";
    return CLOTHGPU_OK;
}

template <typename T>
int step_frames(clothgpu* c, const clothgpu_frame* frames, int count, bool per_cloth) {
    // validate first.  Every frame value must be finite and of sane magnitude: with finite inputs the update can
    // neither produce a non-finite coordinate nor leave the range the shared-seed sqrt/div expansion is exact on
    // (cloth_fused.cuh), so the kernels carry no per-stick range test.  (The Go code would propagate NaN instead.)
    for (int i = 0; i < count; ++i) {
        const clothgpu_frame& f = frames[i];
        if (int rc = validate_frame(f, i)) return rc;
        const int K = f.iterations < 1 ? 1 : f.iterations;
        // attached strips refresh their ghost rows after every launch, detached ones once per frame (by the caller)
        const int depth = 2 * (attached(c) ? std::min(K, c->sweeps_per_launch) : K);
        if (c->desc.strip_rows > 0 && depth > c->ghost)
            return fail(CLOTHGPU_ERR_INVALID, "iterations=%d needs %d ghost rows, handle has %d (raise desc.max_iterations)",
                        K, depth, c->ghost);
        if (per_cloth && (K != (frames[0].iterations < 1 ? 1 : frames[0].iterations) || f.tear_length != frames[0].tear_length ||
                          f.relax_gain != frames[0].relax_gain))
            return fail(CLOTHGPU_ERR_INVALID, "per-cloth frames must share `iterations`, `tear_length` and `relax_gain`");
    }
    const FrameU<T>* dev_frames = nullptr;
    FrameU<T> u0 = fold_frame<T>(frames[0], c->desc.spacing);
    if (per_cloth) {
        // wait until the previous upload has left the pinned staging buffer
        CU_TRY(cudaEventSynchronize(c->frames_copied));
        FrameU<T>* st = (FrameU<T>*)c->frames_host;
        for (int e = 0; e < count; ++e) st[e] = fold_frame<T>(frames[e], c->desc.spacing);
        CU_TRY(cudaMemcpyAsync(c->frames_dev, st, sizeof(FrameU<T>) * count, cudaMemcpyHostToDevice, c->stream));
        CU_TRY(cudaEventRecord(c->frames_copied, c->stream));
        dev_frames = (const FrameU<T>*)c->frames_dev;
    }
    // torn_last, live_last, hl_last of this frame's counter slot were zeroed by the previous frame's first kernel
    // (cloth_types.cuh, prepare_next_frame_counters): no memset node between two frame kernels
    c->ctr = c->ctr_base + c->ctr_slot;
    int sched = c->schedule;
    if (sched == CLOTHGPU_SCHED_RESIDENT) {
        if (!per_cloth && small_ok(c)) return launch_small<T>(c, frames, 1);
        sched = CLOTHGPU_SCHED_AUTO;  // too large for one CTA's shared memory (or an ensemble): what AUTO would take
    }
    if (sched == CLOTHGPU_SCHED_AUTO) {
        // One sweep: the row-streaming kernel, unless the cloth is too small to feed enough independent warps.
        // 4..8 sweeps on cloths that fill the 256-column bands and amortise the pipeline ramp: the wavefront kernel
        // (measured on 4096^2 against the tile kernel: 0.617 vs 0.569 ms at 3 sweeps, 0.683 vs 0.704 at 4, 0.734 vs
        // 0.854 at 5, 0.97 vs 1.48 at 8).
        const int step = std::is_same<T, float>::value ? s1_window_step<float>() : s1_window_step<double>();
        const long long warps = (long long)((c->g.nx + step - 1) / step) * ((c->g.own_rows + 7) / 8) * c->g.ensemble;
        if (u0.iterations == 1 && c->k1_stream && warps >= 2ll * c->sm_count)
            sched = CLOTHGPU_SCHED_ROWSTREAM;
        else if (u0.iterations >= 4 && u0.iterations <= kMaxSweepsPerLaunch && c->wave_auto &&
                 ((c->g.nx >= 192 && c->g.own_rows >= 256) ||
                  (wave_packable(c, dev_frames != nullptr) && (long long)c->g.ensemble * c->g.own_rows >= 256ll * c->sm_count)))
            sched = CLOTHGPU_SCHED_WAVE;
        else
            sched = CLOTHGPU_SCHED_FUSED;
    } else if (sched == CLOTHGPU_SCHED_ROWSTREAM && u0.iterations != 1) {
        sched = CLOTHGPU_SCHED_FUSED;
    }
    if (sched == CLOTHGPU_SCHED_WAVE && (u0.iterations < 2 || u0.iterations > kMaxSweepsPerLaunch)) sched = CLOTHGPU_SCHED_FUSED;
    if (attached(c) && sched == CLOTHGPU_SCHED_STREAMING)
        return fail(CLOTHGPU_ERR_STATE, "attached strips exchange their edge rows inside the fused kernel; the streaming schedule is for detached handles");
    if (c->desc.strip_rows > 0 && !attached(c) && sched != CLOTHGPU_SCHED_STREAMING && u0.iterations > c->sweeps_per_launch)
        return fail(CLOTHGPU_ERR_STATE, "a detached strip cannot split %d sweeps over several launches (ghost rows would go stale between launches); attach the neighbours or use <= %d", u0.iterations, c->sweeps_per_launch);
    int rc;
    if (sched == CLOTHGPU_SCHED_STREAMING) {
        rc = launch_streaming<T>(c, u0, dev_frames);
    } else if (sched == CLOTHGPU_SCHED_ROWSTREAM) {
        rc = launch_stream1<T>(c, u0, dev_frames);
    } else if (sched == CLOTHGPU_SCHED_WAVE) {
        rc = launch_wave<T>(c, u0, dev_frames);
    } else {
        int left = u0.iterations, first = 1;
        rc = CLOTHGPU_OK;
        while (left > 0 && rc == CLOTHGPU_OK) {
            const int k = std::min(left, c->sweeps_per_launch);
            if (!first) CU_TRY(cudaMemsetAsync(&c->ctr->live_last, 0, sizeof(unsigned long long), c->stream));  // the last launch's count stands
            rc = launch_fused<T>(c, u0, dev_frames, k, first);
            first = 0;
            left -= k;
        }
    }
    if (rc == CLOTHGPU_OK) {
        c->frames++;
        c->last_sched = sched;
        c->ctr_slot ^= 1;
    }
    return rc;
}

int step_dispatch(clothgpu* c, const clothgpu_frame* frames, int count, bool per_cloth) {
    return c->precision == CLOTHGPU_F32 ? step_frames<float>(c, frames, count, per_cloth)
                                        : step_frames<double>(c, frames, count, per_cloth);
}

template <typename T>
int init_state(clothgpu* c, int pos_x, int pos_y) {
    // Attached strips: Reset is a strip operation like a frame.  The init kernel rewrites my ghost rows too, so the
    // neighbours' stores of the previous operation into them must have landed first (strip_begin_op), and the
    // neighbours may only store into my ghost rows again once the init is done (strip_end_op).  Every strip resets at
    // the same point of the common call sequence, so all of them are back on buffer 0 afterwards.
    if (int rc = strip_begin_op(c)) return rc;
    c->cur_pos = c->cur_prev = 0;
    Planes<T> P = planes<T>(c, 0, 0);
    init_grid_kernel<T><<<c->sm_count * 8, 256, 0, c->stream>>>(P, c->g, c->ny, c->desc.spacing, pos_x,
                                                                pos_y, c->desc.pin_rule);
    CU_TRY(cudaGetLastError());
    c->launches++;
    CU_TRY(cudaMemsetAsync(c->ctr_base, 0, 2 * sizeof(DevCounters), c->stream));
    c->ctr = c->ctr_base;
    c->ctr_slot = 0;
    c->frames = 0;
    c->census_valid = false;
    return strip_end_op(c);
}

void free_all(clothgpu* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    for (auto& l : c->link)
        if (l.on && l.ipc) cudaIpcCloseMemHandle(l.base);
    cudaFree(c->slab);
    cudaFree(c->ctr_base);
    cudaFree(c->census);
    cudaFreeHost(c->ctr_host);
    cudaFree(c->frames_dev);
    cudaFreeHost(c->frames_host);
    cudaFree(c->small_frames_dev);
    cudaFreeHost(c->small_frames_host);
    cudaFree(c->seg_plan);
    cudaFree(c->seg_starts);
    cudaFree(c->seg_out);
    cudaFree(c->seg_total);
    if (c->frames_copied) cudaEventDestroy(c->frames_copied);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

int check_handle(clothgpu* c) {
    if (!c) return fail(CLOTHGPU_ERR_INVALID, "null handle");
    cudaError_t e = cudaSetDevice(c->device);
    if (e != cudaSuccess) return fail(CLOTHGPU_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
    return CLOTHGPU_OK;
}

}  // namespace

extern "C" {

int clothgpu_frame_default(clothgpu_frame* f) {
    if (!f) return fail(CLOTHGPU_ERR_INVALID, "null frame");
    memset(f, 0, sizeof *f);
    // gui/hud.go:86-92 (row i of the slider table feeds enum value i, gui/hud.go:94-97)
    f->slider[CLOTHGPU_SLIDER_DRAG_FORCE] = 2.0f;
    f->slider[CLOTHGPU_SLIDER_GRAVITY] = 250.0f;
    f->slider[CLOTHGPU_SLIDER_STIFFNESS] = 30.0f;
    f->slider[CLOTHGPU_SLIDER_FRICTION] = 0.98f;
    f->slider[CLOTHGPU_SLIDER_TEAR_DISTANCE] = 15.0f;
    f->drag_force_max = 15.0f;
    f->scroll_y = 50.0f;  // consts.DefaultFocusArea, main.go:84
    f->win_w = 1280;      // consts.WindowSizeX/Y
    f->win_h = 820;
    f->dt = 0.022;        // consts.Delta
    f->iterations = 1;
    f->tear_length = 150.0;
    f->relax_gain = 0.35;
    return CLOTHGPU_OK;
}

int clothgpu_desc_default(clothgpu_desc* d) {
    if (!d) return fail(CLOTHGPU_ERR_INVALID, "null desc");
    memset(d, 0, sizeof *d);
    d->width = 1024;   // main.go:150-158: clothW = Dp(DefaultWindowWidth)
    d->height = 211;   // Dp(640 * 0.33) = 211
    d->spacing = 6;
    d->pos_x = 128;    // (1280 - 1024) / 2
    d->pos_y = 164;    // int(820 * 0.2)
    d->pin_rule = CLOTHGPU_PIN_REFERENCE;
    d->precision = CLOTHGPU_F64;
    d->ensemble = 1;
    d->max_iterations = 1;
    return CLOTHGPU_OK;
}

int clothgpu_create(const clothgpu_desc* desc, clothgpu** out) {
    if (!desc || !out) return fail(CLOTHGPU_ERR_INVALID, "null argument");
    *out = nullptr;
    if (desc->spacing <= 0 || desc->width < 0 || desc->height < 0)
        return fail(CLOTHGPU_ERR_INVALID, "width/height must be >= 0 and spacing > 0");
    if (desc->precision != CLOTHGPU_F64 && desc->precision != CLOTHGPU_F32)
        return fail(CLOTHGPU_ERR_INVALID, "unknown precision %d", desc->precision);
    if (desc->pin_rule < CLOTHGPU_PIN_REFERENCE || desc->pin_rule > CLOTHGPU_PIN_NONE)
        return fail(CLOTHGPU_ERR_INVALID, "unknown pin rule %d", desc->pin_rule);
    const int clothX = desc->width / desc->spacing;   // physics/cloth.go:43
    const int clothY = desc->height / desc->spacing;  // physics/cloth.go:44
    if (desc->pin_rule == CLOTHGPU_PIN_REFERENCE && clothX / 10 == 0)
        return fail(CLOTHGPU_ERR_INVALID,
                    "width/spacing = %d < 10: the reference pin rule divides by zero (physics/cloth.go:72)", clothX);
    const int ensemble = desc->ensemble < 1 ? 1 : desc->ensemble;
    const int nx = clothX + 1, ny = clothY + 1;

    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CLOTHGPU_ERR_NODEVICE, "no CUDA device (%s); clothgpu has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (desc->device < 0 || desc->device >= ndev)
        return fail(CLOTHGPU_ERR_INVALID, "device %d out of range (have %d)", desc->device, ndev);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, desc->device));
    if (prop.major != 10)
        return fail(CLOTHGPU_ERR_NODEVICE, "device %d is sm_%d%d; this library carries sm_100a code only",
                    desc->device, prop.major, prop.minor);
    CU_TRY(cudaSetDevice(desc->device));

    clothgpu* c = new (std::nothrow) clothgpu();
    if (!c) return fail(CLOTHGPU_ERR_NOMEM, "out of host memory");
    c->desc = *desc;
    c->desc.ensemble = ensemble;
    c->ny = ny;
    c->precision = desc->precision;
    c->elem = desc->precision == CLOTHGPU_F32 ? 4 : 8;
    c->device = desc->device;
    c->sm_count = prop.multiProcessorCount;
    Geom& g = c->g;
    g.nx = nx;
    g.pitch = (nx + 15) / 16 * 16;
    g.ensemble = ensemble;
    if (desc->strip_rows > 0) {
        if (desc->strip_row0 < 0 || desc->strip_row0 + desc->strip_rows > ny || (desc->strip_row0 & 1)) {
            delete c;
            return fail(CLOTHGPU_ERR_INVALID, "strip rows [%d,%d) must lie in [0,%d) and start on an even row",
                        desc->strip_row0, desc->strip_row0 + desc->strip_rows, ny);
        }
        const int kmax = desc->max_iterations < 1 ? 1 : desc->max_iterations;
        c->ghost = 2 * kmax;
        const int lo = std::max(0, desc->strip_row0 - c->ghost);
        const int hi = std::min(ny, desc->strip_row0 + desc->strip_rows + c->ghost);
        g.held_row0 = lo;
        g.held_rows = hi - lo;
        g.own_row0 = desc->strip_row0;
        g.own_rows = desc->strip_rows;
    } else {
        g.held_row0 = g.own_row0 = 0;
        g.held_rows = g.own_rows = ny;
    }
    // one spare row keeps the 16-bit flag loads / pair stores of the last row's padding in bounds
    g.cloth_stride = (long long)(g.held_rows + 1) * g.pitch;
    c->plane_elems = (size_t)g.cloth_stride * ensemble;
    if (const char* s = getenv("CLOTHGPU_TILE_ROWS")) c->forced_tile_rows = atoi(s);
    if (const char* s = getenv("CLOTHGPU_FUSED_VARIANT")) c->fused_variant = atoi(s);
    if (const char* s = getenv("CLOTHGPU_K1_STREAM")) c->k1_stream = atoi(s);
    if (const char* s = getenv("CLOTHGPU_S1_ROWS")) c->s1_chunk_rows = atoi(s) & ~1;
    if (const char* s = getenv("CLOTHGPU_WAVE_ROWS")) c->wave_chunk_rows = atoi(s) & ~1;
    if (const char* s = getenv("CLOTHGPU_WAVE_AUTO")) c->wave_auto = atoi(s);
    if (const char* s = getenv("CLOTHGPU_WAVE_PACK")) c->wave_pack = atoi(s);
    if (const char* s = getenv("CLOTHGPU_SMALL_AUTO")) c->small_auto = atoi(s);
    if (const char* s = getenv("CLOTHGPU_S1_RING")) {
        c->s1_ring = atoi(s);
        c->s1_warps_per_sm = c->s1_ring == 2 ? 4 * S1_RING_CTAS : 4 * S1_MIN_CTAS;  // (3 stages: 48 KB per CTA, 4 CTAs fit)
    }
    if (const char* s = getenv("CLOTHGPU_DIRECT_EXPORT")) c->direct_export = atoi(s) != 0;
    if (const char* s = getenv("CLOTHGPU_S1_WARPS")) c->s1_warps_per_sm = std::max(1, atoi(s));
    if (const char* s = getenv("CLOTHGPU_SWEEPS_PER_LAUNCH")) {
        const int v = atoi(s);
        if (v >= 1 && v <= kMaxSweepsPerLaunch) c->sweeps_per_launch = v;
    }

    int rc = CLOTHGPU_OK;
    auto alloc = [&](void** p, size_t bytes) {
        if (rc != CLOTHGPU_OK) return;
        cudaError_t ee = cudaMalloc(p, bytes);
        if (ee != cudaSuccess) {
            rc = fail(ee == cudaErrorMemoryAllocation ? CLOTHGPU_ERR_NOMEM : CLOTHGPU_ERR_CUDA,
                      "cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(ee));
            cudaGetLastError();
            return;
        }
        ee = cudaMemset(*p, 0, bytes);
        if (ee != cudaSuccess) rc = fail(CLOTHGPU_ERR_CUDA, "cudaMemset: %s", cudaGetErrorString(ee));
    };
    {
        size_t off = 0;
        auto take = [&](size_t bytes) {
            const size_t at = off;
            off += (bytes + 255) / 256 * 256;
            return at;
        };
        for (int k = 0; k < 8; ++k) c->off_plane[k] = take(c->plane_elems * c->elem);
        for (int k = 8; k < 10; ++k) c->off_plane[k] = take(c->plane_elems);
        c->off_slots = take(256);
        c->slab_bytes = off;
        alloc(&c->slab, c->slab_bytes);
        if (rc == CLOTHGPU_OK) {
            char* b = (char*)c->slab;
            for (int i = 0; i < 2; ++i) {
                c->X[i] = b + c->off_plane[0 + i];
                c->Y[i] = b + c->off_plane[2 + i];
                c->PX[i] = b + c->off_plane[4 + i];
                c->PY[i] = b + c->off_plane[6 + i];
                c->FL[i] = (uint8_t*)(b + c->off_plane[8 + i]);
            }
            c->slots = (unsigned int*)(b + c->off_slots);
            c->ticket = (unsigned long long*)(b + c->off_slots + 64);
        }
    }
    alloc((void**)&c->ctr_base, 2 * sizeof(DevCounters));
    c->ctr = c->ctr_base;
    alloc((void**)&c->census, 5 * sizeof(unsigned long long));
    alloc((void**)&c->seg_total, 2 * sizeof(unsigned long long));
    if (rc == CLOTHGPU_OK) {
        cudaError_t ee = cudaMallocHost((void**)&c->ctr_host, 2 * sizeof(DevCounters) + 6 * sizeof(unsigned long long));  // + the export's total
        if (ee != cudaSuccess) rc = fail(CLOTHGPU_ERR_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(ee));
    }
    const size_t fbytes = sizeof(FrameU<double>) * (size_t)ensemble;
    alloc(&c->frames_dev, fbytes);
    if (rc == CLOTHGPU_OK) {
        cudaError_t ee = cudaMallocHost(&c->frames_host, fbytes);
        if (ee != cudaSuccess) rc = fail(CLOTHGPU_ERR_NOMEM, "cudaMallocHost: %s", cudaGetErrorString(ee));
    }
    if (rc == CLOTHGPU_OK) {
        cudaError_t ee = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
        if (ee == cudaSuccess) ee = cudaEventCreate(&c->ev0);
        if (ee == cudaSuccess) ee = cudaEventCreate(&c->ev1);
        if (ee == cudaSuccess) ee = cudaEventCreateWithFlags(&c->frames_copied, cudaEventDisableTiming);
        if (ee == cudaSuccess) ee = cudaEventRecord(c->frames_copied, c->stream);
        if (ee != cudaSuccess) rc = fail(CLOTHGPU_ERR_CUDA, "stream/event setup: %s", cudaGetErrorString(ee));
    }
    if (rc == CLOTHGPU_OK)
        rc = c->precision == CLOTHGPU_F32 ? init_state<float>(c, desc->pos_x, desc->pos_y)
                                          : init_state<double>(c, desc->pos_x, desc->pos_y);
    if (rc == CLOTHGPU_OK) {
        cudaError_t ee = cudaStreamSynchronize(c->stream);
        if (ee != cudaSuccess) rc = fail(CLOTHGPU_ERR_CUDA, "init kernel: %s", cudaGetErrorString(ee));
    }
    if (rc != CLOTHGPU_OK) {
        std::string keep = g_last_error;
        free_all(c);
        g_last_error = keep;
        return rc;
    }
    *out = c;
    return CLOTHGPU_OK;
}

int clothgpu_reset(clothgpu* c, int32_t pos_x, int32_t pos_y) {
    if (int rc = check_handle(c)) return rc;
    c->desc.pos_x = pos_x;
    c->desc.pos_y = pos_y;
    return c->precision == CLOTHGPU_F32 ? init_state<float>(c, pos_x, pos_y) : init_state<double>(c, pos_x, pos_y);
}

void clothgpu_destroy(clothgpu* c) {
    if (!c) return;
    cudaSetDevice(c->device);
    if (c->stream) cudaStreamSynchronize(c->stream);
    free_all(c);
}

int clothgpu_step(clothgpu* c, const clothgpu_frame* frame) {
    if (int rc = check_handle(c)) return rc;
    if (!frame) return fail(CLOTHGPU_ERR_INVALID, "null frame");
    CU_TRY(cudaEventRecord(c->ev0, c->stream));
    const int rc = step_dispatch(c, frame, 1, false);
    CU_TRY(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    return rc;
}

int clothgpu_step_n(clothgpu* c, const clothgpu_frame* frames, int32_t n) {
    if (int rc = check_handle(c)) return rc;
    if (!frames || n < 0) return fail(CLOTHGPU_ERR_INVALID, "bad frames/n");
    CU_TRY(cudaEventRecord(c->ev0, c->stream));
    int rc = CLOTHGPU_OK;
    // GUI-sized cloths: all n frames in ONE launch of a resident CTA (cloth_small.cuh) instead of a host loop of launches
    if (small_ok(c) && ((c->schedule == CLOTHGPU_SCHED_AUTO && n >= 2 && c->small_auto) || c->schedule == CLOTHGPU_SCHED_RESIDENT))
        rc = c->precision == CLOTHGPU_F32 ? launch_small<float>(c, frames, n) : launch_small<double>(c, frames, n);
    else
        for (int i = 0; i < n && rc == CLOTHGPU_OK; ++i) rc = step_dispatch(c, frames + i, 1, false);
    CU_TRY(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    return rc;
}

int clothgpu_step_each(clothgpu* c, const clothgpu_frame* frames, int32_t count) {
    if (int rc = check_handle(c)) return rc;
    if (!frames || count != c->g.ensemble)
        return fail(CLOTHGPU_ERR_INVALID, "step_each needs exactly ensemble=%d frames", c ? c->g.ensemble : 0);
    CU_TRY(cudaEventRecord(c->ev0, c->stream));
    const int rc = step_dispatch(c, frames, count, true);
    CU_TRY(cudaEventRecord(c->ev1, c->stream));
    c->timed = true;
    return rc;
}

int clothgpu_sync(clothgpu* c) {
    if (int rc = check_handle(c)) return rc;
    CU_TRY(cudaStreamSynchronize(c->stream));
    return CLOTHGPU_OK;
}

int clothgpu_get_info(clothgpu* c, clothgpu_info* info) {
    if (!c || !info) return fail(CLOTHGPU_ERR_INVALID, "null argument");
    memset(info, 0, sizeof *info);
    info->nx = c->g.nx;
    info->ny = c->ny;
    info->row0 = c->g.own_row0;
    info->rows = c->g.own_rows;
    info->ghost = c->ghost;
    info->pitch = c->g.pitch;
    info->ensemble = c->g.ensemble;
    info->precision = c->precision;
    info->particles = (int64_t)c->g.nx * c->g.own_rows;
    info->sticks = 2ll * c->g.nx * c->ny - c->g.nx - c->ny;
    info->device = c->device;
    info->sm_count = c->sm_count;
    return CLOTHGPU_OK;
}

int clothgpu_set_schedule(clothgpu* c, int32_t schedule) {
    if (!c) return fail(CLOTHGPU_ERR_INVALID, "null handle");
    if (schedule < CLOTHGPU_SCHED_AUTO || schedule > CLOTHGPU_SCHED_RESIDENT)
        return fail(CLOTHGPU_ERR_INVALID, "unknown schedule %d", schedule);
    c->schedule = schedule;
    return CLOTHGPU_OK;
}

int clothgpu_read_state(clothgpu* c, int32_t cloth, double* x, double* y, double* px, double* py,
                        uint8_t* flags) {
    if (int rc = check_handle(c)) return rc;
    if (cloth < 0 || cloth >= c->g.ensemble) return fail(CLOTHGPU_ERR_INVALID, "cloth %d out of range", cloth);
    const Geom& g = c->g;
    const size_t off = (size_t)cloth * g.cloth_stride + (size_t)(g.own_row0 - g.held_row0) * g.pitch;
    const size_t n = (size_t)g.nx * g.own_rows;
    CU_TRY(cudaStreamSynchronize(c->stream));
    void* src[4] = {c->X[c->cur_pos], c->Y[c->cur_pos], c->PX[c->cur_prev], c->PY[c->cur_prev]};
    double* dst[4] = {x, y, px, py};
    std::vector<float> tmp;
    for (int k = 0; k < 4; ++k) {
        if (!dst[k]) continue;
        if (c->precision == CLOTHGPU_F64) {
            CU_TRY(cudaMemcpy2D(dst[k], (size_t)g.nx * 8, (const double*)src[k] + off, (size_t)g.pitch * 8,
                                (size_t)g.nx * 8, g.own_rows, cudaMemcpyDeviceToHost));
        } else {
            tmp.resize(n);
            CU_TRY(cudaMemcpy2D(tmp.data(), (size_t)g.nx * 4, (const float*)src[k] + off, (size_t)g.pitch * 4,
                                (size_t)g.nx * 4, g.own_rows, cudaMemcpyDeviceToHost));
            for (size_t i = 0; i < n; ++i) dst[k][i] = (double)tmp[i];
        }
    }
    if (flags)
        CU_TRY(cudaMemcpy2D(flags, g.nx, c->FL[c->cur_pos] + off, g.pitch, g.nx, g.own_rows, cudaMemcpyDeviceToHost));
    return CLOTHGPU_OK;
}

int clothgpu_write_state(clothgpu* c, int32_t cloth, const double* x, const double* y, const double* px,
                         const double* py, const uint8_t* flags) {
    if (int rc = check_handle(c)) return rc;
    if (cloth < 0 || cloth >= c->g.ensemble) return fail(CLOTHGPU_ERR_INVALID, "cloth %d out of range", cloth);
    const Geom& g = c->g;
    const size_t off = (size_t)cloth * g.cloth_stride + (size_t)(g.own_row0 - g.held_row0) * g.pitch;
    const size_t n = (size_t)g.nx * g.own_rows;
    CU_TRY(cudaStreamSynchronize(c->stream));
    void* dstp[4] = {c->X[c->cur_pos], c->Y[c->cur_pos], c->PX[c->cur_prev], c->PY[c->cur_prev]};
    const double* src[4] = {x, y, px, py};
    // coordinates must be finite (see step_frames); -0.0 is stored as +0.0 — the update never produces -0.0 from
    // other inputs, and the kernels rely on p + 0 == p for every stored value
    for (int k = 0; k < 4; ++k)
        if (src[k])
            for (size_t i = 0; i < n; ++i)
                if (!(std::fabs(src[k][i]) <= kMaxMagnitude))
                    return fail(CLOTHGPU_ERR_INVALID, "write_state: plane %d element %zu is not finite / too large (%g)", k, i, src[k][i]);
    std::vector<float> tmp;
    std::vector<double> tmpd;
    for (int k = 0; k < 4; ++k) {
        if (!src[k]) continue;
        if (c->precision == CLOTHGPU_F64) {
            tmpd.resize(n);
            for (size_t i = 0; i < n; ++i) tmpd[i] = src[k][i] + 0.0;  // -0.0 + 0.0 == +0.0
            CU_TRY(cudaMemcpy2D((double*)dstp[k] + off, (size_t)g.pitch * 8, tmpd.data(), (size_t)g.nx * 8,
                                (size_t)g.nx * 8, g.own_rows, cudaMemcpyHostToDevice));
        } else {
            tmp.resize(n);
            for (size_t i = 0; i < n; ++i) tmp[i] = (float)src[k][i] + 0.0f;
            CU_TRY(cudaMemcpy2D((float*)dstp[k] + off, (size_t)g.pitch * 4, tmp.data(), (size_t)g.nx * 4,
                                (size_t)g.nx * 4, g.own_rows, cudaMemcpyHostToDevice));
        }
    }
    if (flags) {
        // keep the invariants Init establishes: no stick past the last column / last row
        // only the five CLOTHGPU_FLAG_* bits are state (the packed-ensemble schedule parks alive bits in the spare ones
        // while a frame is in flight; a stray bit there must not turn into a stick)
        std::vector<uint8_t> fl(flags, flags + n);
        for (size_t i = 0; i < n; ++i) fl[i] &= 0x1f;
        c->census_valid = false;
        for (int r = 0; r < g.own_rows; ++r) {
            fl[(size_t)r * g.nx + g.nx - 1] &= (uint8_t)~CLOTHGPU_FLAG_ALIVE_H;
            if (g.own_row0 + r == c->ny - 1)
                for (int q = 0; q < g.nx; ++q) fl[(size_t)r * g.nx + q] &= (uint8_t)~CLOTHGPU_FLAG_ALIVE_V;
        }
        CU_TRY(cudaMemcpy2D(c->FL[c->cur_pos] + off, g.pitch, fl.data(), g.nx, g.nx, g.own_rows,
                            cudaMemcpyHostToDevice));
    }
    return CLOTHGPU_OK;
}

namespace {
struct StateFileHeader {
    char magic[8];
    uint32_t header_bytes, version;
    int32_t nx, ny, row0, rows;
    int32_t ensemble, spacing, pos_x, pos_y;
    uint64_t frames, reserved;
};
static_assert(sizeof(StateFileHeader) == 64, "state file header is 64 bytes");
constexpr char kStateMagic[8] = {'C', 'L', 'S', 'T', 'A', 'T', 'E', '1'};

uint64_t fnv1a64(uint64_t h, const void* data, size_t n) {
    const unsigned char* p = (const unsigned char*)data;
    for (size_t i = 0; i < n; ++i) h = (h ^ p[i]) * 0x100000001b3ull;
    return h;
}
constexpr uint64_t kFnvBasis = 0xcbf29ce484222325ull;
}  // namespace

int clothgpu_save_state(clothgpu* c, const char* path) {
    if (int rc = check_handle(c)) return rc;
    if (!path) return fail(CLOTHGPU_ERR_INVALID, "null path");
    const Geom& g = c->g;
    const size_t n = (size_t)g.nx * g.own_rows, nfl = (n + 7) / 8 * 8;
    StateFileHeader h{};
    memcpy(h.magic, kStateMagic, 8);
    h.header_bytes = 64, h.version = 1;
    h.nx = g.nx, h.ny = c->ny, h.row0 = g.own_row0, h.rows = g.own_rows;
    h.ensemble = g.ensemble, h.spacing = c->desc.spacing, h.pos_x = c->desc.pos_x, h.pos_y = c->desc.pos_y;
    h.frames = c->frames;
    FILE* f = fopen(path, "wb");
    if (!f) return fail(CLOTHGPU_ERR_INVALID, "cannot open %s for writing", path);
    uint64_t sum = fnv1a64(kFnvBasis, &h, sizeof h);
    bool ok = fwrite(&h, sizeof h, 1, f) == 1;
    std::vector<double> plane[4];
    for (auto& p : plane) p.resize(n);
    std::vector<uint8_t> fl(nfl, 0);
    int rc = CLOTHGPU_OK;
    for (int e = 0; e < g.ensemble && ok && rc == CLOTHGPU_OK; ++e) {
        rc = clothgpu_read_state(c, e, plane[0].data(), plane[1].data(), plane[2].data(), plane[3].data(), fl.data());
        if (rc != CLOTHGPU_OK) break;
        for (auto& p : plane) {
            sum = fnv1a64(sum, p.data(), n * sizeof(double));
            ok = ok && (n == 0 || fwrite(p.data(), sizeof(double), n, f) == n);
        }
        sum = fnv1a64(sum, fl.data(), nfl);
        ok = ok && (nfl == 0 || fwrite(fl.data(), 1, nfl, f) == nfl);
    }
    ok = ok && fwrite(&sum, sizeof sum, 1, f) == 1;
    ok = (fclose(f) == 0) && ok;
    if (rc != CLOTHGPU_OK) return rc;
    if (!ok) return fail(CLOTHGPU_ERR_INVALID, "short write to %s", path);
    return CLOTHGPU_OK;
}

int clothgpu_load_state(clothgpu* c, const char* path) {
    if (int rc = check_handle(c)) return rc;
    if (!path) return fail(CLOTHGPU_ERR_INVALID, "null path");
    const Geom& g = c->g;
    FILE* f = fopen(path, "rb");
    if (!f) return fail(CLOTHGPU_ERR_INVALID, "cannot open %s", path);
    StateFileHeader h{};
    if (fread(&h, sizeof h, 1, f) != 1 || memcmp(h.magic, kStateMagic, 8) != 0 || h.header_bytes != 64 || h.version != 1) {
        fclose(f);
        return fail(CLOTHGPU_ERR_INVALID, "%s is not a clothgpu state file (CLSTATE1, version 1)", path);
    }
    if (h.nx != g.nx || h.ny != c->ny || h.row0 != g.own_row0 || h.rows != g.own_rows || h.ensemble != g.ensemble) {
        fclose(f);
        return fail(CLOTHGPU_ERR_INVALID, "%s holds %d cloth(s) of %dx%d, rows [%d,%d); the handle holds %d of %dx%d, rows [%d,%d)", path,
                    h.ensemble, h.nx, h.ny, h.row0, h.row0 + h.rows, g.ensemble, g.nx, c->ny, g.own_row0, g.own_row0 + g.own_rows);
    }
    const size_t n = (size_t)g.nx * g.own_rows, nfl = (n + 7) / 8 * 8;
    // first pass: read everything and verify the checksum, so a damaged file leaves the handle untouched
    std::vector<std::vector<double>> planes((size_t)g.ensemble * 4);
    std::vector<std::vector<uint8_t>> flags((size_t)g.ensemble);
    uint64_t sum = fnv1a64(kFnvBasis, &h, sizeof h), stored = 0;
    bool ok = true;
    for (int e = 0; e < g.ensemble && ok; ++e) {
        for (int k = 0; k < 4 && ok; ++k) {
            auto& p = planes[(size_t)e * 4 + k];
            p.resize(n);
            ok = n == 0 || fread(p.data(), sizeof(double), n, f) == n;
            sum = fnv1a64(sum, p.data(), n * sizeof(double));
        }
        flags[e].resize(nfl);
        ok = ok && (nfl == 0 || fread(flags[e].data(), 1, nfl, f) == nfl);
        sum = fnv1a64(sum, flags[e].data(), nfl);
    }
    ok = ok && fread(&stored, sizeof stored, 1, f) == 1;
    fclose(f);
    if (!ok) return fail(CLOTHGPU_ERR_INVALID, "%s is truncated", path);
    if (stored != sum) return fail(CLOTHGPU_ERR_INVALID, "%s: checksum mismatch (file damaged)", path);
    for (int e = 0; e < g.ensemble; ++e) {
        const auto* p = &planes[(size_t)e * 4];
        if (int rc = clothgpu_write_state(c, e, p[0].data(), p[1].data(), p[2].data(), p[3].data(), flags[e].data())) return rc;
    }
    c->frames = h.frames;
    c->census_valid = false;
    return CLOTHGPU_OK;
}

int clothgpu_read_masks(clothgpu* c, int32_t cloth, uint32_t* pin, uint32_t* active, uint32_t* highlighted,
                        uint32_t* alive_h, uint32_t* alive_v) {
    if (int rc = check_handle(c)) return rc;
    if (cloth < 0 || cloth >= c->g.ensemble) return fail(CLOTHGPU_ERR_INVALID, "cloth %d out of range", cloth);
    const Geom& g = c->g;
    const long long n = (long long)g.nx * g.own_rows;
    const long long words = (n + 31) / 32;
    uint32_t* dev = nullptr;
    CU_TRY(cudaMalloc(&dev, (size_t)words * 5 * sizeof(uint32_t)));
    const unsigned blocks = (unsigned)((words * 32 + kStreamThreads - 1) / kStreamThreads);
    if (c->precision == CLOTHGPU_F32)
        pack_masks_kernel<float><<<blocks, kStreamThreads, 0, c->stream>>>(
            planes<float>(c, c->cur_pos, c->cur_prev), g, cloth, dev, words);
    else
        pack_masks_kernel<double><<<blocks, kStreamThreads, 0, c->stream>>>(
            planes<double>(c, c->cur_pos, c->cur_prev), g, cloth, dev, words);
    c->launches++;
    cudaError_t e = cudaStreamSynchronize(c->stream);
    uint32_t* dst[5] = {pin, active, highlighted, alive_h, alive_v};
    for (int b = 0; b < 5 && e == cudaSuccess; ++b)
        if (dst[b]) e = cudaMemcpy(dst[b], dev + (size_t)b * words, (size_t)words * 4, cudaMemcpyDeviceToHost);
    cudaFree(dev);
    if (e != cudaSuccess) return fail(CLOTHGPU_ERR_CUDA, "read_masks: %s", cudaGetErrorString(e));
    return CLOTHGPU_OK;
}

namespace {
// Where an export goes when not into the handle's own device buffer: device-visible addresses of the caller's pinned host
// memory (null = the handle's buffer), the list's capacity there, and where its length is to be written.
struct ExportTarget {
    float4* main = nullptr;
    uint8_t* hl = nullptr;
    uint2* idx = nullptr;
    unsigned long long* total = nullptr;
    size_t cap = 0;
};
// Live-stick export on the handle's stream: a flags-only plan launch + the export launch (cloth_kernels.cuh; ONE launch for
// cloths of a few tiles) into c->seg_out — or, with `direct`, straight into the caller's pinned host memory.
// quads = false: float4 endpoints at [0, n), optional uint2 indices behind them; quads = true: 2 float4 of
// outline per stick.  Highlight bytes behind the 32 B/stick area in both cases.
int compact_enqueue(clothgpu* c, int32_t cloth, bool with_idx, bool quads, float line_width, size_t* max_sticks_out,
                    const ExportTarget* direct = nullptr) {
    if (cloth < 0 || cloth >= c->g.ensemble) return fail(CLOTHGPU_ERR_INVALID, "cloth %d out of range", cloth);
    const Geom& g = c->g;
    const long long n = (long long)g.nx * g.own_rows;
    const long long padded = (long long)g.pitch * g.own_rows;
    const unsigned blocks = (unsigned)std::max<long long>(1, (padded + kSegBlock - 1) / kSegBlock);  // tiles
    if (c->seg_plan_cap < blocks) {
        cudaFree(c->seg_plan);
        cudaFree(c->seg_starts);
        c->seg_plan = nullptr, c->seg_starts = nullptr;
        c->seg_plan_cap = 0;
        CU_TRY(cudaMalloc((void**)&c->seg_plan, (size_t)blocks * sizeof(SegPlan)));
        CU_TRY(cudaMalloc((void**)&c->seg_starts, ((size_t)blocks + 15) / 16 * 16 * sizeof(unsigned int)));
        c->seg_plan_cap = blocks;
    }
    const size_t max_sticks = (size_t)2 * n;
    if (max_sticks > 0xffffffffull || padded > 0xffffffffll)  // the plan and the scatter keep positions in 32 bits
        return fail(CLOTHGPU_ERR_INVALID, "live-stick compaction supports cloths of up to 2^31 particles");
    if (c->seg_out_cap < max_sticks) {
        cudaFree(c->seg_out);
        c->seg_out = nullptr;
        c->seg_out_cap = 0;
        CU_TRY(cudaMalloc(&c->seg_out, std::max<size_t>(1, max_sticks) * 33 + 64));
        c->seg_out_cap = max_sticks;
    }
    float4* d_main = (float4*)c->seg_out;
    uint2* d_idx = (uint2*)(d_main + max_sticks);
    uint8_t* d_hl = (uint8_t*)(d_main + 2 * max_sticks);
    unsigned long long* d_total = c->seg_total;
    size_t list_cap = max_sticks;
    if (direct) {  // the kernel writes the caller's pinned memory itself
        if (direct->main) d_main = direct->main;
        if (direct->hl) d_hl = direct->hl;
        if (direct->idx) d_idx = direct->idx;
        d_total = direct->total, list_cap = direct->cap;
    }
    const bool self_plan = blocks <= kSegSelfPlanTiles;  // one launch: the export blocks plan for themselves
    if (!self_plan) {
        unsigned int* const d_done = reinterpret_cast<unsigned int*>(c->seg_total + 1);  // zero between exports
        segment_plan_kernel<<<std::min(blocks, (unsigned)c->sm_count * (unsigned)SEG_PLAN_BLOCKS_PER_SM), kSegThreads, 0, c->stream>>>(c->FL[c->cur_pos], g, cloth, c->seg_plan, c->seg_starts, blocks, d_done,
                                                                   d_total);
    }
    auto run = [&](auto tag) {
        using T = decltype(tag);
        Planes<T> P = planes<T>(c, c->cur_pos, c->cur_prev);
        auto go = [&](auto kernel, uint2* d_i) {
            if (seg_smem_bytes<T>() + 24 * 1024 > 48 * 1024)  // (static + dynamic above the default limit: opt in)
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)seg_smem_bytes<T>());
            kernel<<<blocks, kSegThreads, seg_smem_bytes<T>(), c->stream>>>(P, g, cloth, self_plan ? nullptr : c->seg_plan, c->seg_starts, line_width,
                                                                              d_main, d_hl, d_i, list_cap, blocks, d_total);
        };
        if (quads)
            go(segment_export_kernel<T, true, false>, nullptr);
        else if (with_idx)
            go(segment_export_kernel<T, false, true>, d_idx);
        else
            go(segment_export_kernel<T, false, false>, nullptr);
    };
    if (c->precision == CLOTHGPU_F32)
        run(float{});
    else
        run(double{});
    CU_TRY(cudaGetLastError());
    c->launches += self_plan ? 1 : 2;
    if (max_sticks_out) *max_sticks_out = max_sticks;
    return CLOTHGPU_OK;
}

// Copy an export back.  Small lists (GUI-sized cloths) are copied at the caller's full capacity together with the count
// and synchronised once — the count is then only checked afterwards; large ones wait for the count and copy exactly n.
// GUI-sized lists skip the copies altogether when the caller's buffers are pinned host memory: the export kernel stores
// straight into them over PCIe (32-byte stores) and writes the list's length next to the counters; one synchronisation.
// Returns false when the buffers do not qualify (pageable memory, alignment, a list too long to be worth it).
constexpr size_t kDirectExportBytes = (size_t)2 << 20;
bool direct_target(clothgpu* c, void* main_dst, void* hl_dst, void* idx_dst, size_t cap, size_t max_sticks, ExportTarget* out) {
    if (!main_dst || !c->direct_export) return false;
    if (std::min(cap, max_sticks) * 33 > kDirectExportBytes) return false;
    auto device_view = [](void* host, size_t align, void** dev) {
        if (!host) return true;  // (not wanted: the kernel writes the handle's own buffer instead)
        cudaPointerAttributes at{};
        if (cudaPointerGetAttributes(&at, host) != cudaSuccess) {
            cudaGetLastError();
            return false;
        }
        if (at.type != cudaMemoryTypeHost || !at.devicePointer || ((uintptr_t)at.devicePointer % align)) return false;
        *dev = at.devicePointer;
        return true;
    };
    void *m = nullptr, *h = nullptr, *i = nullptr;
    if (!device_view(main_dst, 32, &m) || !device_view(hl_dst, 16, &h) || !device_view(idx_dst, 8, &i)) return false;
    out->main = (float4*)m, out->hl = (uint8_t*)h, out->idx = (uint2*)i;
    out->total = (unsigned long long*)(c->ctr_host + 2 * sizeof(DevCounters) + 5 * sizeof(unsigned long long));  // (pinned: one address on both sides)
    out->cap = cap;
    return true;
}
int finish_direct(clothgpu* c, size_t cap, size_t* n_out) {
    CU_TRY(cudaStreamSynchronize(c->stream));
    const unsigned long long total = *(unsigned long long*)(c->ctr_host + 2 * sizeof(DevCounters) + 5 * sizeof(unsigned long long));
    *n_out = (size_t)total;
    if (total > cap) return fail(CLOTHGPU_ERR_CAPACITY, "%llu live sticks, capacity %zu", total, cap);
    return CLOTHGPU_OK;
}

struct ExportCopy {
    void* dst;
    const void* src;
    size_t bytes_per_stick;
};
int copy_export(clothgpu* c, size_t max_sticks, size_t cap, const ExportCopy* copies, int n_copies, size_t* n_out) {
    unsigned long long* total_host = (unsigned long long*)(c->ctr_host + 2 * sizeof(DevCounters) + 5 * sizeof(unsigned long long));
    CU_TRY(cudaMemcpyAsync(total_host, c->seg_total, sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
    bool any = false;
    for (int k = 0; k < n_copies; ++k) any = any || copies[k].dst != nullptr;
    const size_t spec = std::min(cap, max_sticks);
    const bool speculative = any && spec * 33 <= (size_t)(4u << 20);
    if (speculative)
        for (int k = 0; k < n_copies; ++k)
            if (copies[k].dst && spec)
                CU_TRY(cudaMemcpyAsync(copies[k].dst, copies[k].src, spec * copies[k].bytes_per_stick, cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    const unsigned long long total = *total_host;
    *n_out = (size_t)total;
    if (!any) return CLOTHGPU_OK;  // count only
    if (total > cap) return fail(CLOTHGPU_ERR_CAPACITY, "%llu live sticks, capacity %zu", total, cap);
    if (!speculative && total)
        for (int k = 0; k < n_copies; ++k)
            if (copies[k].dst) CU_TRY(cudaMemcpy(copies[k].dst, copies[k].src, (size_t)total * copies[k].bytes_per_stick, cudaMemcpyDeviceToHost));
    return CLOTHGPU_OK;
}
}  // namespace

int clothgpu_compact_segments(clothgpu* c, int32_t cloth) {
    if (int rc = check_handle(c)) return rc;
    return compact_enqueue(c, cloth, false, false, 0.f, nullptr);
}

int clothgpu_read_segments_f32(clothgpu* c, int32_t cloth, float* xyxy, uint8_t* highlighted,
                               uint32_t* sticks_idx, size_t cap, size_t* n_out) {
    if (int rc = check_handle(c)) return rc;
    if (!n_out) return fail(CLOTHGPU_ERR_INVALID, "null n");
    size_t max_sticks = 0;
    ExportTarget direct;
    if (direct_target(c, xyxy, highlighted, sticks_idx, cap, (size_t)2 * c->g.nx * c->g.own_rows, &direct)) {
        if (int rc = compact_enqueue(c, cloth, sticks_idx != nullptr, false, 0.f, &max_sticks, &direct)) return rc;
        return finish_direct(c, cap, n_out);
    }
    if (int rc = compact_enqueue(c, cloth, sticks_idx != nullptr, false, 0.f, &max_sticks)) return rc;
    float4* d_main = (float4*)c->seg_out;
    const ExportCopy copies[3] = {{xyxy, d_main, sizeof(float4)},
                                  {highlighted, (uint8_t*)(d_main + 2 * max_sticks), 1},
                                  {sticks_idx, (uint2*)(d_main + max_sticks), sizeof(uint2)}};
    return copy_export(c, max_sticks, cap, copies, 3, n_out);
}

int clothgpu_host_alloc(size_t bytes, void** p) {
    if (!p) return fail(CLOTHGPU_ERR_INVALID, "null p");
    *p = nullptr;
    const cudaError_t e = cudaHostAlloc(p, std::max<size_t>(bytes, 1), cudaHostAllocPortable | cudaHostAllocMapped);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(e == cudaErrorMemoryAllocation ? CLOTHGPU_ERR_NOMEM : (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver) ? CLOTHGPU_ERR_NODEVICE : CLOTHGPU_ERR_CUDA,
                    "cudaHostAlloc(%zu bytes): %s", bytes, cudaGetErrorString(e));
    }
    return CLOTHGPU_OK;
}
int clothgpu_host_free(void* p) {
    if (!p) return CLOTHGPU_OK;
    const cudaError_t e = cudaFreeHost(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail(CLOTHGPU_ERR_INVALID, "cudaFreeHost: %s (not memory of clothgpu_host_alloc?)", cudaGetErrorString(e));
    }
    return CLOTHGPU_OK;
}

int clothgpu_read_quads_f32(clothgpu* c, int32_t cloth, float line_width, float* quads, uint8_t* highlighted,
                            size_t cap, size_t* n_out) {
    if (int rc = check_handle(c)) return rc;
    if (!n_out) return fail(CLOTHGPU_ERR_INVALID, "null n");
    if (!(line_width >= 0.f) || !(line_width <= 1e30f)) return fail(CLOTHGPU_ERR_INVALID, "line width must be finite and >= 0");
    size_t max_sticks = 0;
    ExportTarget direct;
    if (direct_target(c, quads, highlighted, nullptr, cap, (size_t)2 * c->g.nx * c->g.own_rows, &direct)) {
        if (int rc = compact_enqueue(c, cloth, false, true, line_width, &max_sticks, &direct)) return rc;
        return finish_direct(c, cap, n_out);
    }
    if (int rc = compact_enqueue(c, cloth, false, true, line_width, &max_sticks)) return rc;
    float4* d_main = (float4*)c->seg_out;
    const ExportCopy copies[2] = {{quads, d_main, 2 * sizeof(float4)}, {highlighted, (uint8_t*)(d_main + 2 * max_sticks), 1}};
    return copy_export(c, max_sticks, cap, copies, 2, n_out);
}

int clothgpu_get_counters(clothgpu* c, clothgpu_counters* out) {
    if (int rc = check_handle(c)) return rc;
    if (!out) return fail(CLOTHGPU_ERR_INVALID, "null out");
    const Geom& g = c->g;
    DevCounters* h2 = (DevCounters*)c->ctr_host;  // both slots
    unsigned long long* h5 = (unsigned long long*)(c->ctr_host + 2 * sizeof(DevCounters));
    const bool census = !c->census_valid;
    if (census) {
        // the flag bytes changed behind the frame kernels' back (create, reset, write_state): count them once
        CU_TRY(cudaMemsetAsync(c->census, 0, 5 * sizeof(unsigned long long), c->stream));
        const long long chunks = (long long)g.own_rows * g.pitch / 16;
        dim3 grid((unsigned)std::max<long long>(1, std::min<long long>((chunks + kStreamThreads - 1) / kStreamThreads, c->sm_count * 4)), g.ensemble);
        if (c->precision == CLOTHGPU_F32)
            count_kernel<float><<<grid, kStreamThreads, 0, c->stream>>>(planes<float>(c, c->cur_pos, c->cur_prev), g, c->census);
        else
            count_kernel<double><<<grid, kStreamThreads, 0, c->stream>>>(planes<double>(c, c->cur_pos, c->cur_prev), g, c->census);
        CU_TRY(cudaGetLastError());
        c->launches++;
        CU_TRY(cudaMemcpyAsync(h5, c->census, 5 * sizeof(unsigned long long), cudaMemcpyDeviceToHost, c->stream));
        // the census includes everything the queued frames did: their deltas start again from zero
        static_assert(offsetof(DevCounters, inactive_new) == offsetof(DevCounters, pinned_new) + sizeof(unsigned long long), "layout");
        for (int k = 0; k < 2; ++k) CU_TRY(cudaMemsetAsync(&c->ctr_base[k].pinned_new, 0, 2 * sizeof(unsigned long long), c->stream));
    }
    CU_TRY(cudaMemcpyAsync(h2, c->ctr_base, 2 * sizeof(DevCounters), cudaMemcpyDeviceToHost, c->stream));
    CU_TRY(cudaStreamSynchronize(c->stream));
    // cumulative fields: the sum of the two slots; per-frame fields: the slot of the last frame
    DevCounters sum = h2[c->ctr - c->ctr_base];
    {
        const DevCounters& o = h2[1 - (c->ctr - c->ctr_base)];
        sum.relaxations += o.relaxations, sum.torn_total += o.torn_total, sum.pinned_new += o.pinned_new, sum.inactive_new += o.inactive_new;
    }
    const DevCounters* hc = &sum;
    if (census) {
        c->base.live = h5[0];
        c->base.alive = h5[1];
        c->base.pinned = h5[2];
        c->base.inactive = (unsigned long long)g.nx * g.own_rows * g.ensemble - h5[3];  // the kernel counts ACTIVE particles
        c->base.highlighted = h5[4];
        c->base.torn_total = hc->torn_total;
        c->base.frames = c->frames;
        c->census_valid = true;
    }
    const bool stepped = c->frames > c->base.frames;  // frames since the census: the kernels' own counts stand
    out->frames = c->frames;
    out->live_sticks = stepped ? hc->live_last : c->base.live;
    out->alive_sticks = c->base.alive - (hc->torn_total - c->base.torn_total);
    out->pinned = c->base.pinned + hc->pinned_new;
    out->inactive = c->base.inactive + hc->inactive_new;
    out->highlighted = stepped ? hc->hl_last : c->base.highlighted;
    out->torn_last = hc->torn_last;
    out->relaxations = hc->relaxations;
    return CLOTHGPU_OK;
}

int clothgpu_halo_region(clothgpu* c, int32_t plane, int32_t which, void** dev_ptr, size_t* bytes) {
    if (!c || !dev_ptr || !bytes) return fail(CLOTHGPU_ERR_INVALID, "null argument");
    if (plane < 0 || plane > 4 || which < 0 || which > 3) return fail(CLOTHGPU_ERR_INVALID, "bad plane/which");
    const Geom& g = c->g;
    const int above = g.own_row0 - g.held_row0;
    const int below = (g.held_row0 + g.held_rows) - (g.own_row0 + g.own_rows);
    int row = 0, rows = 0;  // local (held) row index
    switch (which) {
        case 0: row = above; rows = std::min(c->ghost, g.own_rows); break;                      // my top edge
        case 1: rows = std::min(c->ghost, g.own_rows); row = above + g.own_rows - rows; break;  // my bottom edge
        case 2: row = 0; rows = above; break;                                                   // top ghost
        case 3: row = above + g.own_rows; rows = below; break;                                  // bottom ghost
    }
    const size_t es = plane == 4 ? 1 : c->elem;
    char* base = nullptr;
    switch (plane) {
        case 0: base = (char*)c->X[c->cur_pos]; break;
        case 1: base = (char*)c->Y[c->cur_pos]; break;
        case 2: base = (char*)c->PX[c->cur_prev]; break;
        case 3: base = (char*)c->PY[c->cur_prev]; break;
        case 4: base = (char*)c->FL[c->cur_pos]; break;
    }
    *dev_ptr = base + (size_t)row * g.pitch * es;
    *bytes = (size_t)rows * g.pitch * es;
    return CLOTHGPU_OK;
}

namespace {
struct StripBlob {  // what a strip tells its neighbours (clothgpu_strip_export)
    uint32_t magic, version;
    uint64_t pid;
    int32_t device;
    int32_t precision;
    uint64_t raw_base;  // valid inside the exporting process
    cudaIpcMemHandle_t ipc;
    uint64_t off_plane[10];
    uint64_t off_slots;
    int64_t cloth_stride;
    int32_t nx, ny, pitch, ensemble;
    int32_t held_row0, held_rows, own_row0, own_rows;
};
constexpr uint32_t kBlobMagic = 0x434c5347u;  // "CLSG"
}  // namespace

int clothgpu_strip_export(clothgpu* c, void* blob, size_t cap, size_t* len) {
    if (int rc = check_handle(c)) return rc;
    if (!len) return fail(CLOTHGPU_ERR_INVALID, "null len");
    *len = sizeof(StripBlob);
    if (!blob) return CLOTHGPU_OK;  // size query
    if (cap < sizeof(StripBlob)) return fail(CLOTHGPU_ERR_CAPACITY, "blob needs %zu bytes", sizeof(StripBlob));
    StripBlob b{};
    b.magic = kBlobMagic;
    b.version = 1;
    b.pid = (uint64_t)getpid();
    b.device = c->device;
    b.precision = c->precision;
    b.raw_base = (uint64_t)(uintptr_t)c->slab;
    CU_TRY(cudaIpcGetMemHandle(&b.ipc, c->slab));
    for (int k = 0; k < 10; ++k) b.off_plane[k] = c->off_plane[k];
    b.off_slots = c->off_slots;
    b.cloth_stride = c->g.cloth_stride;
    b.nx = c->g.nx, b.ny = c->ny, b.pitch = c->g.pitch, b.ensemble = c->g.ensemble;
    b.held_row0 = c->g.held_row0, b.held_rows = c->g.held_rows, b.own_row0 = c->g.own_row0, b.own_rows = c->g.own_rows;
    memcpy(blob, &b, sizeof b);
    return CLOTHGPU_OK;
}

int clothgpu_strip_attach(clothgpu* c, const void* up_blob, const void* down_blob) {
    if (int rc = check_handle(c)) return rc;
    if (c->desc.strip_rows <= 0) return fail(CLOTHGPU_ERR_STATE, "not a strip handle (desc.strip_rows == 0)");
    if (attached(c)) return fail(CLOTHGPU_ERR_STATE, "already attached");
    const Geom& g = c->g;
    const void* blobs[2] = {up_blob, down_blob};
    clothgpu::Link links[2];
    for (int side = 0; side < 2; ++side) {
        if (!blobs[side]) continue;
        StripBlob b;
        memcpy(&b, blobs[side], sizeof b);
        if (b.magic != kBlobMagic || b.version != 1) return fail(CLOTHGPU_ERR_INVALID, "bad strip blob");
        if (b.nx != g.nx || b.ny != c->ny || b.pitch != g.pitch || b.ensemble != g.ensemble || b.precision != c->precision)
            return fail(CLOTHGPU_ERR_INVALID, "neighbour strip belongs to a different cloth (nx/ny/pitch/ensemble/precision)");
        const int my0 = g.own_row0, my1 = g.own_row0 + g.own_rows;
        const int p_own0 = b.own_row0, p_own1 = b.own_row0 + b.own_rows;
        if (side == 0 ? p_own1 != my0 : p_own0 != my1)
            return fail(CLOTHGPU_ERR_INVALID, "%s neighbour owns rows [%d,%d), mine are [%d,%d): not adjacent",
                        side == 0 ? "upper" : "lower", p_own0, p_own1, my0, my1);
        clothgpu::Link& l = links[side];
        // the neighbour's ghost rows on my side must all be rows I own, and mine must all be rows it owns
        const int ph0 = b.held_row0, ph1 = b.held_row0 + b.held_rows;
        l.row_lo = std::max(my0, ph0);
        l.row_hi = std::min(my1, ph1);
        const int want = side == 0 ? ph1 - p_own1 : p_own0 - ph0;  // ghost rows the neighbour keeps on my side
        if (l.row_hi - l.row_lo != want)
            return fail(CLOTHGPU_ERR_INVALID, "strip of %d rows is thinner than the neighbour's %d ghost rows", g.own_rows, want);
        const int mine_there = side == 0 ? my0 - g.held_row0 : (g.held_row0 + g.held_rows) - my1;
        if (mine_there > b.own_rows)
            return fail(CLOTHGPU_ERR_INVALID, "neighbour strip of %d rows is thinner than my %d ghost rows", b.own_rows, mine_there);
        if (b.pid == (uint64_t)getpid()) {
            l.base = (char*)(uintptr_t)b.raw_base;
            if (b.device != c->device) {
                int can = 0;
                CU_TRY(cudaDeviceCanAccessPeer(&can, c->device, b.device));
                if (!can) return fail(CLOTHGPU_ERR_CUDA, "device %d cannot access device %d", c->device, b.device);
                cudaError_t e = cudaDeviceEnablePeerAccess(b.device, 0);
                if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) CU_TRY(e);
                cudaGetLastError();
            }
        } else {
            void* base = nullptr;
            CU_TRY(cudaIpcOpenMemHandle(&base, b.ipc, cudaIpcMemLazyEnablePeerAccess));
            l.base = (char*)base;
            l.ipc = true;
        }
        for (int k = 0; k < 10; ++k) l.off_plane[k] = (size_t)b.off_plane[k];
        l.off_slots = (size_t)b.off_slots;
        l.cloth_stride = b.cloth_stride;
        l.held_row0 = b.held_row0;
        l.held_rows = b.held_rows;
        l.on = true;
    }
    // epoch slots start at 0 (slab is zero-filled at create); make sure nothing is in flight
    CU_TRY(cudaStreamSynchronize(c->stream));
    c->link[0] = links[0];
    c->link[1] = links[1];
    c->epoch = 0;
    return CLOTHGPU_OK;
}

int clothgpu_strip_push(clothgpu* c) {
    if (int rc = check_handle(c)) return rc;
    if (!attached(c)) return fail(CLOTHGPU_ERR_STATE, "strip is not attached to its neighbours");
    if (int rc = strip_begin_op(c)) return rc;
    auto run = [&](auto tag) -> int {
        using T = decltype(tag);
        Planes<T> mine = planes<T>(c, c->cur_pos, c->cur_prev);
        for (int side = 0; side < 2; ++side) {
            if (!c->link[side].on) continue;
            const PeerOut<T> p = peer_out<T>(c, side, c->cur_pos, c->cur_prev);
            const long long n = (long long)(p.row_hi - p.row_lo) * c->g.pitch;
            dim3 grid((unsigned)std::min<long long>((n + kStreamThreads - 1) / kStreamThreads, 4096), c->g.ensemble);
            strip_push_kernel<T><<<grid, kStreamThreads, 0, c->stream>>>(mine, c->g, p);
            CU_TRY(cudaGetLastError());
            c->launches++;
        }
        return CLOTHGPU_OK;
    };
    if (int rc = c->precision == CLOTHGPU_F32 ? run(float{}) : run(double{})) return rc;
    return strip_end_op(c);
}

int clothgpu_set_stream(clothgpu* c, void* cuda_stream) {
    if (int rc = check_handle(c)) return rc;
    CU_TRY(cudaStreamSynchronize(c->stream));
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    c->stream = (cudaStream_t)cuda_stream;
    c->own_stream = false;
    return CLOTHGPU_OK;
}

int clothgpu_last_step_ms(clothgpu* c, float* ms) {
    if (int rc = check_handle(c)) return rc;
    if (!ms) return fail(CLOTHGPU_ERR_INVALID, "null ms");
    if (!c->timed) return fail(CLOTHGPU_ERR_STATE, "no step has been issued yet");
    CU_TRY(cudaEventSynchronize(c->ev1));
    CU_TRY(cudaEventElapsedTime(ms, c->ev0, c->ev1));
    return CLOTHGPU_OK;
}

int clothgpu_launch_count(clothgpu* c, uint64_t* launches) {
    if (!c || !launches) return fail(CLOTHGPU_ERR_INVALID, "null argument");
    *launches = c->launches;
    return CLOTHGPU_OK;
}

int clothgpu_last_schedule(clothgpu* c, int32_t* schedule, const char** kernel_name) {
    if (!c) return fail(CLOTHGPU_ERR_INVALID, "null handle");
    if (c->last_sched < 0 || !c->last_kernel) return fail(CLOTHGPU_ERR_STATE, "no step has been issued yet");
    if (schedule) *schedule = c->last_sched;
    if (kernel_name) *kernel_name = c->last_kernel;
    return CLOTHGPU_OK;
}

int clothgpu_selftest_stickmath(int32_t device, uint64_t seed, uint64_t samples, const double* s_list, size_t n_list,
                                clothgpu_stickmath_report* report) {
    if (!report) return fail(CLOTHGPU_ERR_INVALID, "null report");
    static_assert(sizeof(StickMathReport) == sizeof(clothgpu_stickmath_report), "report layouts must agree");
    memset(report, 0, sizeof *report);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CLOTHGPU_ERR_NODEVICE, "no CUDA device (%s); clothgpu has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(CLOTHGPU_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(CLOTHGPU_ERR_NODEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    CU_TRY(cudaSetDevice(device));
    for (size_t i = 0; i < n_list; ++i)
        if (!s_list || !(s_list[i] >= 1.0) || !(s_list[i] < 0x1p900))
            return fail(CLOTHGPU_ERR_INVALID, "s_list[%zu] is outside [1, 2^900)", i);
    StickMathReport* d_rep = nullptr;
    double* d_list = nullptr;
    CU_TRY(cudaMalloc((void**)&d_rep, sizeof(StickMathReport)));
    int rc = CLOTHGPU_OK;
    auto step = [&](cudaError_t ee, const char* what) {
        if (rc == CLOTHGPU_OK && ee != cudaSuccess) rc = fail(CLOTHGPU_ERR_CUDA, "%s: %s", what, cudaGetErrorString(ee));
    };
    step(cudaMemset(d_rep, 0, sizeof(StickMathReport)), "cudaMemset");
    const double gain = 0.35;  // physics/constraint.go:42
    if (rc == CLOTHGPU_OK && samples) {
        stickmath_random_kernel<<<prop.multiProcessorCount * 8, 256>>>(seed, samples, gain, d_rep);
        step(cudaGetLastError(), "stickmath_random_kernel");
    }
    if (rc == CLOTHGPU_OK && n_list) {
        step(cudaMalloc((void**)&d_list, n_list * sizeof(double)), "cudaMalloc");
        if (rc == CLOTHGPU_OK) step(cudaMemcpy(d_list, s_list, n_list * sizeof(double), cudaMemcpyHostToDevice), "cudaMemcpy");
        if (rc == CLOTHGPU_OK) {
            const unsigned blocks = (unsigned)std::min<size_t>((n_list * 64 + 255) / 256, (size_t)prop.multiProcessorCount * 8);
            stickmath_list_kernel<<<blocks, 256>>>(d_list, n_list, gain, d_rep);
            step(cudaGetLastError(), "stickmath_list_kernel");
        }
    }
    step(cudaDeviceSynchronize(), "self-test kernels");
    if (rc == CLOTHGPU_OK) step(cudaMemcpy(report, d_rep, sizeof *report, cudaMemcpyDeviceToHost), "cudaMemcpy");
    cudaFree(d_list);
    cudaFree(d_rep);
    return rc;
}

int clothgpu_selftest_stickmath_f32(int32_t device, clothgpu_stickmath_report* report) {
    if (!report) return fail(CLOTHGPU_ERR_INVALID, "null report");
    memset(report, 0, sizeof *report);
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(CLOTHGPU_ERR_NODEVICE, "no CUDA device (%s); clothgpu has no CPU fallback",
                    e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(CLOTHGPU_ERR_INVALID, "device %d out of range (have %d)", device, ndev);
    cudaDeviceProp prop;
    CU_TRY(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) return fail(CLOTHGPU_ERR_NODEVICE, "device %d is sm_%d%d; this library carries sm_100a code only", device, prop.major, prop.minor);
    CU_TRY(cudaSetDevice(device));
    StickMathReport* d_rep = nullptr;
    CU_TRY(cudaMalloc((void**)&d_rep, sizeof(StickMathReport)));
    cudaError_t ee = cudaMemset(d_rep, 0, sizeof(StickMathReport));
    if (ee == cudaSuccess) {
        stickmath_f32_exhaustive_kernel<<<prop.multiProcessorCount * 16, 256>>>(0.35f, d_rep);
        ee = cudaGetLastError();
    }
    if (ee == cudaSuccess) ee = cudaDeviceSynchronize();
    if (ee == cudaSuccess) ee = cudaMemcpy(report, d_rep, sizeof *report, cudaMemcpyDeviceToHost);
    cudaFree(d_rep);
    if (ee != cudaSuccess) return fail(CLOTHGPU_ERR_CUDA, "binary32 self-test: %s", cudaGetErrorString(ee));
    return CLOTHGPU_OK;
}

const char* clothgpu_last_error(void) { return g_last_error.c_str(); }

const char* clothgpu_version(void) { return "clothgpu 0.1 (sm_100a)"; }

}  // extern "C"
