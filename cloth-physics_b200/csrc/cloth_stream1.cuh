// cloth_stream1.cuh — the reference's own frame (ONE relaxation sweep, physics/cloth.go:96-100) as a single streaming
// pass over HBM: no block barriers, every warp independent.
//
// With one sweep per frame the launch is bound by HBM, not by arithmetic, so the job is to keep loads in flight.
// A lane owns NP (even, odd) column pairs of every plane row — one pair: 16 bytes in binary64, 8 in binary32; the code is
// written for any NP — so a warp owns a window of 64 NP columns, contiguous per warp and plane, and marches down a chunk
// of rows two at a time, holding a sliding window of three rows in registers:
//
//     load rows (r, r+1)  ->  particle_update (cloth.go:92-94; previous positions go straight back to HBM)
//       H-even   the lane's own pair(s)                                          in registers
//       H-odd    odd column of a pair  <->  even column of the next pair         in registers inside the lane,
//                                                                                two warp shuffles each way across lanes
//       V-even   rows (r, r+1)                                                   in registers
//       V-odd    rows (r-1, r), r-1 carried from the previous trip               in registers
//     store rows (r-1, r), carry row r+1
//
// so every particle is read once and written once.  The colour order H-even, H-odd, V-even, V-odd of the whole-cloth
// pass is respected because each stick only needs its two endpoints to have finished the previous colours, which the
// row pipeline guarantees (V-even of a pair after both H passes of both rows, V-odd after the V-even of both pairs it
// joins).  Column halo: lane 0's first column misses its H-odd stick to the left and lane 31's last column the one to
// the right, so windows advance by 30 lanes' worth of columns and lanes 0 / 31 are recomputed by the neighbouring warp
// (except at the cloth's edge, where those sticks do not exist).  Row halo: a chunk reads two rows above and below its
// owned rows.  Loads run two trips ahead of the arithmetic: every lane owns a small ring of shared-memory slots that
// cp.async fills (a lane only reads back what it copied itself, so this needs no barrier and no registers for the data in
// flight).
#pragma once
#include "cloth_fused.cuh"
#include "cloth_kernels.cuh"
#include "cloth_types.cuh"

namespace clothgpu_impl {

#ifndef S1_MIN_CTAS
#define S1_MIN_CTAS 4
#endif
#ifndef S1_RING_CTAS
#define S1_RING_CTAS 5
#endif
constexpr int kS1Threads = 128;       // 4 independent warps per CTA
#ifndef S1_PINS_KNOWN
#define S1_PINS_KNOWN 0  // 1: one vote per trip instead of one per pass on "is any endpoint pinned" — measured 3.11 -> 3.07 ms with
#endif                   // a drag but 3.05 -> 3.07 ms idle on 16384^2, and 0.2167 -> 0.2236 ms idle on 4096^2: not taken
constexpr int kS1Pins = S1_PINS_KNOWN ? kPinsKnown : kPinsVote;
#ifndef S1_F32_COMMON_FIRST
#define S1_F32_COMMON_FIRST 1  // (binary32 takes the branch-free common case first: 2.10 vs 2.31 ms idle, 2.23 vs 2.34 ms with a drag on 16384^2)
#endif
#ifndef S1_PIN_VOTE
#define S1_PIN_VOTE 2  // 0: never, 1: every instantiation, 2: frames with a drag only
#endif
#ifndef S1_LATE_CLAMP
#define S1_LATE_CLAMP 1  // (measured on 16384^2: 3.10 -> 3.07 ms idle, 3.23 -> 3.19 ms with the drag; 4096^2 0.2177 -> 0.2139 ms)
#endif
#ifndef S1_UPDATE_ALL
#define S1_UPDATE_ALL 1  // (measured on 16384^2: 3.06 -> 3.00 ms idle, 3.17 -> 3.12 ms with the drag; 0 = every lane tests whether its particle exists first)
#endif
#ifndef S1_UNROLL
#define S1_UNROLL 1
#endif
constexpr int kS1Unroll = S1_UNROLL;  // trips per pass of the row loop (2 removes the register moves of the loop-carried
                                      // rows but measured 3.93 ms at 96 registers — spills — and 3.37 ms at 120 registers / 4 CTAs, against 3.20 ms)

// What a lane owns of one plane row: 16 bytes = NP (even, odd) column pairs.
template <typename T>
struct S1Lane;
template <>
struct S1Lane<double> {
    using vec = double2;
    static constexpr int NP = 1;
    static __device__ __forceinline__ double get(const vec& v, int i) { return i == 0 ? v.x : v.y; }
    static __device__ __forceinline__ void set(vec& v, int i, double a) { (i == 0 ? v.x : v.y) = a; }
    static __device__ __forceinline__ uint32_t load_flags(const uint8_t* p) { return *reinterpret_cast<const unsigned short*>(p); }
    static __device__ __forceinline__ void store_flags(uint8_t* p, uint32_t f) { *reinterpret_cast<unsigned short*>(p) = (unsigned short)f; }
};
// binary32: also ONE pair per lane (8 bytes).  Two pairs per lane (float4, `-DS1_F32_TWO_PAIRS`, which the kernel below
// supports) measured 2.43 vs 2.49 ms per frame on 16384^2 with the branch-free binary32 stick arithmetic (and 4.80 vs
// 3.77 ms with the library's branchy IEEE sequences before it): the binary32 path is bound by issue slots, not by bytes.
template <>
struct S1Lane<float> {
#ifdef S1_F32_TWO_PAIRS
    using vec = float4;
    static constexpr int NP = 2;
    static __device__ __forceinline__ float get(const vec& v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }
    static __device__ __forceinline__ void set(vec& v, int i, float a) { (i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w) = a; }
    static __device__ __forceinline__ uint32_t load_flags(const uint8_t* p) { return *reinterpret_cast<const uint32_t*>(p); }
    static __device__ __forceinline__ void store_flags(uint8_t* p, uint32_t f) { *reinterpret_cast<uint32_t*>(p) = f; }
#else
    using vec = float2;
    static constexpr int NP = 1;
    static __device__ __forceinline__ float get(const vec& v, int i) { return i == 0 ? v.x : v.y; }
    static __device__ __forceinline__ void set(vec& v, int i, float a) { (i == 0 ? v.x : v.y) = a; }
    static __device__ __forceinline__ uint32_t load_flags(const uint8_t* p) { return *reinterpret_cast<const unsigned short*>(p); }
    static __device__ __forceinline__ void store_flags(uint8_t* p, uint32_t f) { *reinterpret_cast<unsigned short*>(p) = (unsigned short)f; }
#endif
};
template <typename T>
__host__ __device__ constexpr int s1_window_cols() { return 64 * S1Lane<T>::NP; }   // columns a warp loads
template <typename T>
__host__ __device__ constexpr int s1_window_step() { return 60 * S1Lane<T>::NP; }   // columns a window advances by (30 lanes owned + 1 halo lane per side)
constexpr int kS1WindowStep = 60;     // (binary64; kept for the host's AUTO heuristics)

struct Stream1Plan {
    int windows;     // column windows per cloth
    int chunk_rows;  // owned rows per chunk (even)
    int chunks;      // row chunks per cloth (strip)
    long long tasks; // windows * chunks * ensemble
};

template <typename T>
struct RowRaw {  // one row of a lane's columns, as loaded
    typename S1Lane<T>::vec x, y, px, py;
    uint32_t fl;  // byte c = column c of the lane
};

// RING == 0: the next trip's loads wait in registers.  RING >= 2: every lane owns RING slots of 8 x 16 B in shared memory
// that cp.async fills RING trips ahead (a lane only ever reads back what it copied itself, so no barrier is involved):
// more loads in flight, 34 fewer live registers.
template <typename T, bool PER_CLOTH, bool DRAG, int RING>
__global__ void __launch_bounds__(kS1Threads, RING ? S1_RING_CTAS : S1_MIN_CTAS)
    stream1_frame_kernel(Planes<T> in, Planes<T> out, Geom g, Stream1Plan pl, const FrameU<T> u0,
                         const FrameU<T>* __restrict__ per_cloth, DevCounters* ctr, const PeerOut<T> peer_up,
                         const PeerOut<T> peer_dn, const StripSync ss) {
    using V2 = typename Vec2<T>::type;          // one particle's (x, y)
    using L = S1Lane<T>;
    using VL = typename L::vec;                 // 16 bytes of one plane row
    constexpr int NP = L::NP, NC = 2 * NP;      // column pairs / columns per lane
    // The particle loop as "branch-free common case first, the rule itself only when a lane needs it" (below): +10 % in
    // binary32, whose launch is bound by issue slots — and -20 % to -32 % in binary64 (3.88 / 4.17 vs 3.15-3.19 ms on
    // 16384^2 with the slots refilled after / before the arithmetic): keeping the inputs for the rare path costs 16
    // registers the binary64 kernel does not have (88 bytes of spills in the loop).  So: binary32 only.
    constexpr bool kCommonFirst = S1_F32_COMMON_FIRST && (sizeof(T) == 4);
    constexpr bool kEarlyRefill = sizeof(T) == 8;  // binary64: a trip's slots are refilled before its arithmetic
    constexpr unsigned FULL = 0xffffffffu;
    const int lane = threadIdx.x & 31;
    prepare_next_frame_counters(ctr);
    const long long task = (long long)blockIdx.x * (kS1Threads / 32) + (threadIdx.x >> 5);
    if (task >= pl.tasks) return;  // whole warp
    const int w = (int)(task % pl.windows);
    // Row chunks are issued edges first (top, bottom, then the interior from the top down): the chunks that store into
    // the neighbouring strips' ghost rows — and publish the epoch the neighbours' next frame waits for — then finish in
    // the first wave instead of the last.
    const int ci = (int)((task / pl.windows) % pl.chunks);
    const int chunk = ci == 0 ? 0 : ci == 1 ? pl.chunks - 1 : ci - 1;
    const int e = (int)(task / ((long long)pl.windows * pl.chunks));
    const FrameU<T> u = pick_frame<T, PER_CLOTH>(u0, per_cloth, e);
    const bool dragging = DRAG && (u.buttons & CLOTHGPU_MOUSE_DRAGGING) != 0;

    const int c0 = w * s1_window_step<T>();
    const int ce = c0 + NC * lane;  // my first column
    const bool ex_e = ce < g.nx;    // my columns exist from the left: column ce + c exists iff ce + c < nx
    // lanes whose columns all come out right in this window (see the header): they are stored by this warp
    const bool own_cols = ex_e && (lane >= 1 || w == 0) && (lane <= 30 || c0 + s1_window_cols<T>() >= g.nx);
    const int held_end = g.held_row0 + g.held_rows;
    const int own_end = g.own_row0 + g.own_rows;
    const int ra = g.own_row0 + chunk * pl.chunk_rows, rb = min(ra + pl.chunk_rows, own_end);  // owned rows [ra, rb)
    const int rs = max(g.held_row0, ra - 2), re = min(held_end, rb + 2);                       // rows the pipeline reads
    const long long cloth_base = (long long)e * g.cloth_stride;
    // Attached strips: a chunk that stores rows a neighbour keeps as ghosts (the only chunks that read ghost rows
    // themselves) waits for that neighbour's epoch first; everybody else starts right away (cloth_types.cuh, StripSync).
    const bool st_up = ss.on() && ra < peer_up.row_hi && rb > peer_up.row_lo;
    const bool st_dn = ss.on() && ra < peer_dn.row_hi && rb > peer_dn.row_lo;
    if (st_up) ss.wait(0);
    if (st_dn) ss.wait(1);

    auto load_row = [&](int gy) -> RowRaw<T> {
        RowRaw<T> r;
#pragma unroll
        for (int c = 0; c < NC; ++c) L::set(r.x, c, T(0)), L::set(r.y, c, T(0)), L::set(r.px, c, T(0)), L::set(r.py, c, T(0));
        r.fl = 0;
        if (gy < re && ex_e) {
            const long long at = cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce;
            r.x = *reinterpret_cast<const VL*>(in.x + at);
            r.y = *reinterpret_cast<const VL*>(in.y + at);
            r.px = *reinterpret_cast<const VL*>(in.px + at);
            r.py = *reinterpret_cast<const VL*>(in.py + at);
            r.fl = L::load_flags(in.fl + at);
        }
        return r;
    };
    // E[p] / O[p] = the even / odd column of pair p
    auto store_xy = [&](int gy, const V2 (&pe)[NP], const V2 (&po)[NP], const uint32_t (&fe)[NP], const uint32_t (&fo)[NP]) {
        if (!own_cols || gy < ra || gy >= rb) return;
        const long long at = cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce;
        VL ox2, oy2;
        uint32_t ff = 0;
#pragma unroll
        for (int p = 0; p < NP; ++p) {
            L::set(ox2, 2 * p, pe[p].x), L::set(ox2, 2 * p + 1, po[p].x), L::set(oy2, 2 * p, pe[p].y), L::set(oy2, 2 * p + 1, po[p].y);
            ff |= ((fe[p] & 0xffu) | ((fo[p] & 0xffu) << 8)) << (16 * p);
        }
#if S1_STCS  // streaming stores: the planes are far larger than L2 and are read again only by the NEXT frame
        __stcs(reinterpret_cast<VL*>(out.x + at), ox2);
        __stcs(reinterpret_cast<VL*>(out.y + at), oy2);
#else
        *reinterpret_cast<VL*>(out.x + at) = ox2;
        *reinterpret_cast<VL*>(out.y + at) = oy2;
#endif
        L::store_flags(out.fl + at, ff);
        if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {  // the neighbouring strip's ghost copy of this row
            const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + ce;
            *reinterpret_cast<VL*>(peer_up.x + pat) = ox2;
            *reinterpret_cast<VL*>(peer_up.y + pat) = oy2;
            L::store_flags(peer_up.fl + pat, ff);
        }
        if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
            const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + ce;
            *reinterpret_cast<VL*>(peer_dn.x + pat) = ox2;
            *reinterpret_cast<VL*>(peer_dn.y + pat) = oy2;
            L::store_flags(peer_dn.fl + pat, ff);
        }
    };
    auto store_prev = [&](int gy, const VL& vpx, const VL& vpy) {
        if (!own_cols || gy < ra || gy >= rb) return;
        const long long at = cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce;
#if S1_STCS
        __stcs(reinterpret_cast<VL*>(out.px + at), vpx);
        __stcs(reinterpret_cast<VL*>(out.py + at), vpy);
#else
        *reinterpret_cast<VL*>(out.px + at) = vpx;
        *reinterpret_cast<VL*>(out.py + at) = vpy;
#endif
        if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {
            const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + ce;
            *reinterpret_cast<VL*>(peer_up.px + pat) = vpx;
            *reinterpret_cast<VL*>(peer_up.py + pat) = vpy;
        }
        if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
            const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + ce;
            *reinterpret_cast<VL*>(peer_dn.px + pat) = vpx;
            *reinterpret_cast<VL*>(peer_dn.py + pat) = vpy;
        }
    };
    auto live_of = [](uint32_t fh, uint32_t ft, uint32_t alive) -> uint32_t {
        return ((fh & alive) && (fh & ft & CLOTHGPU_FLAG_ACTIVE)) ? 1u : 0u;  // cloth.go:97
    };
    auto shfl_v2 = [&](const V2& v, bool down) -> V2 {
        V2 r;
        r.x = down ? __shfl_down_sync(FULL, v.x, 1) : __shfl_up_sync(FULL, v.x, 1);
        r.y = down ? __shfl_down_sync(FULL, v.y, 1) : __shfl_up_sync(FULL, v.y, 1);
        return r;
    };

    unsigned int visits = 0, torn = 0;
    ParticleEvents events;
    // the row carried from the previous trip (global row r - 1): after V-even, waiting for its V-odd stick
    V2 cE[NP], cO[NP];
    uint32_t cfe[NP], cfo[NP];
#pragma unroll
    for (int p = 0; p < NP; ++p) cE[p].x = cE[p].y = cO[p].x = cO[p].y = T(0), cfe[p] = cfo[p] = 0;

    // ---- prefetch machinery -------------------------------------------------------------------------------------
    extern __shared__ __align__(16) unsigned char s1_smem[];
    // slot (stage, q, plane) of this lane: 32 lanes x sizeof(VL) contiguous, conflict-free
    auto slot = [&](int stage, int q, int plane) -> VL* {
        return reinterpret_cast<VL*>(s1_smem) + ((((threadIdx.x >> 5) * (RING ? RING : 1) + stage) * 8 + q * 4 + plane) * 32 + lane);
    };
    auto issue_trip = [&](int r, int stage) {  // rows r, r+1 -> stage (RING > 0 only)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int gy = r + q;
            const bool ok = gy < re && ex_e;
            const long long at = ok ? cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce : 0;
            cp_async_lane<sizeof(VL)>(slot(stage, q, 0), in.x + at, ok);
            cp_async_lane<sizeof(VL)>(slot(stage, q, 1), in.y + at, ok);
            cp_async_lane<sizeof(VL)>(slot(stage, q, 2), in.px + at, ok);
            cp_async_lane<sizeof(VL)>(slot(stage, q, 3), in.py + at, ok);
        }
        cp_async_commit();
    };
    // (the value is used RING trips later; nothing may touch it before — masking the non-existent columns right here
    //  made every warp wait for the load at once: 17 % of all stall samples of the launch sat on that one instruction)
    auto load_flags = [&](int gy) -> uint32_t {
        uint32_t f = 0;
        if (gy < re && ex_e) f = L::load_flags(in.fl + cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce);
        return f;
    };
    // columns of the row padding (ce + c >= nx) do not exist: their flag bytes are dropped where the flags are used
    uint32_t fl_keep = 0;
#pragma unroll
    for (int c = 0; c < NC; ++c)
        if (ce + c < g.nx) fl_keep |= 0xffu << (8 * c);

    RowRaw<T> nxt0, nxt1;
    uint32_t fq[RING ? 2 * RING : 1];  // flags of the trips in flight (RING > 0)
    if (RING == 0) {
        nxt0 = load_row(rs), nxt1 = load_row(rs + 1);
    } else {
#pragma unroll
        for (int k = 0; k < RING; ++k) {
            issue_trip(rs + 2 * k, k);
            fq[2 * k] = load_flags(rs + 2 * k), fq[2 * k + 1] = load_flags(rs + 2 * k + 1);
        }
    }
    int stage = 0;
#pragma unroll(kS1Unroll)
    for (int r = rs; r < re; r += 2) {
        RowRaw<T> reg0, reg1;  // RING == 0 only: this trip's rows, in registers
        if (RING == 0) {
            // (a ping-pong of two register sets instead of these copies measured slower: twice the code, worse schedule)
            reg0 = nxt0, reg1 = nxt1;
            reg0.fl &= fl_keep, reg1.fl &= fl_keep;
            nxt0 = load_row(r + 2);  // in flight during this trip's arithmetic (rows >= re load nothing)
            nxt1 = load_row(r + 3);
        } else {
            cp_async_wait<RING - 1>();  // the oldest group (this trip) has landed
            if (kEarlyRefill) {  // binary64: this trip's rows move to registers and their slots are refilled at once
                reg0.x = *slot(stage, 0, 0), reg0.y = *slot(stage, 0, 1), reg0.px = *slot(stage, 0, 2), reg0.py = *slot(stage, 0, 3);
                reg1.x = *slot(stage, 1, 0), reg1.y = *slot(stage, 1, 1), reg1.px = *slot(stage, 1, 2), reg1.py = *slot(stage, 1, 3);
                reg0.fl = fq[0] & fl_keep, reg1.fl = fq[RING ? 1 : 0] & fl_keep;
#pragma unroll
                for (int k = 0; k + 1 < RING; ++k) fq[2 * k] = fq[2 * k + 2], fq[2 * k + 1] = fq[2 * k + 3];
                issue_trip(r + 2 * RING, stage);  // refill the slots just read (always commits, so group counting stays uniform)
                fq[RING ? 2 * RING - 2 : 0] = load_flags(r + 2 * RING), fq[RING ? 2 * RING - 1 : 0] = load_flags(r + 2 * RING + 1);
                stage = stage + 1 == RING ? 0 : stage + 1;
            }
        }
        const uint32_t fl_q[2] = {(RING && !kEarlyRefill) ? fq[0] & fl_keep : reg0.fl, (RING && !kEarlyRefill) ? fq[RING ? 1 : 0] & fl_keep : reg1.fl};
        // row q of the trip as loaded: from the ring slots (they stay valid until the refill below) or from the registers
        auto fetch_row = [&](int q) -> RowRaw<T> {
            RowRaw<T> rw;
            if (RING == 0 || kEarlyRefill) {
                rw = q ? reg1 : reg0;
            } else {
                rw.x = *slot(stage, q, 0), rw.y = *slot(stage, q, 1), rw.px = *slot(stage, q, 2), rw.py = *slot(stage, q, 3);
                rw.fl = fl_q[q];
            }
            return rw;
        };
        V2 E[2][NP], O[2][NP];
        uint32_t fe[2][NP], fo[2][NP];
        // ---- loop over particles, cloth.go:92-94.  kCommonFirst: first the branch-free common case for all particles of
        // ---- the trip (particle_update_common); if any lane of the warp has a pinned particle, one within the mouse's reach
        // ---- or one leaving the window — rare — the rule itself re-runs on the kept inputs.  Otherwise: the rule itself.
        VL vpx[2], vpy[2];
        bool clamp_e[2][NP] = {}, clamp_o[2][NP] = {};  // S1_LATE_CLAMP: particles that left the window in this trip
        bool rare = !kCommonFirst;
        if (kCommonFirst) {
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const RowRaw<T> rw = fetch_row(q);
            vpx[q] = rw.px, vpy[q] = rw.py;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                E[q][p].x = L::get(rw.x, 2 * p), O[q][p].x = L::get(rw.x, 2 * p + 1);
                E[q][p].y = L::get(rw.y, 2 * p), O[q][p].y = L::get(rw.y, 2 * p + 1);
                fe[q][p] = (rw.fl >> (16 * p)) & 0xffu, fo[q][p] = (rw.fl >> (16 * p + 8)) & 0xffu;
                T px0 = L::get(vpx[q], 2 * p), px1 = L::get(vpx[q], 2 * p + 1), py0 = L::get(vpy[q], 2 * p), py1 = L::get(vpy[q], 2 * p + 1);
                // (rows / columns that do not exist hold zeros and zero flags: whatever the common case computes for them is
                //  discarded below — they are never stored — but they must not drag the warp into the rare path)
                const bool r0 = particle_update_common(E[q][p].x, E[q][p].y, px0, py0, fe[q][p], u);
                const bool r1 = particle_update_common(O[q][p].x, O[q][p].y, px1, py1, fo[q][p], u);
                const bool row_ok = r + q < re;
                rare |= (r0 && row_ok && ce + 2 * p < g.nx) || (r1 && row_ok && ce + 2 * p + 1 < g.nx);
                L::set(vpx[q], 2 * p, px0), L::set(vpx[q], 2 * p + 1, px1), L::set(vpy[q], 2 * p, py0), L::set(vpy[q], 2 * p + 1, py1);
            }
        }
        }
        if (!kCommonFirst || __any_sync(FULL, rare)) {
            auto update_rows = [&](auto may_pin) {
            constexpr bool kMayPin = decltype(may_pin)::value;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const RowRaw<T> rw = fetch_row(q);  // the kept inputs
                vpx[q] = rw.px, vpy[q] = rw.py;
                const bool row_ok = r + q < re && ex_e;
                const bool counted = own_cols && r + q >= ra && r + q < rb;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    E[q][p].x = L::get(rw.x, 2 * p), O[q][p].x = L::get(rw.x, 2 * p + 1);
                    E[q][p].y = L::get(rw.y, 2 * p), O[q][p].y = L::get(rw.y, 2 * p + 1);
                    fe[q][p] = (rw.fl >> (16 * p)) & 0xffu, fo[q][p] = (rw.fl >> (16 * p + 8)) & 0xffu;
#if S1_UPDATE_ALL
                    {   // every lane runs the rule, also on the zeros of particles that do not exist (rows past the chunk, row
                        // padding); those are put back to zeros — coordinates and flag byte — in a block only the warps at
                        // the cloth's right and lower edge enter, so no merge of two register sets precedes the sticks
                        const uint32_t fe_in = fe[q][p], fo_in = fo[q][p];
                        T px0 = L::get(vpx[q], 2 * p), px1 = L::get(vpx[q], 2 * p + 1), py0 = L::get(vpy[q], 2 * p), py1 = L::get(vpy[q], 2 * p + 1);
                        const bool ok_e = row_ok && (NP == 1 || ce + 2 * p < g.nx), ok_o = row_ok && ce + 2 * p + 1 < g.nx;
                        // (the window clamp — an event — is applied to the whole trip below, behind one vote)
                        // (a vote on "the mouse is near one of the warp's particles" in front of :108-149 as well measured
                        //  3.00 vs 3.02 ms on 16384^2 but 0.218 vs 0.215 ms on 4096^2 with a drag: not taken)
                        clamp_e[q][p] = particle_update<T, !S1_LATE_CLAMP, kMayPin>(E[q][p].x, E[q][p].y, px0, py0, fe[q][p], u) && ok_e;
                        clamp_o[q][p] = particle_update<T, !S1_LATE_CLAMP, kMayPin>(O[q][p].x, O[q][p].y, px1, py1, fo[q][p], u) && ok_o;
                        if (!__all_sync(FULL, ok_o)) {  // (ok_o implies ok_e)
                            if (!ok_e) E[q][p].x = E[q][p].y = px0 = py0 = T(0), fe[q][p] = 0;
                            if (!ok_o) O[q][p].x = O[q][p].y = px1 = py1 = T(0), fo[q][p] = 0;
                        }
                        L::set(vpx[q], 2 * p, px0), L::set(vpx[q], 2 * p + 1, px1), L::set(vpy[q], 2 * p, py0), L::set(vpy[q], 2 * p + 1, py1);
                        if (counted) events.note2(fe_in, fe[q][p], fo_in, fo[q][p]);
                    }
#else
                    if (row_ok) {
                        const uint32_t fe_in = fe[q][p], fo_in = fo[q][p];
                        T px0 = L::get(vpx[q], 2 * p), px1 = L::get(vpx[q], 2 * p + 1), py0 = L::get(vpy[q], 2 * p), py1 = L::get(vpy[q], 2 * p + 1);
                        if (NP == 1 || ce + 2 * p < g.nx) particle_update(E[q][p].x, E[q][p].y, px0, py0, fe[q][p], u);
                        if (ce + 2 * p + 1 < g.nx) particle_update(O[q][p].x, O[q][p].y, px1, py1, fo[q][p], u);
                        L::set(vpx[q], 2 * p, px0), L::set(vpx[q], 2 * p + 1, px1), L::set(vpy[q], 2 * p, py0), L::set(vpy[q], 2 * p + 1, py1);
                        if (counted) events.note2(fe_in, fe[q][p], fo_in, fo[q][p]);
                    }
#endif
                }
            }
            };
            // pins are the cloth's top row and what ctrl-clicks add: one vote over the trip's flag bytes decides whether the
            // rule runs with or without its pinned branch.  Measured per instantiation (16384^2 / 4096^2): with a drag 3.22 ->
            // 3.14-3.19 ms / 0.2235 -> 0.2170 ms, without one 3.08 -> 3.07-3.11 ms / 0.2142 -> 0.2206 ms — so only there.
            constexpr bool kPinVote = S1_UPDATE_ALL && (S1_PIN_VOTE == 1 || (S1_PIN_VOTE == 2 && DRAG));
            if (kPinVote && !__any_sync(FULL, ((fl_q[0] | fl_q[1]) & (CLOTHGPU_FLAG_PIN * 0x01010101u)) != 0u)) update_rows(std::false_type{});
            else update_rows(std::true_type{});
#if S1_UPDATE_ALL && S1_LATE_CLAMP
            {
                bool any_clamp = false;
#pragma unroll
                for (int q = 0; q < 2; ++q)
#pragma unroll
                    for (int p = 0; p < NP; ++p) any_clamp |= clamp_e[q][p] | clamp_o[q][p];
                if (__any_sync(FULL, any_clamp)) {  // a particle of the trip left the window (:164-178)
#pragma unroll
                    for (int q = 0; q < 2; ++q)
#pragma unroll
                        for (int p = 0; p < NP; ++p) {
                            T px0 = L::get(vpx[q], 2 * p), px1 = L::get(vpx[q], 2 * p + 1), py0 = L::get(vpy[q], 2 * p), py1 = L::get(vpy[q], 2 * p + 1);
                            if (clamp_e[q][p]) particle_window_clamp(E[q][p].x, E[q][p].y, px0, py0, u);
                            if (clamp_o[q][p]) particle_window_clamp(O[q][p].x, O[q][p].y, px1, py1, u);
                            L::set(vpx[q], 2 * p, px0), L::set(vpx[q], 2 * p + 1, px1), L::set(vpy[q], 2 * p, py0), L::set(vpy[q], 2 * p + 1, py1);
                        }
                }
            }
#endif
        } else {
            // the common case wrote every lane, also the particles that do not exist (zero-filled rows past the chunk, the
            // row padding): back to the zeros they were loaded as (their flag bytes are zero either way)
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const bool row_ok = r + q < re;
#pragma unroll
                for (int c = 0; c < NC; ++c)
                    if (!(row_ok && ce + c < g.nx)) {
                        V2& pt = (c & 1) ? O[q][c >> 1] : E[q][c >> 1];
                        pt.x = pt.y = T(0);
                        L::set(vpx[q], c, T(0)), L::set(vpy[q], c, T(0));
                    }
            }
        }
        if (RING != 0 && !kEarlyRefill) {
#pragma unroll
            for (int k = 0; k + 1 < RING; ++k) fq[2 * k] = fq[2 * k + 2], fq[2 * k + 1] = fq[2 * k + 3];
            issue_trip(r + 2 * RING, stage);  // refill the slots just read (always commits, so group counting stays uniform)
            fq[2 * RING - 2] = load_flags(r + 2 * RING), fq[2 * RING - 1] = load_flags(r + 2 * RING + 1);
            stage = stage + 1 == RING ? 0 : stage + 1;
        }
#pragma unroll
        for (int q = 0; q < 2; ++q)
            if (r + q < re && ex_e) store_prev(r + q, vpx[q], vpy[q]);
        const bool own_r[2] = {r >= ra && r < rb, r + 1 >= ra && r + 1 < rb};
        bool touched = false;
        // pins are the cloth's top row and what ctrl-clicks add: ONE vote per trip over every particle its four passes can
        // touch (the two rows, the row carried over from the trip before; a neighbouring lane's particle is some lane's
        // own) tells the passes whether any of them needs per-endpoint multipliers (constraint.go:46-53)
        bool trip_pins = true;
        if (kS1Pins == kPinsKnown) {
            uint32_t any_flag = 0;
#pragma unroll
            for (int p = 0; p < NP; ++p) any_flag |= fe[0][p] | fo[0][p] | fe[1][p] | fo[1][p] | cfe[p] | cfo[p];
            trip_pins = __any_sync(FULL, (any_flag & CLOTHGPU_FLAG_PIN) != 0u);
        }
#if defined(S1_NO_MATH)  // timing experiment only: what the memory system gives this access pattern without the sticks
        if (u.iterations == 12345)
#endif
        {
        {  // ---- H-even: the lane's own pairs, rows r and r+1
            V2* hh[2 * NP];
            V2* tt[2 * NP];
            uint32_t live = 0, pins = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int k = q * NP + p;
                    hh[k] = &E[q][p], tt[k] = &O[q][p];
                    live |= live_of(fe[q][p], fo[q][p], CLOTHGPU_FLAG_ALIVE_H) << k;
                    pins |= ((fe[q][p] & CLOTHGPU_FLAG_PIN) | ((fo[q][p] & CLOTHGPU_FLAG_PIN) << 1)) << (2 * k);
                }
            const uint32_t tore = relax_n<T, DRAG, 2 * NP, kS1Pins>(hh, tt, live, pins, u, dragging, touched, 0xffffffffu, trip_pins);
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int k = q * NP + p;
                    if (own_cols && own_r[q]) visits += (live >> k) & 1u, torn += (tore >> k) & 1u;
                    if (DRAG && (tore & (1u << k))) fe[q][p] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_H;
                }
        }
        {  // ---- H-odd: the odd column of a pair (head) and the even column of the next pair (tail): inside the lane for
           // ---- p < NP - 1, the even column of lane + 1 for the last pair
            V2 Tn[2];
            uint32_t fn[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                Tn[q] = shfl_v2(E[q][0], true);
                fn[q] = __shfl_down_sync(FULL, fe[q][0], 1);
            }
            V2* hh[2 * NP];
            V2* tt[2 * NP];
            uint32_t live = 0, pins = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q)
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int k = q * NP + p;
                    const bool cross = p == NP - 1;
                    const uint32_t ftail = cross ? fn[q] : fe[q][cross ? 0 : p + 1];
                    hh[k] = &O[q][p], tt[k] = cross ? &Tn[q] : &E[q][cross ? 0 : p + 1];
                    if (!cross || lane < 31) live |= live_of(fo[q][p], ftail, CLOTHGPU_FLAG_ALIVE_H) << k;
                    pins |= ((fo[q][p] & CLOTHGPU_FLAG_PIN) | ((ftail & CLOTHGPU_FLAG_PIN) << 1)) << (2 * k);
                }
            const uint32_t tore = relax_n<T, DRAG, 2 * NP, kS1Pins>(hh, tt, live, pins, u, dragging, touched, 0xffffffffu, trip_pins);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const int k = q * NP + p;
                    if (own_cols && own_r[q]) visits += (live >> k) & 1u, torn += (tore >> k) & 1u;
                    if (DRAG && (tore & (1u << k))) fo[q][p] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_H;
                }
                const V2 back = shfl_v2(Tn[q], false);  // the relaxed tail returns to its lane
                if (lane >= 1) E[q][0] = back;
            }
        }
        {  // ---- V-even: rows (r, r+1), all columns
            V2* hh[NC];
            V2* tt[NC];
            uint32_t live = 0, pins = 0;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                hh[2 * p] = &E[0][p], tt[2 * p] = &E[1][p], hh[2 * p + 1] = &O[0][p], tt[2 * p + 1] = &O[1][p];
                live |= (live_of(fe[0][p], fe[1][p], CLOTHGPU_FLAG_ALIVE_V) | (live_of(fo[0][p], fo[1][p], CLOTHGPU_FLAG_ALIVE_V) << 1)) << (2 * p);
                pins |= ((fe[0][p] & CLOTHGPU_FLAG_PIN) | ((fe[1][p] & CLOTHGPU_FLAG_PIN) << 1) | ((fo[0][p] & CLOTHGPU_FLAG_PIN) << 2) |
                         ((fo[1][p] & CLOTHGPU_FLAG_PIN) << 3)) << (4 * p);
            }
            const uint32_t tore = relax_n<T, DRAG, NC, kS1Pins>(hh, tt, live, pins, u, dragging, touched, 0xffffffffu, trip_pins);
            if (own_cols && own_r[0]) visits += __popc(live), torn += __popc(tore);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (DRAG && (tore & (1u << (2 * p)))) fe[0][p] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
                if (DRAG && (tore & (2u << (2 * p)))) fo[0][p] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
            }
        }
        {  // ---- V-odd: rows (r-1, r); r-1 is the carried row
            V2* hh[NC];
            V2* tt[NC];
            uint32_t live = 0, pins = 0;
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                hh[2 * p] = &cE[p], tt[2 * p] = &E[0][p], hh[2 * p + 1] = &cO[p], tt[2 * p + 1] = &O[0][p];
                live |= (live_of(cfe[p], fe[0][p], CLOTHGPU_FLAG_ALIVE_V) | (live_of(cfo[p], fo[0][p], CLOTHGPU_FLAG_ALIVE_V) << 1)) << (2 * p);
                pins |= ((cfe[p] & CLOTHGPU_FLAG_PIN) | ((fe[0][p] & CLOTHGPU_FLAG_PIN) << 1) | ((cfo[p] & CLOTHGPU_FLAG_PIN) << 2) |
                         ((fo[0][p] & CLOTHGPU_FLAG_PIN) << 3)) << (4 * p);
            }
            const uint32_t tore = relax_n<T, DRAG, NC, kS1Pins>(hh, tt, live, pins, u, dragging, touched, 0xffffffffu, trip_pins);
            if (own_cols && r - 1 >= ra && r - 1 < rb) visits += __popc(live), torn += __popc(tore);
#pragma unroll
            for (int p = 0; p < NP; ++p) {
                if (DRAG && (tore & (1u << (2 * p)))) cfe[p] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
                if (DRAG && (tore & (2u << (2 * p)))) cfo[p] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
            }
        }
        }
        store_xy(r - 1, cE, cO, cfe, cfo);  // both rows have now seen all four colours
        store_xy(r, E[0], O[0], fe[0], fo[0]);
#pragma unroll
        for (int p = 0; p < NP; ++p) cE[p] = E[1][p], cO[p] = O[1][p], cfe[p] = fe[1][p], cfo[p] = fo[1][p];
    }
    // the last carried row has no row below it when the cloth (strip) ends here
    store_xy(rs + ((re - rs + 1) & ~1) - 1, cE, cO, cfe, cfo);

    if (st_up || st_dn) {  // warp-uniform: my share of the neighbours' ghost rows is written
        __threadfence_system();
        __syncwarp();
        if (lane == 0) {
            if (st_up) ss.done(0, (unsigned)(min(rb, peer_up.row_hi) - max(ra, peer_up.row_lo)));
            if (st_dn) ss.done(1, (unsigned)(min(rb, peer_dn.row_hi) - max(ra, peer_dn.row_lo)));
        }
    }

    events.flush(ctr);
    for (int o = 16; o > 0; o >>= 1) {
        visits += __shfl_down_sync(FULL, visits, o);
        torn += __shfl_down_sync(FULL, torn, o);
    }
    if (lane == 0) {
        // one sweep: the sticks visited and not torn are the live sticks the next frame starts with
        if (visits != torn) atomicAdd(&ctr->live_last, (unsigned long long)(visits - torn));
        if (visits) atomicAdd(&ctr->relaxations, (unsigned long long)visits);
        if (torn) {
            atomicAdd(&ctr->torn_total, (unsigned long long)torn);
            atomicAdd(&ctr->torn_last, (unsigned long long)torn);
        }
    }
}

}  // namespace clothgpu_impl
