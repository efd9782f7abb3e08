// cloth_kernels.cuh — streaming schedule (one launch per pass) and the auxiliary kernels.
//
//   init_grid_kernel        <- (*Cloth).Init                          physics/cloth.go:42-81
//   verlet_kernel           <- loop over particles                    physics/cloth.go:92-94
//   relax_pass_kernel<C>    <- loop over sticks, one colour at a time physics/cloth.go:96-100
//   count_kernel / pack_masks_kernel / segment_export_kernel: read-back helpers
//
// The streaming schedule moves 64 B per stick relaxation through L2/HBM (2 endpoints x 16 B, read and
// write) and is the correctness cross-check and the "unfused" companion figure of DESIGN.md; the
// fused schedule in cloth_fused.cuh is the fast path.
#pragma once
#include "cloth_types.cuh"

namespace clothgpu_impl {

constexpr int kStreamThreads = 256;

__device__ __forceinline__ unsigned long long block_sum_ull(unsigned long long v) {
    // warp shuffle reduction, then one shared-memory hop across warps
    __shared__ unsigned long long warp_part[32];
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0) warp_part[w] = v;
    __syncthreads();
    unsigned long long t = 0;
    if (w == 0) {
        t = lane < (int)((blockDim.x + 31) >> 5) ? warp_part[lane] : 0ull;
        for (int o = 16; o > 0; o >>= 1) t += __shfl_down_sync(0xffffffffu, t, o);
    }
    __syncthreads();
    return t;  // valid in thread 0
}

// (*Cloth).Init: rest grid, px = x, everything active, every existing stick alive, Init pins.
template <typename T>
__global__ void init_grid_kernel(Planes<T> P, Geom g, int ny, int spacing, int pos_x, int pos_y,
                                 int pin_rule) {
    const long long per_cloth = (long long)g.held_rows * g.pitch;
    const long long total = per_cloth * g.ensemble;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int e = (int)(i / per_cloth);
        const long long r = i - (long long)e * per_cloth;
        const int ly = (int)(r / g.pitch), cx = (int)(r - (long long)ly * g.pitch);
        const int gy = g.held_row0 + ly;
        const long long at = (long long)e * g.cloth_stride + r;
        T x = T(0), y = T(0);
        uint8_t fl = 0;
        if (cx < g.nx) {
            x = (T)(pos_x + cx * spacing);  // cloth.go:53-56: integer arithmetic, then float64()
            y = (T)(pos_y + gy * spacing);
            fl = CLOTHGPU_FLAG_ACTIVE;
            if (cx + 1 < g.nx) fl |= CLOTHGPU_FLAG_ALIVE_H;  // cloth.go:66-70
            if (gy + 1 < ny) fl |= CLOTHGPU_FLAG_ALIVE_V;    // cloth.go:61-65
            if (gy == 0) {                                   // cloth.go:72-75
                if (pin_rule == CLOTHGPU_PIN_REFERENCE) {
                    if (cx % ((g.nx - 1) / 10) == 0) fl |= CLOTHGPU_FLAG_PIN;
                } else if (pin_rule == CLOTHGPU_PIN_TOP_ROW) {
                    fl |= CLOTHGPU_FLAG_PIN;
                }
            }
        }
        P.x[at] = x;
        P.y[at] = y;
        P.px[at] = x;
        P.py[at] = y;
        P.fl[at] = fl;
    }
}

// Loop over particles (physics/cloth.go:92-94), in place.  One thread per particle of the held rows.
// `per_cloth` (optional) holds one FrameU per cloth of an ensemble; otherwise every cloth uses u0.
// (PER_CLOTH is a template parameter, not a run-time test: nvcc 12.9 miscompiles the run-time form
//  `u = u0; if (per_cloth) u = per_cloth[e];` — it keeps reading the first field from the kernel
//  parameter bank.  tests/test_gpu_parity.py::test_ensemble_cloths_are_independent guards this.)
template <typename T, bool PER_CLOTH>
__device__ __forceinline__ FrameU<T> pick_frame(const FrameU<T>& u0, const FrameU<T>* __restrict__ per_cloth, int e) {
    if (PER_CLOTH) return per_cloth[e];
    return u0;
}

template <typename T, bool PER_CLOTH>
__global__ void __launch_bounds__(kStreamThreads)
    verlet_kernel(Planes<T> P, Geom g, const FrameU<T> u0, const FrameU<T>* __restrict__ per_cloth, DevCounters* ctr) {
    const int e = blockIdx.y;
    const FrameU<T> u = pick_frame<T, PER_CLOTH>(u0, per_cloth, e);
    const long long n = (long long)g.held_rows * g.pitch;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    prepare_next_frame_counters(ctr);
    ParticleEvents events;
    if (i < n && (int)(i % g.pitch) < g.nx) {
        const long long at = (long long)e * g.cloth_stride + i;
        T x = P.x[at], y = P.y[at], px = P.px[at], py = P.py[at];
        uint32_t fl = P.fl[at];
        const uint32_t fl_in = fl;
        particle_update(x, y, px, py, fl, u);
        P.x[at] = x;
        P.y[at] = y;
        P.px[at] = px;
        P.py[at] = py;
        P.fl[at] = (uint8_t)fl;
        const int gy = g.held_row0 + (int)(i / g.pitch);
        if (gy >= g.own_row0 && gy < g.own_row0 + g.own_rows) events.note(fl_in, fl);  // ghost rows belong to the neighbour
    }
    events.flush(ctr);
}

// One colour pass over the held rows, in place.  COLOUR: 0 H-even, 1 H-odd, 2 V-even, 3 V-odd
// (parity of the head particle's global column / row).  Each colour is a perfect matching, so the
// threads of one pass never touch the same particle.
template <typename T, int COLOUR, bool PER_CLOTH>
__global__ void __launch_bounds__(kStreamThreads)
    relax_pass_kernel(Planes<T> P, Geom g, const FrameU<T> u0, const FrameU<T>* __restrict__ per_cloth,
                      DevCounters* ctr, int last_sweep) {
    const int e = blockIdx.y;
    const FrameU<T> u = pick_frame<T, PER_CLOTH>(u0, per_cloth, e);
    constexpr bool kVertical = COLOUR >= 2;
    constexpr int kOdd = COLOUR & 1;
    unsigned long long visits = 0, torn = 0;
    long long head = -1, tail = -1;
    bool owned = false;
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (!kVertical) {
        const int per_row = max(1, g.nx / 2);  // upper bound on sticks of one colour per row
        const int ly = (int)(i / per_row), k = (int)(i % per_row);
        const int c = 2 * k + kOdd;
        if (ly < g.held_rows && c + 1 < g.nx) {
            head = (long long)ly * g.pitch + c;
            tail = head + 1;
            const int gy = g.held_row0 + ly;
            owned = gy >= g.own_row0 && gy < g.own_row0 + g.own_rows;
        }
    } else {
        // first held row with the wanted global parity
        const int first = ((g.held_row0 & 1) == kOdd) ? 0 : 1;
        const int j = (int)(i / g.nx), c = (int)(i % g.nx);
        const int ly = first + 2 * j;
        if (ly + 1 < g.held_rows) {
            head = (long long)ly * g.pitch + c;
            tail = head + g.pitch;
            const int gy = g.held_row0 + ly;
            owned = gy >= g.own_row0 && gy < g.own_row0 + g.own_rows;
        }
    }
    if (head >= 0) {
        const long long base = (long long)e * g.cloth_stride;
        head += base;
        tail += base;
        const uint32_t f1 = P.fl[head], f2 = P.fl[tail];
        const uint32_t alive_bit = kVertical ? CLOTHGPU_FLAG_ALIVE_V : CLOTHGPU_FLAG_ALIVE_H;
        if ((f1 & alive_bit) && (f1 & f2 & CLOTHGPU_FLAG_ACTIVE)) {  // cloth.go:97
            T x1 = P.x[head], y1 = P.y[head], x2 = P.x[tail], y2 = P.y[tail];
            const bool tore = stick_relax(x1, y1, x2, y2, (f1 & CLOTHGPU_FLAG_PIN) != 0,
                                          (f2 & CLOTHGPU_FLAG_PIN) != 0, u);
            P.x[head] = x1;
            P.y[head] = y1;
            P.x[tail] = x2;
            P.y[tail] = y2;
            if (tore) P.fl[head] = (uint8_t)(f1 & ~alive_bit);  // tear = clear the alive bit in place
            if (owned) {
                visits = 1;
                torn = tore ? 1 : 0;
            }
        }
    }
    const unsigned long long v = block_sum_ull(visits);
    const unsigned long long t = block_sum_ull(torn);
    if (threadIdx.x == 0) {
        // the sticks the frame's last sweep visited and did not tear are the live sticks the next frame starts with
        if (last_sweep && v != t) atomicAdd(&ctr->live_last, v - t);
        if (v) atomicAdd(&ctr->relaxations, v);
        if (t) {
            atomicAdd(&ctr->torn_total, t);
            atomicAdd(&ctr->torn_last, t);
        }
    }
}

// Census of the owned rows: {live sticks, alive sticks, pinned, inactive, highlighted}.
// 16 flag bytes per thread and trip (128-bit loads), counted four bytes at a time with byte-parallel masks; padding
// columns and the spare row hold zero flags, so they add nothing (inactive = existing - active is fixed up by the host
// side of the sum: the kernel returns the ACTIVE count in slot 3).
template <typename T>
__global__ void __launch_bounds__(kStreamThreads)
    count_kernel(Planes<T> P, Geom g, unsigned long long* out5) {
    const int e = blockIdx.y;
    const long long chunks = (long long)g.own_rows * g.pitch / 16;  // pitch is a multiple of 16
    const uint8_t* base = P.fl + (long long)e * g.cloth_stride + (long long)(g.own_row0 - g.held_row0) * g.pitch;
    unsigned int live = 0, alive = 0, pinned = 0, active = 0, hl = 0;
    constexpr uint32_t k1 = 0x01010101u;
    for (long long c = blockIdx.x * (long long)blockDim.x + threadIdx.x; c < chunks; c += (long long)gridDim.x * blockDim.x) {
        const uint4 w4 = *reinterpret_cast<const uint4*>(base + c * 16);
        const uint4 b4 = *reinterpret_cast<const uint4*>(base + c * 16 + g.pitch);  // the row below (ghost / spare row at the end)
        const uint32_t nxt = base[c * 16 + 16];                                     // right neighbour of the chunk's last byte
        const uint32_t w[5] = {w4.x, w4.y, w4.z, w4.w, nxt};
        const uint32_t bl[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t f = w[q];
            const uint32_t act = (f >> 1) & k1, ah = (f >> 3) & k1, av = (f >> 4) & k1;
            const uint32_t act_right = (act >> 8) | ((w[q + 1] >> 1) << 24 & 0x01000000u);
            const uint32_t act_below = (bl[q] >> 1) & k1;
            pinned += __popc(f & k1);
            active += __popc(act);
            hl += __popc((f >> 2) & k1);
            alive += __popc(ah) + __popc(av);
            live += __popc(ah & act & act_right) + __popc(av & act & act_below);  // cloth.go:97
        }
    }
    unsigned long long v[5] = {live, alive, pinned, active, hl};
    for (int k = 0; k < 5; k++) {
        const unsigned long long s = block_sum_ull(v[k]);
        if (threadIdx.x == 0 && s) atomicAdd(&out5[k], s);
    }
}

// Flag byte -> five 1-bit planes over the owned particles in row-major order WITHOUT row padding
// (bit i = particle i = x + y*nx), assembled with warp ballots.
template <typename T>
__global__ void __launch_bounds__(kStreamThreads)
    pack_masks_kernel(Planes<T> P, Geom g, int cloth, uint32_t* __restrict__ out, long long words) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;  // particle index
    const long long n = (long long)g.own_rows * g.nx;
    uint32_t f = 0;
    if (i < n) {
        const int ly = (int)(i / g.nx), cx = (int)(i % g.nx);
        f = P.fl[(long long)cloth * g.cloth_stride +
                 (long long)(g.own_row0 - g.held_row0 + ly) * g.pitch + cx];
    }
    const long long w = i >> 5;
#pragma unroll
    for (int b = 0; b < 5; b++) {
        const uint32_t m = __ballot_sync(0xffffffffu, (f >> b) & 1u);
        if ((threadIdx.x & 31) == 0 && w < words) out[(long long)b * words + w] = m;
    }
}

// ---- live-stick list (render export; physics/cloth.go:108-114, 126-134) ------------------------------
// Sticks are listed in the order Cloth.Init created them (physics/cloth.go:61-70): by TAIL particle in
// row-major order, the vertical stick (from the particle above) before the horizontal one (from the
// particle to the left).  A stick is listed when it is still in the slice and both endpoints are
// active.  ONE launch: every block of kSegBlock consecutive particles counts its sticks from the flag bytes, obtains
// its place in the list by a decoupled look-back over the blocks before it, and scatters.
constexpr int kSegPerThread = 16;
constexpr int kSegBlock = kStreamThreads * kSegPerThread;

// Decoupled look-back status of one block: epoch (29 bits) | flag (2 bits) | value (33 bits).  The epoch is the launch
// number, so the array is never cleared: a word of an earlier launch simply reads as "nothing yet".
constexpr unsigned long long kSegFlagAggregate = 1ull, kSegFlagPrefix = 2ull;
__device__ __forceinline__ unsigned long long seg_status(unsigned int epoch, unsigned long long flag, unsigned long long value) {
    return ((unsigned long long)(epoch & 0x1fffffffu) << 35) | (flag << 33) | (value & 0x1ffffffffull);
}

// addSegment / normal of the reference's render loops (physics/cloth.go:150-168) for one stick a -> b: the four float32
// corners {a+n, b+n, b-n, a-n} of its outline quad, n = the stick's normal scaled to the line width.  math.Hypot is
// max * sqrt(1 + (min/max)^2) in binary64 (hypot_amd64.s); everything else is float32 arithmetic, one rounding per
// operation (-fmad=false, IEEE division).
__device__ __forceinline__ void segment_quad(const float4 s, float line_width, float4& q0, float4& q1) {
    const float dx = s.z - s.x, dy = s.w - s.y;  // dir := b.Sub(a)
    const float rx = +dy, ry = -dx;              // dir.X, dir.Y = +dir.Y, -dir.X
    double p = fabs((double)rx), q = fabs((double)ry);
    if (p < q) {
        const double t = p;
        p = q, q = t;
    }
    double d = 0.0;
    if (p != 0.0) {
        q = q / p;
        d = p * sqrt(1.0 + q * q);
    }
    float nx = 0.f, ny = 0.f;  // f32.Point{} when |d| < 1e-5
    if (!(fabs(d) < 1e-5)) {
        const float sc = line_width / (float)d;
        nx = rx * sc, ny = ry * sc;
    }
    q0 = make_float4(s.x + nx, s.y + ny, s.z + nx, s.w + ny);
    q1 = make_float4(s.z - nx, s.w - ny, s.x - nx, s.y - ny);
}

// One block per kSegBlock particles of the owned rows in PADDED row-major order (padding columns hold zero flags and add
// nothing, so the order of the listed sticks is the unpadded one); blocks take their index from a ticket counter, so a
// block's predecessors are always running or done.  Every warp owns 512 consecutive particles.  It loads their flag
// bytes and those of the row above once (16 bytes per lane), keeps them in shared memory and counts the sticks with
// byte-parallel tests: tail t lists its vertical stick when t and the particle above are active and that particle's
// ALIVE_V bit is set, its horizontal stick likewise with the particle to the left.  Warp 0 publishes the block's count,
// sums the counts of the blocks before it (decoupled look-back, 32 blocks per step) and publishes the inclusive prefix;
// from there on the warps run independently: 32 particles per trip, flags from shared memory, the coordinate loads of
// the next trip in flight while this trip's sticks are ranked with two ballots, converted and staged in a warp-private
// buffer at their rank, then copied out as ONE contiguous, fully coalesced run (before ranking the sticks of consecutive
// lanes are not adjacent in the output: a lane has 0, 1 or 2 of them).  QUADS: the run is written as outline quads
// (2 float4 per stick, addSegment) instead of endpoint pairs.
#ifndef SEG_MIN_BLOCKS
#define SEG_MIN_BLOCKS 4
#endif
#ifndef SEG_PREFETCH2
#define SEG_PREFETCH2 0
#endif
#ifndef SEG_FULL_FAST
#define SEG_FULL_FAST 1
#endif
template <typename T, bool QUADS>
__global__ void __launch_bounds__(kStreamThreads, SEG_MIN_BLOCKS)
    segment_export_kernel(Planes<T> P, Geom g, int cloth, unsigned long long* __restrict__ status, unsigned long long* ticket,
                          unsigned long long ticket_base, unsigned int epoch, unsigned int n_blocks, float line_width,
                          float4* __restrict__ out, uint8_t* __restrict__ hl, uint2* __restrict__ idx,
                          unsigned long long cap, unsigned long long* __restrict__ total_out) {
    constexpr int kWarps = kStreamThreads / 32, kPerWarp = kSegBlock / kWarps, kTrips = kPerWarp / 32;
    static_assert(kPerWarp == 32 * 16, "one uint4 of flag bytes per lane");
    __shared__ unsigned int warp_tot[kWarps];
    __shared__ unsigned int s_bid;
    __shared__ unsigned long long s_excl;
    __shared__ __align__(16) uint8_t s_f[kWarps][kPerWarp + 16];   // [15] = the particle left of the warp's first one
    __shared__ __align__(16) uint8_t s_fu[kWarps][kPerWarp];
    __shared__ float4 s_xy[kWarps][64];
    __shared__ uint2 s_idx[kWarps][64];
    // the highlight bytes of all sticks of the warp (at most 1024), shifted so that shared and global addresses agree
    // mod 16: they leave as 128-bit stores at the end instead of two one-byte stores per lane and trip
    __shared__ __align__(16) uint8_t s_hl[kWarps][2 * kPerWarp + 16];
    const long long n = (long long)g.own_rows * g.pitch;
    const long long base = (long long)cloth * g.cloth_stride + (long long)(g.own_row0 - g.held_row0) * g.pitch;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const uint32_t below = (1u << lane) - 1u;
    if (threadIdx.x == 0) s_bid = (unsigned int)(atomicAdd(ticket, 1ull) - ticket_base);
    __syncthreads();
    const unsigned int bid = s_bid;
    const long long warp_first = (long long)bid * kSegBlock + (long long)w * kPerWarp;
    const uint8_t* fl = P.fl + base;
    const T* X = P.x + base;
    const T* Y = P.y + base;

    {   // ---- flags of this warp's 512 particles -> shared memory; their sticks ---------------------------------------------
        constexpr uint32_t k1 = 0x01010101u;
        const long long off = warp_first + lane * 16;
        unsigned int cnt = 0;
        uint4 w4 = make_uint4(0, 0, 0, 0), u4 = make_uint4(0, 0, 0, 0);
        uint32_t left = 0;
        if (off < n) {
            const int ly = (int)(off / g.pitch), cx = (int)(off - (long long)ly * g.pitch);
            w4 = *reinterpret_cast<const uint4*>(fl + off);
            if (g.own_row0 + ly > g.held_row0) u4 = *reinterpret_cast<const uint4*>(fl + off - g.pitch);  // a ghost row of a strip, or absent at the cloth's top
            if (cx > 0) left = fl[off - 1];
            const uint32_t wq[4] = {w4.x, w4.y, w4.z, w4.w}, up[4] = {u4.x, u4.y, u4.z, u4.w};
            uint32_t carry = left;  // flag byte of the particle left of word q
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t f = wq[q], act = (f >> 1) & k1;
                const uint32_t fl_left = (f << 8) | (carry & 0xffu);  // each byte: the flags of its left neighbour
                cnt += __popc(act & (up[q] >> 4) & (up[q] >> 1) & k1) + __popc(act & (fl_left >> 3) & (fl_left >> 1) & k1);
                carry = f >> 24;
            }
        }
        *reinterpret_cast<uint4*>(&s_f[w][16 + lane * 16]) = w4;
        *reinterpret_cast<uint4*>(&s_fu[w][lane * 16]) = u4;
        if (lane == 0) s_f[w][15] = (uint8_t)left;  // (0 when the first particle starts a row: no stick to its left)
        for (int o = 16; o > 0; o >>= 1) cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        if (lane == 0) warp_tot[w] = cnt;
    }
    __syncthreads();
    // rows and columns in 32-bit arithmetic: one division per lane, then every trip advances the lane by 32 particles
    const int row0 = (int)(warp_first / g.pitch), cx0 = (int)(warp_first - (long long)row0 * g.pitch);
    const int n_here = (int)min((long long)kPerWarp, n - warp_first);  // particles of this warp (<= 0: none)
    X += warp_first, Y += warp_first;
    const unsigned int pitch = (unsigned int)g.pitch;
    unsigned int f_cx = (unsigned int)(cx0 + lane) % pitch;                     // column / tail index of the lane's particle
    unsigned int f_tail = (unsigned int)(row0 + (cx0 + lane) / (int)pitch) * (unsigned int)g.nx + f_cx;  // in the NEXT fetch

    constexpr uint32_t kV = 1u << 24, kH = 1u << 25;
    struct Trip {
        T x, y, xu, yu, xl0, yl0;  // own position, the particle above, the particle left of lane 0's
        uint32_t fl3;              // flag bytes: own | above << 8 | left << 16; kV / kH = the lane lists that stick
        unsigned int tail;         // index of the particle in the (unpadded) cloth
    };
    auto fetch = [&](int trip) -> Trip {
        Trip t;
        t.x = t.y = t.xu = t.yu = t.xl0 = t.yl0 = T(0);
        t.fl3 = 0u, t.tail = f_tail;
        const int li = trip * 32 + lane;
        if (li < n_here) {
            const uint32_t f = s_f[w][16 + li];
            t.fl3 = f;
            if (f & CLOTHGPU_FLAG_ACTIVE) {
                const uint32_t fu = s_fu[w][li], fw = f_cx > 0 ? s_f[w][15 + li] : 0u;
                t.fl3 |= (fu << 8) | (fw << 16);
                if ((fu & (CLOTHGPU_FLAG_ALIVE_V | CLOTHGPU_FLAG_ACTIVE)) == (CLOTHGPU_FLAG_ALIVE_V | CLOTHGPU_FLAG_ACTIVE)) t.fl3 |= kV;
                if ((fw & (CLOTHGPU_FLAG_ALIVE_H | CLOTHGPU_FLAG_ACTIVE)) == (CLOTHGPU_FLAG_ALIVE_H | CLOTHGPU_FLAG_ACTIVE)) t.fl3 |= kH;
            }
            // every lane loads its own position (its right neighbour may need it as the head of a horizontal stick)
            t.x = X[li], t.y = Y[li];
            if (t.fl3 & kV) t.xu = X[li - (int)pitch], t.yu = Y[li - (int)pitch];
            if (lane == 0 && (t.fl3 & kH)) t.xl0 = X[li - 1], t.yl0 = Y[li - 1];
        }
        // the lane's particle of the next trip: 32 further on, rows are `pitch` apart and hold nx particles
        f_cx += 32, f_tail += 32;
        while (f_cx >= pitch) f_cx -= pitch, f_tail += (unsigned int)g.nx - pitch;
        return t;
    };

    Trip cur = fetch(0);  // in flight during the look-back below
#if SEG_PREFETCH2
    Trip ahead = cur;     // the trip after the next one: two trips of loads in flight
    if (1 < kTrips) ahead = fetch(1);
#endif
    if (w == 0) {   // ---- decoupled look-back: where this block's sticks start in the list ------------------------------------
        unsigned long long block_total = 0;
        for (int k = 0; k < kWarps; ++k) block_total += warp_tot[k];
        volatile unsigned long long* st = status;
        if (lane == 0) {
            st[bid] = seg_status(epoch, bid == 0 ? kSegFlagPrefix : kSegFlagAggregate, block_total);
            __threadfence();
        }
        unsigned long long excl = 0;
        long long look = (long long)bid - 1;  // nearest predecessor this step looks at (lane 0)
        while (look >= 0) {
            const long long at = look - lane;
            unsigned long long word = seg_status(epoch, kSegFlagPrefix, 0);  // before block 0: an empty prefix
            if (at >= 0) word = st[at];
            const bool valid = (unsigned int)(word >> 35) == (epoch & 0x1fffffffu) && ((word >> 33) & 3ull) != 0ull;
            const bool is_prefix = valid && ((word >> 33) & 3ull) == kSegFlagPrefix;
            const uint32_t bp = __ballot_sync(0xffffffffu, is_prefix), binv = __ballot_sync(0xffffffffu, !valid);
            const int first_p = bp ? __ffs(bp) - 1 : 32;
            const uint32_t need = first_p >= 31 ? 0xffffffffu : ((2u << first_p) - 1u);
            if (binv & need) {  // a block this step depends on has not published yet
                __nanosleep(32);
                continue;
            }
            unsigned long long v = (lane <= first_p) ? (word & 0x1ffffffffull) : 0ull;
            for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
            excl += __shfl_sync(0xffffffffu, v, 0);
            if (first_p < 32) break;
            look -= 32;
        }
        if (lane == 0) {
            if (bid != 0) {
                st[bid] = seg_status(epoch, kSegFlagPrefix, excl + block_total);
                __threadfence();
            }
            s_excl = excl;
            if (bid + 1 == n_blocks) *total_out = excl + block_total;
        }
    }
    __syncthreads();
    // (output positions fit 32 bits: the host side rejects cloths with 2^32 or more possible sticks)
    unsigned int run = (unsigned int)s_excl;
    for (int k = 0; k < w; ++k) run += warp_tot[k];
    const unsigned int cap32 = (unsigned int)min(cap, 0xffffffffull);
    const unsigned int run0 = run, hl_shift = run0 & 15u;  // (hl itself is 16-byte aligned)
    unsigned int hl_at = hl_shift;                         // where this trip's highlight bytes go in s_hl[w]

#pragma unroll 1
    for (int trip = 0; trip < kTrips; ++trip) {
        if (trip * 32 >= n_here) break;  // warp-uniform
#if SEG_PREFETCH2
        Trip nxt = ahead;
        if (trip + 2 < kTrips) ahead = fetch(trip + 2);
#else
        Trip nxt = cur;
        if (trip + 1 < kTrips) nxt = fetch(trip + 1);  // its loads are in flight while this trip is written out
#endif
        const bool v = (cur.fl3 & kV) != 0, h = (cur.fl3 & kH) != 0;
        const uint32_t bv = __ballot_sync(0xffffffffu, v), bh = __ballot_sync(0xffffffffu, h);
        const unsigned int total = __popc(bv) + __popc(bh);
        if (total) {  // warp-uniform
            const float x2 = (float)cur.x, y2 = (float)cur.y;  // cloth.go:110-111: float32(c.p2.x)
            // head of the horizontal stick = the lane to the left (lane 0: loaded separately)
            float xl = __shfl_up_sync(0xffffffffu, x2, 1), yl = __shfl_up_sync(0xffffffffu, y2, 1);
            if (lane == 0) xl = (float)cur.xl0, yl = (float)cur.yl0;
            const uint32_t hl_own = cur.fl3 & CLOTHGPU_FLAG_HIGHLIGHT;
#if SEG_FULL_FAST
            if ((bv & bh) == 0xffffffffu && run + 64u <= cap32) {
                // every lane lists both of its sticks (an intact piece of cloth: nearly every trip): lane l's pair goes to
                // run + 2l, run + 2l + 1 — no ranking, no staging, each lane stores 32 (64) contiguous bytes
                const unsigned int o = run + 2u * lane;
                const float4 sv = make_float4((float)cur.xu, (float)cur.yu, x2, y2), sh = make_float4(xl, yl, x2, y2);
                s_hl[w][hl_at + 2u * lane] = (hl_own & (cur.fl3 >> 8)) != 0;
                s_hl[w][hl_at + 2u * lane + 1u] = (hl_own & (cur.fl3 >> 16)) != 0;
                if (QUADS) {
                    float4 q0, q1, q2, q3;
                    segment_quad(sv, line_width, q0, q1);
                    segment_quad(sh, line_width, q2, q3);
                    out[2ull * o] = q0, out[2ull * o + 1] = q1, out[2ull * o + 2] = q2, out[2ull * o + 3] = q3;
                } else {
                    out[o] = sv, out[o + 1] = sh;
                    if (idx) idx[o] = make_uint2(cur.tail - g.nx, cur.tail), idx[o + 1] = make_uint2(cur.tail - 1, cur.tail);
                }
                run += 64u, hl_at += 64u;
                cur = nxt;
                continue;
            }
#endif
            unsigned int r = __popc(bv & below) + __popc(bh & below);  // lane l's sticks come after those of lanes < l
            if (v) {
                s_xy[w][r] = make_float4((float)cur.xu, (float)cur.yu, x2, y2);
                s_hl[w][hl_at + r] = (hl_own & (cur.fl3 >> 8)) != 0;  // cloth.go:127-128: both ends highlighted
                if (!QUADS) s_idx[w][r] = make_uint2(cur.tail - g.nx, cur.tail);
                ++r;
            }
            if (h) {
                s_xy[w][r] = make_float4(xl, yl, x2, y2);
                s_hl[w][hl_at + r] = (hl_own & (cur.fl3 >> 16)) != 0;
                if (!QUADS) s_idx[w][r] = make_uint2(cur.tail - 1, cur.tail);
            }
            __syncwarp();
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                const unsigned int j = q * 32 + lane, o = run + j;
                if (j < total && o < cap32) {
                    if (QUADS) {
                        float4 q0, q1;
                        segment_quad(s_xy[w][j], line_width, q0, q1);
                        out[2ull * o] = q0, out[2ull * o + 1] = q1;
                    } else {
                        out[o] = s_xy[w][j];
                        if (idx) idx[o] = s_idx[w][j];
                    }
                }
            }
            __syncwarp();  // the staging buffer is rewritten by the next trip
            run += total, hl_at += total;
        }
        cur = nxt;
    }
    // ---- highlight bytes [run0, run) of this warp: whole 16-byte groups as one store, the ragged ends byte by byte ----
    __syncwarp();
    const unsigned int first16 = run0 - hl_shift;  // 16-aligned output position of s_hl[w][0]
    for (unsigned int c = lane * 16u; c < hl_at; c += 32u * 16u) {
        const unsigned int o = first16 + c;
        if (c >= hl_shift && c + 16u <= hl_at && o + 16u <= cap32) {
            *reinterpret_cast<uint4*>(hl + o) = *reinterpret_cast<const uint4*>(&s_hl[w][c]);
        } else {
            for (unsigned int b = 0; b < 16u; ++b)
                if (c + b >= hl_shift && c + b < hl_at && o + b < cap32) hl[o + b] = s_hl[w][c + b];
        }
    }
}

// ---- row strips over several GPUs: the explicit edge-row push and the whole-operation handshake ---------------------
// The frame kernels hand-shake by themselves, piece by piece (cloth_types.cuh, StripSync).  The two host-driven strip
// operations — clothgpu_strip_push after a clothgpu_write_state, and clothgpu_reset — are rare and keep the simple
// stream-ordered form: a wait kernel (both neighbours have finished operation k-1), the operation, a signal kernel
// that publishes k+1 in both neighbours' slots.
__global__ void strip_wait_kernel(const unsigned int* my_slots, int wait_up, int wait_dn, unsigned int need) {
    // epochs only grow; the signed difference tolerates wrap-around
    if (wait_up)
        while ((int)(ld_acquire_sys(my_slots + 0) - need) < 0) __nanosleep(64);
    if (wait_dn)
        while ((int)(ld_acquire_sys(my_slots + 1) - need) < 0) __nanosleep(64);
}

__global__ void strip_signal_kernel(unsigned int* up_slot, unsigned int* dn_slot, unsigned int epoch) {
    __threadfence_system();  // everything this stream wrote before (incl. peer stores of the frame kernel)
    if (up_slot) st_release_sys(up_slot, epoch);
    if (dn_slot) st_release_sys(dn_slot, epoch);
}

// Copy my owned rows that a neighbour keeps as ghosts into its memory (all five planes).  Used after the state was
// written from the host (clothgpu_write_state) — per-frame traffic goes through the fused kernel itself.
template <typename T>
__global__ void __launch_bounds__(kStreamThreads)
    strip_push_kernel(Planes<T> mine, Geom g, PeerOut<T> peer) {
    const int e = blockIdx.y;
    const int rows = peer.row_hi - peer.row_lo;
    const long long n = (long long)rows * g.pitch;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int r = (int)(i / g.pitch), cx = (int)(i - (long long)r * g.pitch);
        const int gy = peer.row_lo + r;
        const long long at = (long long)e * g.cloth_stride + (long long)(gy - g.held_row0) * g.pitch + cx;
        const long long pat = (long long)e * peer.cloth_stride + (long long)(gy - peer.held_row0) * g.pitch + cx;
        peer.x[pat] = mine.x[at];
        peer.y[pat] = mine.y[at];
        peer.px[pat] = mine.px[at];
        peer.py[pat] = mine.py[at];
        peer.fl[pat] = mine.fl[at];
    }
}

}  // namespace clothgpu_impl
