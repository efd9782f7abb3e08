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
// active.  TWO launches over TILES of kSegBlock consecutive particles of the owned rows in PADDED row-major order
// (padding columns hold zero flags and add nothing, so the order of the listed sticks is the unpadded one):
//   segment_plan_kernel    reads the flag bytes only (2 of the 66 bytes per particle the export moves): which sticks every
//                          tile lists, where they go inside the tile's part of the list, and — by the last block to
//                          finish — where every tile's part starts;
//   segment_export_kernel  streams the coordinates and writes the list: with the plan in hand no block waits for another.
// A cloth of at most kSegSelfPlanTiles tiles (the GUI's) takes ONE launch: every export block plans its own tile and counts the
// tiles before it.
constexpr unsigned int kSegSelfPlanTiles = 4;
#ifndef SEG_PLAN_BLOCKS_PER_SM
#define SEG_PLAN_BLOCKS_PER_SM 6  // (all blocks of the plan kernel resident at once: 40 registers x 256 threads)
#endif
constexpr int kSegThreads = 256, kSegPerThread = 16;
constexpr int kSegBlock = kSegThreads * kSegPerThread;

// The plan of one tile.  Thread t of the planning block owns particles [16 t, 16 t + 16) of the tile; bit i of its masks is
// particle 16 t + i.  Read back as 32-bit words, word k holds the 32 particles of CHUNK k — the ballots of the export's trips.
struct __align__(16) SegPlan {
    uint16_t mv[kSegThreads], mh[kSegThreads];  // the particle lists its vertical / horizontal stick
    uint16_t hv[kSegThreads], hh[kSegThreads];  // ... and that stick is highlighted (both ends are, cloth.go:127-128)
    uint16_t pre[kSegThreads];                  // sticks of the tile before particle 16 t
    uint32_t total, any_hl, pad[2];             // sticks of the tile; whether any of them is highlighted
};
static_assert(sizeof(SegPlan) % 16 == 0, "moved by bulk copies");

template <int kThreads>
__device__ __forceinline__ unsigned int block_exclusive_scan(unsigned int v, unsigned int* warp_sums, unsigned int& total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned int incl = v;
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned int up_v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += up_v;
    }
    if (lane == 31) warp_sums[w] = incl;
    __syncthreads();
    unsigned int before = 0, all = 0;
#pragma unroll
    for (int k = 0; k < kThreads / 32; ++k) {
        const unsigned int t = warp_sums[k];
        if (k < w) before += t;
        all += t;
    }
    total = all;
    return before + incl - v;
}

// The flag bytes of 16 consecutive particles and of the 16 above them (one uint4 each) decide byte-parallel which sticks
// are listed: tail t lists its vertical stick when t and the particle above are active and that particle's ALIVE_V bit is
// set, its horizontal stick likewise with the particle to the left.  The four answers per particle are gathered into
// 16-bit masks.  `fl` = the flag bytes of the owned rows, `off` = the first particle (padded row-major, a multiple of 16),
// `n` = owned rows * pitch.  Returns the sticks listed, bit 20 set when one of them is highlighted (sums of 256 of these
// keep both apart: a tile lists at most 8192 sticks).
constexpr unsigned int kSegCountMask = 0xfffffu;
struct SegFlags16 {
    uint4 own, up;
    uint32_t left;  // flag byte of the particle left of the first one (0 at the start of a row)
};
__device__ __forceinline__ SegFlags16 seg_load16(const uint8_t* __restrict__ fl, const Geom& g, unsigned int n, unsigned int off) {
    SegFlags16 f;
    f.own = f.up = make_uint4(0, 0, 0, 0), f.left = 0;
    if (off < n) {
        const unsigned int pitch = (unsigned int)g.pitch, ly = off / pitch, cx = off - ly * pitch;
        f.own = *reinterpret_cast<const uint4*>(fl + off);
        if (g.own_row0 + (int)ly > g.held_row0) f.up = *reinterpret_cast<const uint4*>(fl + off - pitch);  // a ghost row of a strip, or absent at the cloth's top
        if (cx > 0) f.left = fl[off - 1];
    }
    return f;
}
__device__ __forceinline__ unsigned int seg_masks16(const SegFlags16& fl16, uint32_t& mv, uint32_t& mh, uint32_t& hv, uint32_t& hh) {
    constexpr uint32_t k1 = 0x01010101u, kGather = 0x01020408u;  // (m & k1) * kGather >> 24: bit 0 of bytes 0..3 -> bits 0..3
    mv = mh = hv = hh = 0;
    uint32_t carry = fl16.left;  // flag byte of the particle left of word q
    const uint32_t wq[4] = {fl16.own.x, fl16.own.y, fl16.own.z, fl16.own.w}, up[4] = {fl16.up.x, fl16.up.y, fl16.up.z, fl16.up.w};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        const uint32_t f = wq[q], act = (f >> 1) & k1;
        const uint32_t lf = (f << 8) | (carry & 0xffu);  // each byte: the flags of its left neighbour
        const uint32_t v = act & (up[q] >> 4) & (up[q] >> 1) & k1, h = act & (lf >> 3) & (lf >> 1) & k1;
        mv |= ((v * kGather) >> 24) << (4 * q);
        mh |= ((h * kGather) >> 24) << (4 * q);
        hv |= (((v & (f >> 2) & (up[q] >> 2)) * kGather) >> 24) << (4 * q);
        hh |= (((h & (f >> 2) & (lf >> 2)) * kGather) >> 24) << (4 * q);
        carry = f >> 24;
    }
    return (unsigned int)(__popc(mv) + __popc(mh)) | ((hv | hh) ? (kSegCountMask + 1u) : 0u);
}

// The plans of the tiles, a few tiles per block (the next tile's flag bytes are in flight while this one's are counted).
// The last block to finish turns the tiles' totals (left in `starts`, which holds a multiple of 16 entries) into their
// starts: 4096 tiles per step, every thread 16 consecutive ones.
__global__ void __launch_bounds__(kSegThreads)
    segment_plan_kernel(const uint8_t* __restrict__ flags, Geom g, int cloth, SegPlan* __restrict__ plan, unsigned int* __restrict__ starts,
                        unsigned int n_tiles, unsigned int* done, unsigned long long* __restrict__ total_out) {
    __shared__ unsigned int warp_sums[kSegThreads / 32];
    __shared__ bool s_last;
    const unsigned int n = (unsigned int)g.own_rows * (unsigned int)g.pitch;
    const uint8_t* fl = flags + (long long)cloth * g.cloth_stride + (long long)(g.own_row0 - g.held_row0) * g.pitch;
    SegFlags16 next = seg_load16(fl, g, n, blockIdx.x * (unsigned int)kSegBlock + threadIdx.x * 16u);
    for (unsigned int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const SegFlags16 cur = next;
        if (tile + gridDim.x < n_tiles) next = seg_load16(fl, g, n, (tile + gridDim.x) * (unsigned int)kSegBlock + threadIdx.x * 16u);
        uint32_t mv, mh, hv, hh;
        const unsigned int v = seg_masks16(cur, mv, mh, hv, hh);
        unsigned int all;
        const unsigned int before = block_exclusive_scan<kSegThreads>(v, warp_sums, all);
        SegPlan& mine = plan[tile];
        mine.mv[threadIdx.x] = (uint16_t)mv, mine.mh[threadIdx.x] = (uint16_t)mh;
        mine.hv[threadIdx.x] = (uint16_t)hv, mine.hh[threadIdx.x] = (uint16_t)hh;
        mine.pre[threadIdx.x] = (uint16_t)(before & kSegCountMask);
        if (threadIdx.x == 0) {
            mine.total = all & kSegCountMask, mine.any_hl = all >> 20;
            starts[tile] = all & kSegCountMask;  // (turned into the tile's start below)
        }
        __syncthreads();  // warp_sums is reused by the next tile
    }
    if (threadIdx.x == 0) {
        __threadfence();
        s_last = atomicAdd(done, 1u) == gridDim.x - 1u;
    }
    __syncthreads();
    if (!s_last) return;
    // ---- the last block: exclusive prefix over the tiles' totals ---------------------------------------------------------------
    __threadfence();
    unsigned int carried = 0;
    for (unsigned int lo = 0; lo < n_tiles; lo += kSegThreads * 16u) {
        const unsigned int at = lo + threadIdx.x * 16u;
        uint4 t4[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            t4[j] = make_uint4(0, 0, 0, 0);
            if (at + 4u * j < n_tiles) t4[j] = __ldcg(reinterpret_cast<const uint4*>(starts + at) + j);
        }
        unsigned int t[16], sum = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) t[4 * j] = t4[j].x, t[4 * j + 1] = t4[j].y, t[4 * j + 2] = t4[j].z, t[4 * j + 3] = t4[j].w;
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            if (at + i >= n_tiles) t[i] = 0;  // (entries behind the last tile are not initialised)
            sum += t[i];
        }
        unsigned int chunk_total;
        unsigned int run = carried + block_exclusive_scan<kSegThreads>(sum, warp_sums, chunk_total);
#pragma unroll
        for (int i = 0; i < 16; ++i) {
            const unsigned int mine_i = t[i];
            t[i] = run, run += mine_i;
        }
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (at + 4u * j < n_tiles) reinterpret_cast<uint4*>(starts + at)[j] = make_uint4(t[4 * j], t[4 * j + 1], t[4 * j + 2], t[4 * j + 3]);
        carried += chunk_total;
        __syncthreads();  // warp_sums is reused by the next step
    }
    if (threadIdx.x == 0) *total_out = carried, *done = 0u;  // (the counter is ready for the next export)
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

// 32 contiguous, 32-byte aligned bytes in ONE store (sm_100: 256-bit global stores): a whole sector per lane, whether
// the list lives in HBM or — written directly by the kernel — in pinned host memory.
__device__ __forceinline__ void st_pair(float4* p, const float4& a, const float4& b) {
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x),
                 "f"(b.y), "f"(b.z), "f"(b.w)
                 : "memory");
}

// One block per tile.  The plan arrives by one bulk copy; every warp owns 512 consecutive particles (16 chunks) and a
// private ring of kSegStages buffers in shared memory, each holding x, y of kSegStageN consecutive particles and of the
// kSegStageN above them, which one lane hands to the TMA engine — four bulk copies per buffer, completing on the buffer's
// mbarrier — the moment the block starts: kSegStages * kSegStageN * 32 bytes in flight per warp at no register cost,
// refilled as soon as a buffer has been consumed.  Per chunk: its two masks and its coordinates from shared memory, the
// particle to the left by shuffle.  An intact piece of cloth — nearly every chunk — stores lane l's two sticks straight to
// run + 2l, run + 2l + 1; otherwise the sticks are ranked with the masks, staged in a warp-private buffer at their rank
// and copied out as ONE contiguous, coalesced run.  QUADS: outline quads (2 float4 per stick, addSegment) instead of
// endpoint pairs; IDX: the particle indices of every stick's ends as well.  Highlight bytes: all zero for nearly every
// tile and then written as such; otherwise assembled in shared memory and written as 128-bit stores shifted so that shared
// and global addresses agree mod 16.
#ifndef SEG_MIN_BLOCKS
#define SEG_MIN_BLOCKS 3
#endif
#ifndef SEG_STAGE_N
#define SEG_STAGE_N 64
#endif
#ifndef SEG_STAGES
#define SEG_STAGES 3  // (measured on 4096^2: 3 buffers x 3 blocks per SM 152 us, 2 x 4 156 us, 4 x 32 particles 160 us)
#endif
#ifndef SEG_FULL_FAST
#define SEG_FULL_FAST 1
#endif
constexpr int kSegStageN = SEG_STAGE_N, kSegStages = SEG_STAGES;
// dynamic shared memory of the export kernel: the warps' coordinate rings
template <typename T>
constexpr size_t seg_smem_bytes() {
    return (size_t)(kSegThreads / 32) * kSegStages * 4 * kSegStageN * sizeof(T);
}
template <typename T, bool QUADS, bool IDX>
__global__ void __launch_bounds__(kSegThreads, SEG_MIN_BLOCKS)
    segment_export_kernel(Planes<T> P, Geom g, int cloth, const SegPlan* __restrict__ plan, const unsigned int* __restrict__ starts,
                          float line_width, float4* __restrict__ out, uint8_t* __restrict__ hl, uint2* __restrict__ idx,
                          unsigned long long cap, unsigned int n_tiles, unsigned long long* __restrict__ total_out) {
    constexpr int kWarps = kSegThreads / 32;
    constexpr int kPerWarp = kSegBlock / kWarps, kFills = kPerWarp / kSegStageN, kTripsPerFill = kSegStageN / 32;
    static_assert(kPerWarp % kSegStageN == 0 && kSegStageN % 32 == 0, "whole chunks per buffer, 16-byte copies");
    __shared__ SegPlan s_plan;
    __shared__ __align__(8) unsigned long long s_bar_plan, s_bar[kWarps][kSegStages];  // one mbarrier per warp and ring buffer
    __shared__ float4 s_xy[kWarps][64];
    __shared__ uint2 s_idx[IDX ? kWarps : 1][64];
    __shared__ __align__(16) uint8_t s_hl[2 * kSegBlock + 16];
    extern __shared__ __align__(16) unsigned char seg_dyn[];
    const long long n = (long long)g.own_rows * g.pitch;
    const long long base = (long long)cloth * g.cloth_stride + (long long)(g.own_row0 - g.held_row0) * g.pitch;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned int bid = blockIdx.x;
    const long long tile_first = (long long)bid * kSegBlock;
    const int n_tile = (int)min((long long)kSegBlock, n - tile_first);  // particles of this tile
    const T* X = P.x + base + tile_first;
    const T* Y = P.y + base + tile_first;
    const int pitch = g.pitch;
    if (lane == 0) {
        if (w == 0) mbar_init(&s_bar_plan, 1);
        for (int d = 0; d < kSegStages; ++d) mbar_init(&s_bar[w][d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    // ---- the warp's coordinate ring ------------------------------------------------------------------------------------------
    T* const ring = reinterpret_cast<T*>(seg_dyn) + (size_t)w * kSegStages * 4 * kSegStageN;  // [stage][x | y | xu | yu][kSegStageN]
    const int warp_lo = w * kPerWarp;        // first particle of the warp in the tile
    const int warp_n = n_tile - warp_lo;     // particles of the tile from there on (<= 0: none for this warp)
    // the row above the first held row does not exist (no stick hangs from it either: those shared bytes stay unread)
    const int warp_no_up = g.own_row0 > g.held_row0 ? 0 : (int)max(0ll, min((long long)kPerWarp, (long long)pitch - (tile_first + warp_lo)));
    // fill number f of the warp = particles [warp_lo + f * kSegStageN, + kSegStageN) of the tile -> buffer f % kSegStages
    auto fill_count = [&](int f) { return max(0, min(kSegStageN, warp_n - f * kSegStageN)); };  // (a multiple of 16)
    auto issue_fill = [&](int f) {  // one lane
        const int cnt = fill_count(f);
        if (cnt <= 0) return;
        const int at = warp_lo + f * kSegStageN;
        const int no_up = max(0, min(cnt, warp_no_up - f * kSegStageN));
        T* const buf = ring + (size_t)(f % kSegStages) * 4 * kSegStageN;
        unsigned long long* const bar = &s_bar[w][f % kSegStages];
        mbar_arrive_expect_tx(bar, (unsigned)((2 * cnt + 2 * (cnt - no_up)) * (int)sizeof(T)));
        bulk_g2s(buf, X + at, (unsigned)(cnt * (int)sizeof(T)), bar);
        bulk_g2s(buf + kSegStageN, Y + at, (unsigned)(cnt * (int)sizeof(T)), bar);
        if (cnt > no_up) {
            bulk_g2s(buf + 2 * kSegStageN + no_up, X + at + no_up - pitch, (unsigned)((cnt - no_up) * (int)sizeof(T)), bar);
            bulk_g2s(buf + 3 * kSegStageN + no_up, Y + at + no_up - pitch, (unsigned)((cnt - no_up) * (int)sizeof(T)), bar);
        }
    };
    // the particle left of the warp's first one (head of lane 0's horizontal stick in its first chunk)
    float carry_x = 0.f, carry_y = 0.f;
    T carry_tx = T(0), carry_ty = T(0);  // (converted when first used: the load stays off the warp's critical path)
    if (lane == 0) {
        if (w == 0 && plan) {
            mbar_arrive_expect_tx(&s_bar_plan, (unsigned)sizeof(SegPlan));
            bulk_g2s(&s_plan, plan + bid, (unsigned)sizeof(SegPlan), &s_bar_plan);
        }
#pragma unroll
        for (int f = 0; f < kSegStages && f < kFills; ++f) issue_fill(f);
        if (warp_n > 0 && tile_first + warp_lo > 0) carry_tx = X[warp_lo - 1], carry_ty = Y[warp_lo - 1];
    }
    // (output positions fit 32 bits: the host side rejects cloths with 2^32 or more possible sticks)
    unsigned int run0 = 0;
    if (plan) {  // block-uniform
        run0 = __ldg(starts + bid);
        mbar_wait(&s_bar_plan, 0);
    } else {
        // a cloth of a few tiles (the GUI's): no plan launch — the block plans its own tile and counts the tiles before it
        const uint8_t* fl = P.fl + base;
        unsigned int* const warp_sums = reinterpret_cast<unsigned int*>(s_xy);
        for (unsigned int t = 0; t <= bid; ++t) {
            uint32_t mv, mh, hv, hh;
            unsigned int all;
            const unsigned int before = block_exclusive_scan<kSegThreads>(
                seg_masks16(seg_load16(fl, g, (unsigned int)n, t * (unsigned int)kSegBlock + threadIdx.x * 16u), mv, mh, hv, hh), warp_sums, all);
            if (t < bid) {
                run0 += all & kSegCountMask;
            } else {
                s_plan.mv[threadIdx.x] = (uint16_t)mv, s_plan.mh[threadIdx.x] = (uint16_t)mh;
                s_plan.hv[threadIdx.x] = (uint16_t)hv, s_plan.hh[threadIdx.x] = (uint16_t)hh;
                s_plan.pre[threadIdx.x] = (uint16_t)(before & kSegCountMask);
                if (threadIdx.x == 0) {
                    s_plan.total = all & kSegCountMask, s_plan.any_hl = all >> 20;
                    if (bid + 1u == n_tiles) *total_out = (unsigned long long)run0 + (all & kSegCountMask);
                }
            }
            __syncthreads();
        }
    }
    const unsigned int hl_shift = run0 & 15u;  // (hl itself is 16-byte aligned)
    const unsigned int cap32 = (unsigned int)min(cap, 0xffffffffull);
    const uint32_t below = (1u << lane) - 1u;
    const bool any_hl = s_plan.any_hl != 0u;  // block-uniform
    if (any_hl) {
        for (int c = threadIdx.x * 16; c < 2 * kSegBlock + 16; c += kSegThreads * 16) *reinterpret_cast<uint4*>(&s_hl[c]) = make_uint4(0, 0, 0, 0);
        __syncthreads();
    }
    const uint32_t* tv = reinterpret_cast<const uint32_t*>(s_plan.mv);
    const uint32_t* th = reinterpret_cast<const uint32_t*>(s_plan.mh);

    // ---- trips --------------------------------------------------------------------------------------------------------------
    auto do_trip = [&](const int k, const T* buf) {  // chunk k of the tile, its coordinates at buf[plane * kSegStageN + lane]
        const uint32_t bv = tv[k], bh = th[k];
        const unsigned int total = __popc(bv) + __popc(bh);
        const float x2 = (float)buf[lane], y2 = (float)buf[kSegStageN + lane];  // cloth.go:110-111: float32(c.p2.x)
        // head of the horizontal stick = the lane to the left (lane 0: the last lane of the chunk before)
        float xl = __shfl_up_sync(0xffffffffu, x2, 1), yl = __shfl_up_sync(0xffffffffu, y2, 1);
        if (lane == 0) xl = carry_x, yl = carry_y;
        carry_x = __shfl_sync(0xffffffffu, x2, 31), carry_y = __shfl_sync(0xffffffffu, y2, 31);
        if (total == 0) return;  // warp-uniform
        const unsigned int rel = s_plan.pre[2 * k];  // the chunk's place in the tile's part of the list
        const unsigned int run = run0 + rel, hl_at = hl_shift + rel;
        float xu = 0.f, yu = 0.f;
        if ((bv >> lane) & 1u) xu = (float)buf[2 * kSegStageN + lane], yu = (float)buf[3 * kSegStageN + lane];
        uint32_t bhv = 0, bhh = 0;
        if (any_hl) bhv = reinterpret_cast<const uint32_t*>(s_plan.hv)[k], bhh = reinterpret_cast<const uint32_t*>(s_plan.hh)[k];
        unsigned int tail = 0;  // index of the lane's particle in the (unpadded) cloth: only the index list needs it
        if (IDX) {
            const long long p = tile_first + k * 32 + lane;
            const int row = (int)(p / pitch);
            tail = (unsigned int)row * (unsigned int)g.nx + (unsigned int)(p - (long long)row * pitch);
        }
#if SEG_FULL_FAST
        if ((bv & bh) == 0xffffffffu && run + 64u <= cap32) {
            // every lane lists both of its sticks: lane l's pair goes to run + 2l, run + 2l + 1 — no ranking, no staging,
            // each lane stores 32 (64) contiguous bytes
            const unsigned int o = run + 2u * lane;
            const float4 sv = make_float4(xu, yu, x2, y2), sh = make_float4(xl, yl, x2, y2);
            if (bhv | bhh) s_hl[hl_at + 2u * lane] = (bhv >> lane) & 1u, s_hl[hl_at + 2u * lane + 1u] = (bhh >> lane) & 1u;
            if (QUADS) {
                float4 q0, q1, q2, q3;
                segment_quad(sv, line_width, q0, q1);
                segment_quad(sh, line_width, q2, q3);
                st_pair(out + 2ull * o, q0, q1), st_pair(out + 2ull * o + 2, q2, q3);
            } else {
                if (run & 1u) out[o] = sv, out[o + 1] = sh;  // (warp-uniform)
                else st_pair(out + o, sv, sh);
                if (IDX) idx[o] = make_uint2(tail - g.nx, tail), idx[o + 1] = make_uint2(tail - 1, tail);
            }
            return;
        }
#endif
        const bool v = (bv >> lane) & 1u, h = (bh >> lane) & 1u;
        unsigned int r = __popc(bv & below) + __popc(bh & below);  // lane l's sticks come after those of lanes < l
        if (v) {
            s_xy[w][r] = make_float4(xu, yu, x2, y2);
            if ((bhv >> lane) & 1u) s_hl[hl_at + r] = 1;
            if (IDX) s_idx[w][r] = make_uint2(tail - g.nx, tail);
            ++r;
        }
        if (h) {
            s_xy[w][r] = make_float4(xl, yl, x2, y2);
            if ((bhh >> lane) & 1u) s_hl[hl_at + r] = 1;
            if (IDX) s_idx[w][r] = make_uint2(tail - 1, tail);
        }
        __syncwarp();
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const unsigned int j = q * 32 + lane, o = run + j;
            if (j < total && o < cap32) {
                if (QUADS) {
                    float4 q0, q1;
                    segment_quad(s_xy[w][j], line_width, q0, q1);
                    st_pair(out + 2ull * o, q0, q1);
                } else {
                    out[o] = s_xy[w][j];
                    if (IDX) idx[o] = s_idx[w][j];
                }
            }
        }
        __syncwarp();  // the staging buffer is rewritten by the next trip
    };
#pragma unroll 1
    for (int f = 0; f < kFills; ++f) {
        if (fill_count(f) <= 0) break;  // warp-uniform
        const int d = f % kSegStages;
        mbar_wait(&s_bar[w][d], (unsigned)(f / kSegStages) & 1u);  // the buffer's coordinates have landed
        if (f == 0) carry_x = (float)carry_tx, carry_y = (float)carry_ty;
        const T* buf = ring + (size_t)d * 4 * kSegStageN;
#pragma unroll
        for (int t = 0; t < kTripsPerFill; ++t) do_trip((warp_lo + f * kSegStageN) / 32 + t, buf + 32 * t);
        __syncwarp();
        if (lane == 0 && f + kSegStages < kFills) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the reads above precede the refill
            issue_fill(f + kSegStages);
        }
    }
    // ---- highlight bytes of the tile's sticks: hl[run0, run0 + total) <- s_hl[hl_shift, hl_end), or zeros.  Whole 16-byte
    // groups as one store each; the ragged first and last group byte by byte, one thread per byte.
    if (any_hl) __syncthreads();
    const unsigned int hl_end = hl_shift + s_plan.total;
    const unsigned int first16 = run0 - hl_shift;  // 16-aligned output position of s_hl[0]
    const unsigned int body_lo = hl_shift ? 16u : 0u, body_hi = hl_end & ~15u;
    for (unsigned int c = body_lo + threadIdx.x * 16u; c < body_hi; c += kSegThreads * 16u) {
        const uint4 v16 = any_hl ? *reinterpret_cast<const uint4*>(&s_hl[c]) : make_uint4(0, 0, 0, 0);
        if (first16 + c + 16u <= cap32) *reinterpret_cast<uint4*>(hl + first16 + c) = v16;
        else
            for (unsigned int b = 0; b < 16u; ++b)
                if (first16 + c + b < cap32) hl[first16 + c + b] = any_hl ? s_hl[c + b] : (uint8_t)0;
    }
    if (threadIdx.x < 32) {
        // lanes 0..15: the group before body_lo; lanes 16..31: the group from body_hi on (the same group when all the
        // tile's bytes fall into one: then only the first half writes)
        const unsigned int c = lane < 16 ? (unsigned int)lane : max(body_hi, body_lo) + (unsigned int)lane - 16u;
        const bool mine = lane < 16 ? (body_lo != 0u) : (body_hi >= body_lo);
        if (mine && c >= hl_shift && c < hl_end && first16 + c < cap32) hl[first16 + c] = any_hl ? s_hl[c] : (uint8_t)0;
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
