// cloth_wave.cuh — the wavefront schedule for frames with several sweeps: ONE launch per frame, rows streamed through
// a ring of shared memory, the K sweeps pipelined along the row direction.
//
// The tile kernel (cloth_fused.cuh) loads a 128 x 100 tile, runs K sweeps on it and stores it: its arithmetic is 1.5x
// redundant for K = 8 (a 16-wide stale rim on all four sides) and its load / store phases do not overlap the sweeps.
// Here a CTA owns a band of 256 columns and a chunk of rows, and streams the rows downwards:
//
//     entry    rows (2q, 2q+1) arrive (cp.async staging ring, three pairs ahead), particle loop (cloth.go:92-94),
//              H-even of sweep 0, the state words of their 2x2 blocks, then they join the ring of 4K+4 rows; the pair
//              that has passed all K sweeps leaves for HBM from the ring slots the new pair is about to take
//     stage j  (sweep j, j = 0..K-1) works 2 row pairs behind stage j-1:
//                A_j(p) = H-odd(j) + V-even(j)    on rows (2p,   2p+1),   at step s = p + 2j
//                B_j(p) = V-odd(j) + H-even(j+1)  on rows (2p+1, 2p+2),   at step s = p + 2j + 1
//              (the same 2x2-blocks-in-registers passes as the tile kernel)
//
// A step = { all A_j } barrier { all B_j } barrier with the entry work alongside, so every barrier interval carries the
// work of all K sweeps.  Dependencies: B_j(p) needs A_j(p) (previous step) and A_j(p+1) (same step, before the barrier);
// A_{j+1}(p) needs B_j(p-1) and B_j(p) (both in earlier steps) — hence the lag of two pairs per sweep.
// Rows are never recomputed (only the 2K halo rows above and below a chunk, trapezoid-wise as in the tile kernel), the
// column halo is 2K of 256 instead of 2K of 128, and HBM traffic is spread evenly over the whole launch.
//
// Threads: 16 worker warps (8 groups of 64 column pairs; a group serves two stages of one half-band) in two sets
// (stages 0..K/2-1 and the rest, see kWaveSets below) + 4 entry warps (one column pair per thread).  Every worker group
// executes the same instruction count per step; everything that is not a sweep is the entry warps' job, and they only
// join the end-of-step barrier (they have a whole step for their ~800 instructions at a fifth of a scheduler's issue
// slots).  Validity of the stale rim is decided per item (rows, warp-uniform) and per thread and stage (columns),
// exactly as in the tile kernel.  Local rows are offset by 2 so that pair 0 is an all-zero dummy pair: B_j(0) = rows
// (1, 2) then gives the first real row its H-even pass without a special case.
#pragma once
#include <type_traits>

#include "cloth_fused.cuh"
#include "cloth_stream1.cuh"

namespace clothgpu_impl {

constexpr int kWaveTW = 256, kWaveHW = 128;
#ifndef WAVE_WORKERS
#define WAVE_WORKERS 512
#endif
constexpr int kWaveWorkers = WAVE_WORKERS, kWaveEntry = 128, kWaveThreads = kWaveWorkers + kWaveEntry;  // (384 workers: experiments with K <= 6)
constexpr int kWaveGroups = kWaveWorkers / 64;  // worker groups of 64 column pairs
#ifndef WAVE_SETS
#define WAVE_SETS 2
#endif
#ifndef WAVE_STAGGER
#define WAVE_STAGGER 0
#endif
// Two worker sets: set X (groups 0-3) runs the first half of the sweep stages, set Y (groups 4-7) the second half.
// Barriers per step: alpha -> beta among the two groups that share their stages (128 threads: B_j only needs A_j),
// end of step within a set (X's with the entry warps); the sets are chained by two producer/consumer barriers ("X has
// finished step s", "Y has finished step s": bar.arrive / bar.sync), so they may drift apart by up to one step and one
// set's barrier waits, loads and address arithmetic overlap the other set's binary64 work on the same scheduler.
// Named barriers: 1 entry warps only; 2, 13 (X) and 14, 15 (Y) alpha -> beta; 12 end of step X + entry; 3 end of step Y;
// 4-7 "X done" and 8-11 "Y done" (step mod 4).
constexpr int kWaveSets = WAVE_SETS;
constexpr bool kWaveStagger = WAVE_STAGGER != 0;  // Y may start step s only after X's alpha phase of step s
constexpr int kBarEntry = 1, kBarX = 2, kBarY = 3, kBarXDone = 4 /* .. 7 */, kBarYDone = 8 /* .. 11 */, kBarXEnd = 12;
constexpr int kWaveSetThreads = kWaveWorkers / 2;
// With two worker sets the entry warps also write the finished rows back to HBM (they wait for "set Y has finished step
// s - 2" before they refill those ring slots anyway), so the last sweep stage does the same work as every other stage.
constexpr bool kWaveEntryStores = kWaveSets == 2;
static_assert(kWaveSets == 1 || kWaveWorkers == 512, "the two-set stage mapping is written for 8 worker groups");

#ifdef WAVE_ABLATE_BAR  // timing experiment only (results are wrong): what do the barriers cost?
__device__ __forceinline__ void bar_sync(int, int) {}
__device__ __forceinline__ void bar_arrive(int, int) {}
#else
__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }
#endif
constexpr int kWaveDepth = 3;                   // row pairs in flight in the staging ring
#ifndef WAVE_SLACK
#define WAVE_SLACK 0
#endif
// extra row pairs in the ring beyond the 2K + 2 the pipeline needs: set X (with the entry warps) may then run that many
// more steps ahead of set Y before the entry warps must wait for Y (2 + kWaveSlack steps)
constexpr int kWaveSlack = WAVE_SLACK;
#ifndef WAVE_TMA
#define WAVE_TMA 0
#endif
// Optional staging by the TMA engine (-DWAVE_TMA=1; 2 = refill right after the staging reads): one elected entry thread
// issues a bulk copy per plane row (256 columns = 2 KB in f64) that completes on a shared-memory mbarrier, instead of 128
// threads x 8 cp.async of 16 bytes each.  Parity-green, but MEASURED 2.7 % slower than the cp.async ring on 16384^2
// (17.5 vs 17.0 ms, profiles/README.md), so the cp.async ring stays the default.
constexpr bool kWaveTma = WAVE_TMA != 0;

#ifndef WAVE_BOTH
#define WAVE_BOTH 0
#endif
constexpr bool kWaveBoth = WAVE_BOTH != 0;      // both items of a thread at once (4 sticks in flight) or one after the other

struct WavePlan {
    int K, h, RB;     // sweeps, halo = 2K, ring rows = 4K + 4
    int bands;        // column bands per cloth
    int first_w, mid_w;
    int columns;      // band columns = bands * ensemble
    int span;         // row units per CTA (even)
    // Packed ensembles (pack != 0): cloths of at most 128 columns, two side by side in the 256-column band and many stacked
    // on top of each other, streamed as ONE tall virtual cloth (the Geom handed to the kernel describes the virtual cloth:
    // nx = 256, rows = pairs * rows_v).  A cloth's last column and last row head no stick, so the members do not interact;
    // no halo at all and one pipeline fill per CTA instead of per cloth.  Only the entry warps know about it: the
    // real geometry below tells them where a ring row lives.
    int pack;
    int nx, ny;       // real cloth size
    int rows_v;       // virtual rows per cloth (ny rounded up to even: row pairs never mix two cloths)
    int ensemble;     // real number of cloths
    int pitch;        // real elements between rows
    long long cloth_stride;  // real elements between cloths
};

template <typename T>
__host__ __device__ constexpr size_t wave_smem_bytes(int RB) {
    return (size_t)RB * kWaveHW * 2 * 2 * sizeof(T)                      // xyE, xyO
           + (size_t)RB * kWaveTW                                        // fl8
           + (size_t)(RB / 2) * kWaveHW * 2 * sizeof(uint16_t)           // stA, stB
           + (size_t)kWaveDepth * 2 * 4 * kWaveHW * 2 * sizeof(T)        // staging: pairs x rows x planes x 128 column pairs
           + 64;                                                         // one mbarrier per staging slot
}

template <typename T, bool PER_CLOTH, bool DRAG, bool PACK = false>
__global__ void __launch_bounds__(kWaveThreads, 1)
    wave_frame_kernel(Planes<T> in, Planes<T> out, Geom g, WavePlan pl, const FrameU<T> u0,
                      const FrameU<T>* __restrict__ per_cloth, DevCounters* ctr, const PeerOut<T> peer_up,
                      const PeerOut<T> peer_dn, const StripSync ss) {
    using V2 = typename Vec2<T>::type;
    constexpr int TW = kWaveTW, HW = kWaveHW;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int RB = pl.RB, HP = RB >> 1, K = pl.K, h = pl.h;
    V2* xyE = reinterpret_cast<V2*>(smem_raw);
    V2* xyO = xyE + RB * HW;
    uint8_t* fl8 = reinterpret_cast<uint8_t*>(xyO + RB * HW);
    uint16_t* stA = reinterpret_cast<uint16_t*>(fl8 + (size_t)RB * TW);
    uint16_t* stB = stA + HP * HW;
    V2* stage = reinterpret_cast<V2*>(stB + HP * HW);  // [kWaveDepth][2][4][HW]
    unsigned long long* mbar = reinterpret_cast<unsigned long long*>(stage + kWaveDepth * 2 * 4 * HW);  // [kWaveDepth]

    const int tid = threadIdx.x;
    prepare_next_frame_counters(ctr);
    unsigned int acc_live = 0, acc_torn = 0, acc_torn_post = 0, acc_extra = 0;
    ParticleEvents events;
    unsigned int stage_parity = 0;  // entry warps: bit d = phase parity of staging slot d's mbarrier (runs on across chunks)
    if (kWaveTma && tid == kWaveWorkers) {
        for (int d = 0; d < kWaveDepth; ++d) mbar_init(&mbar[d], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (kWaveTma) __syncthreads();

    // ---- the CTA's share of the work: a contiguous span of the row units of all band columns (band b of cloth e),
    // ---- laid end to end, so that every CTA streams the same number of rows whatever the cloth's shape.  A span that
    // ---- runs over the end of a band column continues at the top of the next one (the pipeline restarts).
    const int rows_e = (g.own_rows + 1) & ~1;  // row units per band column (even, so chunks start on even rows)
    const long long total_units = (long long)pl.columns * rows_e;
    long long ucur = (long long)blockIdx.x * pl.span;
    const long long uend = min(total_units, ucur + pl.span);
    while (ucur < uend) {  // block-uniform
    const int colidx = (int)(ucur / rows_e), r_start = (int)(ucur - (long long)colidx * rows_e);
    const int r_stop = (int)min((long long)rows_e, r_start + (uend - ucur));
    ucur += r_stop - r_start;
    if (r_start >= g.own_rows) continue;  // the padding unit of an odd-height cloth
    const int e = colidx / pl.bands;
    const FrameU<T> u = pick_frame<T, PER_CLOTH>(u0, per_cloth, e);
    const bool dragging = DRAG && ((PACK && PER_CLOTH) || (u.buttons & CLOTHGPU_MOUSE_DRAGGING) != 0);

    // ---- band (columns) and chunk (rows) ---------------------------------------------------------------------------
    const int bx = colidx % pl.bands;
    const int ix0 = bx == 0 ? 0 : pl.first_w + (bx - 1) * pl.mid_w;  // interior columns [ix0, ix1)
    const int ix1 = min(g.nx, bx == 0 ? pl.first_w : ix0 + pl.mid_w);
    const int ox = bx == 0 ? 0 : ix0 - h;  // band origin column
    const int cols = min(TW, g.nx - ox);
    const bool halo_left = ox > 0, halo_right = ox + TW < g.nx;
    const int lx0 = ix0 - ox, lx1 = ix1 - ox;
    const int held_end = g.held_row0 + g.held_rows, own_end = g.own_row0 + g.own_rows;
    const int ra = g.own_row0 + r_start, rb = min(g.own_row0 + r_stop, own_end);  // owned rows of this chunk
    const int y0 = max(g.held_row0, ra - h), y1 = min(held_end, rb + h);                            // rows streamed
    const bool halo_top = y0 > g.held_row0, halo_bottom = y1 < held_end;
    const int n_lr = y1 - y0 + 2;  // local rows: 0, 1 = dummy zero pair, real rows [2, n_lr)
    const int ly0 = ra - y0 + 2, ly1 = rb - y0 + 2;  // owned rows, local
    const long long cloth_base = (long long)e * g.cloth_stride;
    auto row_off = [&](int lr) -> long long {  // element offset of local row lr, column ox
        return cloth_base + (long long)(y0 + lr - 2 - g.held_row0) * g.pitch + ox;
    };
    auto slot = [&](int lr) -> int { return lr % RB; };

    // Attached strips: a chunk that stores rows a neighbour keeps as ghosts (the only chunks whose halo reaches into ghost
    // rows) waits for that neighbour's epoch before it loads — the entry warps do all global traffic — and reports when
    // its stores are out (cloth_types.cuh, StripSync); every other chunk starts right away.
    const bool st_up = ss.on() && ra < peer_up.row_hi && rb > peer_up.row_lo;
    const bool st_dn = ss.on() && ra < peer_dn.row_hi && rb > peer_dn.row_lo;

    auto owned = [&](int lr, int c) -> bool { return lr >= ly0 && lr < ly1 && c >= lx0 && c < lx1; };
    // a torn stick (after the block states were derived): clear the alive bit of its head's flag byte
    auto tear = [&](int lr, int c, uint32_t alive_bit, int sweep) {
        fl8[slot(lr) * TW + c] &= (uint8_t)~alive_bit;
        if (owned(lr, c)) acc_torn += 1, acc_torn_post += 1, acc_extra += (unsigned)(sweep + 1);
    };

    // ---- prologue: the dummy pair -----------------------------------------------------------------------------------
    for (int i = tid; i < 2 * HW; i += kWaveThreads) {
        V2 z;
        z.x = z.y = T(0);
        xyE[i] = z, xyO[i] = z;  // rows 0 and 1 are slots 0 and 1
    }
    for (int i = tid; i < 2 * TW; i += kWaveThreads) fl8[i] = 0;
    // (stA of pair 0 is never read: A_j(0) is skipped; stB of pair 0 is written by the entry warps at step 0)

    const int last_pair = (n_lr + 1) >> 1;  // entry also delivers one all-zero pair past the end (row n_lr may be a B tail)
    const int n_steps = last_pair + 2 * K + 1 + kWaveSlack;  // (the last pair leaves the ring 2K + 2 + slack steps after it came)

    // ---- the state word of a 2x2 block (cloth_fused.cuh): live bits of its four sticks, pin bits of its corners ------
    auto live_bit = [](uint32_t fh, uint32_t ft, uint32_t alive) -> uint32_t {
        return ((fh & alive) && (fh & ft & CLOTHGPU_FLAG_ACTIVE)) ? 1u : 0u;  // cloth.go:97
    };
    // (tear4: packed ensembles with one frame block per cloth — bits 12..15 say which of the block's four sticks belong to
    //  a cloth that is being dragged, in the order of the live bits)
    auto pack = [&](uint32_t fa, uint32_t fb, uint32_t fc, uint32_t fd, uint32_t first_alive, uint32_t second_alive,
                    bool first_ok, uint32_t tear4 = 0u) -> uint32_t {
        uint32_t st = 0;
        if (first_ok) st |= live_bit(fa, fb, first_alive) | (live_bit(fc, fd, first_alive) << 1);
        st |= (live_bit(fa, fc, second_alive) << 2) | (live_bit(fb, fd, second_alive) << 3);
        const uint32_t pa = fa & CLOTHGPU_FLAG_PIN, pb = fb & CLOTHGPU_FLAG_PIN, pc = fc & CLOTHGPU_FLAG_PIN,
                       pd = fd & CLOTHGPU_FLAG_PIN;  // CLOTHGPU_FLAG_PIN == 1
        st |= (pa << 4) | (pb << 5) | (pc << 6) | (pd << 7);
        st |= (pa << 8) | (pc << 9) | (pb << 10) | (pd << 11);
        st |= tear4 << 12;
        return st;
    };

    if (tid >= kWaveWorkers) {
        // =============================== entry warps ===============================================================
        const int et = tid - kWaveWorkers;  // 0..127: the column pair this thread delivers, both rows of a pair
        if (st_up) ss.wait(0);
        if (st_dn) ss.wait(1);
        const int c = 2 * et;
        // where local row lr of this thread's column pair lives in HBM, and whether it exists at all
        const int pk_half = PACK ? et >> 6 : 0, pk_cc = PACK ? 2 * (et & 63) : c;
        const int col_end = PACK ? pl.nx : cols;  // the pair's first column exists iff pk_cc < col_end
        int pk_lr = 2, pk_m = 0, pk_rr = 0;  // packed ensembles: local row 2q of the pair being delivered = row pk_rr of cloth pair pk_m
        if (PACK) pk_m = y0 / pl.rows_v, pk_rr = y0 - pk_m * pl.rows_v;
        // Packed ensembles with one frame block per cloth: the entry warps work with the frame of the cloth their column
        // half of the pair being delivered belongs to (reloaded when the pair moves on to the next cloth); the sweeps only
        // need what the host checked to be common to all cloths (rest length, gain, tear length) plus, per stick, whether
        // its cloth is being dragged — that goes into the block state words.
        // (the frame is read through the pointer where it is used — all threads of a half read the same words, they stay in
        //  L1 — instead of 40 more live registers per entry thread)
        constexpr bool kPerClothPack = PACK && PER_CLOTH;
        const FrameU<T>* uep = per_cloth;  // only dereferenced when kPerClothPack
        uint32_t dr_own = dragging ? 1u : 0u, dr_other = dr_own, dr_prev = dr_own;
        int ue_m = -1;  // cloth pair `uep` belongs to
        auto cloth_frame = [&](int m) {
            if (!kPerClothPack || m == ue_m) return;
            ue_m = m;
            const int own = min(2 * m + pk_half, pl.ensemble - 1), other = min(2 * m + 1 - pk_half, pl.ensemble - 1);
            uep = per_cloth + own;
            dr_own = (uep->buttons & CLOTHGPU_MOUSE_DRAGGING) ? 1u : 0u;
            dr_other = (per_cloth[other].buttons & CLOTHGPU_MOUSE_DRAGGING) ? 1u : 0u;
        };
        struct RowAt {
            long long at;   // element offset of (row, first column of the pair)
            bool ok;        // the row exists (and the pair's first column with it)
            bool last_row;  // packed ensembles: the cloth's last row heads no vertical stick
        };
        auto row_at = [&](int lr) -> RowAt {
            RowAt r;
            r.last_row = false;
            if (!PACK) {
                r.ok = lr >= 2 && lr < n_lr && c < cols;
                r.at = r.ok ? row_off(lr) + c : 0;
                return r;
            }
            // (cloth pair, row) of lr from those of the pair being delivered: the rows asked for are at most 2K + 2 pairs
            // behind and kWaveDepth pairs ahead, so a few steps instead of a division
            int rr = pk_rr + (lr - pk_lr), m = pk_m;
            while (rr >= pl.rows_v) rr -= pl.rows_v, ++m;
            while (rr < 0) rr += pl.rows_v, --m;
            const int cid = 2 * m + pk_half;
            r.ok = lr >= 2 && lr < n_lr && rr < pl.ny && cid < pl.ensemble && pk_cc < pl.nx;
            r.at = r.ok ? (long long)cid * pl.cloth_stride + (long long)rr * pl.pitch + pk_cc : 0;
            r.last_row = rr == pl.ny - 1;
            return r;
        };
        auto st_slot = [&](int q, int row, int plane) -> V2* {
            return stage + ((((q % kWaveDepth) * 2 + row) * 4 + plane) * HW + et);
        };
        // columns a bulk copy brings in per plane row: this band's columns, rounded up to 16 bytes (the row padding up to
        // `pitch` exists and is never a live particle)
        const unsigned copy_bytes = (unsigned)((cols * (int)sizeof(T) + 15) & ~15);
        auto issue_pair = [&](int q) {  // local rows 2q, 2q+1 -> staging
            if (kWaveTma) {
                if (et != 0 || q > last_pair) return;  // one thread drives the TMA engine; nothing past the chunk's end
                unsigned long long* b = &mbar[q % kWaveDepth];
                const int rows_ok = (2 * q < n_lr ? 1 : 0) + (2 * q + 1 < n_lr ? 1 : 0);
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier reads of the slot precede the refill
                mbar_arrive_expect_tx(b, 4u * copy_bytes * (unsigned)rows_ok);
                for (int row = 0; row < rows_ok; ++row) {
                    const long long at = row_off(2 * q + row);
                    V2* dst = stage + (((q % kWaveDepth) * 2 + row) * 4) * HW;
                    bulk_g2s(dst, in.x + at, copy_bytes, b);
                    bulk_g2s(dst + HW, in.y + at, copy_bytes, b);
                    bulk_g2s(dst + 2 * HW, in.px + at, copy_bytes, b);
                    bulk_g2s(dst + 3 * HW, in.py + at, copy_bytes, b);
                }
                return;
            }
#pragma unroll
            for (int row = 0; row < 2; ++row) {
                const RowAt ra_ = row_at(2 * q + row);
                const bool ok = ra_.ok;
                const long long at = ra_.at;
                cp_async_lane<sizeof(V2)>(st_slot(q, row, 0), in.x + at, ok);
                cp_async_lane<sizeof(V2)>(st_slot(q, row, 1), in.y + at, ok);
                cp_async_lane<sizeof(V2)>(st_slot(q, row, 2), in.px + at, ok);
                cp_async_lane<sizeof(V2)>(st_slot(q, row, 3), in.py + at, ok);
            }
            cp_async_commit();
        };
        auto load_flags = [&](int q, int row) -> uint32_t {
            const RowAt ra_ = row_at(2 * q + row);
            uint32_t f = 0;
            if (ra_.ok) {
                f = *reinterpret_cast<const unsigned short*>(in.fl + ra_.at);
                // (the value is consumed a step later: without PACK nothing touches it here, so the warp does not wait for
                //  the load; the non-existent column nx of the row padding is masked at the point of use)
                if (PACK) {
                    if (pk_cc + 1 >= col_end) f &= 0xffu;
                    // Next to / below a cloth's edge lies ANOTHER cloth here: the edge particles must head no stick.  Their
                    // alive bits (meaningless, but part of the state the caller may have written) are parked in the spare
                    // bits 5 / 6 of the flag byte — no pass looks at those — and restored when the row is written back.
                    uint32_t hide = 0;
                    if (ra_.last_row) hide |= CLOTHGPU_FLAG_ALIVE_V | (CLOTHGPU_FLAG_ALIVE_V << 8);
                    if (pk_cc == pl.nx - 1) hide |= CLOTHGPU_FLAG_ALIVE_H;
                    if (pk_cc + 1 == pl.nx - 1) hide |= CLOTHGPU_FLAG_ALIVE_H << 8;
                    f = (f & ~hide) | ((f & hide) << 2);  // ALIVE_H (8) -> 0x20, ALIVE_V (16) -> 0x40
                }
            }
            return f;
        };
#pragma unroll
        for (int q = 1; q <= kWaveDepth; ++q) issue_pair(q);
        uint32_t fnext[2] = {load_flags(1, 0), load_flags(1, 1)};
        __syncthreads();  // (matches the workers' prologue barrier)
        int eslot = 2;  // ring slot of local row 2q (q = 1 at the first step)
        uint32_t fprev[2] = {0u, 0u};  // flags of the previous pair's second row (columns 2et, 2et+1); the dummy pair has none
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            const int q = s + 1;  // the pair delivered during this step
            // the slots refilled in this step held the pair that set Y wrote back in step s - 2
            if (kWaveSets == 2 && s >= 2 + kWaveSlack) bar_sync(kBarYDone + ((s - 2 - kWaveSlack) & 3), kWaveSetThreads + kWaveEntry);
            if (kWaveEntryStores) {
                // ---- the pair that has passed all K sweeps (local rows 2qo, 2qo+1, in the ring slots refilled below) leaves ----
                const int qo = q - HP;
                if (qo >= 1 && c >= lx0 && c < lx1) {
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int lr = 2 * qo + w;
                        const RowAt ra_ = row_at(lr);
                        if (lr >= ly0 && lr < ly1 && ra_.ok) {
                            const V2 pe = xyE[(eslot + w) * HW + et], po = xyO[(eslot + w) * HW + et];
                            uint32_t ff = *reinterpret_cast<const uint16_t*>(fl8 + (eslot + w) * TW + c);
                            if (PACK) ff = (ff & 0x1f1fu) | ((ff & 0x6060u) >> 2);  // the parked edge bits come back
                            const long long at = ra_.at;
                            const int gy = y0 + lr - 2, gx = ox + c;
                            V2 ox2, oy2;
                            ox2.x = pe.x, ox2.y = po.x, oy2.x = pe.y, oy2.y = po.y;
                            *reinterpret_cast<V2*>(out.x + at) = ox2;
                            *reinterpret_cast<V2*>(out.y + at) = oy2;
                            *reinterpret_cast<unsigned short*>(out.fl + at) = (unsigned short)ff;
                            if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {
                                const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + gx;
                                *reinterpret_cast<V2*>(peer_up.x + pat) = ox2;
                                *reinterpret_cast<V2*>(peer_up.y + pat) = oy2;
                                *reinterpret_cast<unsigned short*>(peer_up.fl + pat) = (unsigned short)ff;
                            }
                            if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
                                const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + gx;
                                *reinterpret_cast<V2*>(peer_dn.x + pat) = ox2;
                                *reinterpret_cast<V2*>(peer_dn.y + pat) = oy2;
                                *reinterpret_cast<unsigned short*>(peer_dn.fl + pat) = (unsigned short)ff;
                            }
                        }
                    }
                }
            }
            if (q <= last_pair) {
                if (kWaveTma) {
                    mbar_wait(&mbar[q % kWaveDepth], (stage_parity >> (q % kWaveDepth)) & 1u);
                    stage_parity ^= 1u << (q % kWaveDepth);
                } else {
                    cp_async_wait<kWaveDepth - 1>();
                }
                cloth_frame(pk_m);  // (pk_m: the cloth pair of the pair being delivered)
                const bool drag_e = kPerClothPack ? dr_own != 0u : dragging;
                const uint32_t fkeep = (PACK || pk_cc + 1 < col_end) ? 0xffffu : 0xffu;
                const uint32_t fcur[2] = {fnext[0] & fkeep, fnext[1] & fkeep};
                fnext[0] = load_flags(q + 1, 0), fnext[1] = load_flags(q + 1, 1);
                V2 p0[2], p1[2];
                uint32_t f0[2], f1[2];
#pragma unroll
                for (int row = 0; row < 2; ++row) {
                    const int lr = 2 * q + row;
                    V2 vx = *st_slot(q, row, 0), vy = *st_slot(q, row, 1);
                    V2 vpx = *st_slot(q, row, 2), vpy = *st_slot(q, row, 3);
                    if (kWaveTma && !(lr < n_lr && c < cols)) {  // a bulk copy does not zero-fill what does not exist
                        vx.x = vx.y = vy.x = vy.y = T(0);
                        vpx = vx, vpy = vx;
                    }
                    p0[row].x = vx.x, p1[row].x = vx.y, p0[row].y = vy.x, p1[row].y = vy.y;
                    f0[row] = fcur[row] & 0xffu, f1[row] = fcur[row] >> 8;
                    const RowAt ra_ = row_at(lr);
                    if (ra_.ok) {
                        T px0 = vpx.x, px1 = vpx.y, py0 = vpy.x, py1 = vpy.y;
                        const uint32_t f0_in = f0[row], f1_in = f1[row];
                        if (kPerClothPack) {
                            particle_update(p0[row].x, p0[row].y, px0, py0, f0[row], *uep);
                            if (pk_cc + 1 < col_end) particle_update(p1[row].x, p1[row].y, px1, py1, f1[row], *uep);
                        } else {
                            particle_update(p0[row].x, p0[row].y, px0, py0, f0[row], u);
                            if (pk_cc + 1 < col_end) particle_update(p1[row].x, p1[row].y, px1, py1, f1[row], u);
                        }
                        if (owned(lr, c)) {  // the owner of a particle writes its previous position
                            events.note2(f0_in & 0x1fu, f0[row] & 0x1fu, f1_in & 0x1fu, f1[row] & 0x1fu);
                            const long long at = ra_.at;
                            const int gy = y0 + lr - 2, gx = ox + c;
                            V2 ox2, oy2;
                            ox2.x = px0, ox2.y = px1, oy2.x = py0, oy2.y = py1;
                            *reinterpret_cast<V2*>(out.px + at) = ox2;
                            *reinterpret_cast<V2*>(out.py + at) = oy2;
                            if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {
                                const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + gx;
                                *reinterpret_cast<V2*>(peer_up.px + pat) = ox2;
                                *reinterpret_cast<V2*>(peer_up.py + pat) = oy2;
                            }
                            if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
                                const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + gx;
                                *reinterpret_cast<V2*>(peer_dn.px + pat) = ox2;
                                *reinterpret_cast<V2*>(peer_dn.py + pat) = oy2;
                            }
                        }
                    }
                }
                if (!kWaveTma) issue_pair(q + kWaveDepth);  // refill the staging slots just read (always commits: uniform group counting)
#if WAVE_TMA == 2  // refill as early as possible: one more entry-only barrier right after the staging reads
                bar_sync(kBarEntry, kWaveEntry);
                issue_pair(q + kWaveDepth);
#endif
                {  // H-even of sweep 0: the two rows of this column pair
                    V2* hh[2] = {&p0[0], &p0[1]};
                    V2* tt[2] = {&p1[0], &p1[1]};
                    uint32_t live = 0, pins = 0;
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        if ((f0[w] & CLOTHGPU_FLAG_ALIVE_H) && (f0[w] & f1[w] & CLOTHGPU_FLAG_ACTIVE)) live |= 1u << w;  // cloth.go:97
                        pins |= ((f0[w] & CLOTHGPU_FLAG_PIN) | ((f1[w] & CLOTHGPU_FLAG_PIN) << 1)) << (2 * w);
                    }
                    bool touched = false;
                    // (the sticks only read what all cloths share — rest length, gain, tear length — so `u` serves every cloth)
                    const uint32_t tore = relax_n<T, DRAG, 2>(hh, tt, live, pins, u, drag_e, touched);
                    if (DRAG && tore) {
#pragma unroll
                        for (int w = 0; w < 2; ++w)
                            if (tore & (1u << w)) {
                                f0[w] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_H;
                                if (owned(2 * q + w, c)) acc_torn += 1, acc_extra += 1;
                            }
                    }
                }
#pragma unroll
                for (int row = 0; row < 2; ++row) {
                    const int sl = eslot + row;
                    xyE[sl * HW + et] = p0[row];
                    xyO[sl * HW + et] = p1[row];
                    *reinterpret_cast<uint16_t*>(fl8 + sl * TW + c) = (uint16_t)(f0[row] | (f1[row] << 8));
                }
                // ---- block states of this pair, derived HERE and not by sweep stage 0: the entry warps have the flags in
                // ---- registers and time to spare, and every worker group then does the same amount of work per step ----
                bar_sync(kBarEntry, kWaveEntry);  // the entry warps' flag bytes are in the ring
                if (WAVE_TMA == 1) issue_pair(q + kWaveDepth);  // every entry thread has read the staging slot: refill it
                {   // group A, pair q: rows (2q, 2q+1) x columns (2et+1, 2et+2); heads a (H, V), c (H), b (V)
                    const int er = (et + 1) & (HW - 1), cl = c + 1, cr = 2 * er;
                    const uint32_t fb = fl8[eslot * TW + cr], fd = fl8[(eslot + 1) * TW + cr];
                    // heads: a (H-odd, V-even), c (H-odd): this thread's odd column; b (V-even): the even column to the right,
                    // which belongs to the other cloth of the pair when it lies in the other half of the band
                    const uint32_t dr_b = (er >> 6) == pk_half ? dr_own : dr_other;
                    const uint32_t tearA = kPerClothPack ? (dr_own | (dr_own << 1) | (dr_own << 2) | (dr_b << 3)) : 0u;
                    const uint32_t st = pack(f1[0], fb, f1[1], fd, CLOTHGPU_FLAG_ALIVE_H, CLOTHGPU_FLAG_ALIVE_V, et + 1 < HW, tearA);
                    stA[(eslot >> 1) * HW + et] = (uint16_t)st;
                    const bool oa = owned(2 * q, cl), oc = owned(2 * q + 1, cl), ob = owned(2 * q, cr);
                    acc_live += (oa ? (st & 1u) + ((st >> 2) & 1u) : 0u) + (oc ? (st >> 1) & 1u : 0u) + (ob ? (st >> 3) & 1u : 0u);
                }
                {   // group B, pair q-1: rows (2q-1, 2q) x columns (2et, 2et+1); heads a (V, H), c (V), b (H)
                    const int s2p = eslot == 0 ? RB - 2 : eslot - 2;  // slot of row 2(q-1)
                    // heads: a (V-odd, H-even), c (V-odd): row 2q-1, the cloth of the previous pair; b (H-even): row 2q
                    const uint32_t tearB = kPerClothPack ? (dr_prev | (dr_prev << 1) | (dr_prev << 2) | (dr_own << 3)) : 0u;
                    const uint32_t st = pack(fprev[0], f0[0], fprev[1], f1[0], CLOTHGPU_FLAG_ALIVE_V, CLOTHGPU_FLAG_ALIVE_H, true, tearB);
                    stB[(s2p >> 1) * HW + et] = (uint16_t)st;
                    const bool oa = owned(2 * q - 1, c), oc = owned(2 * q - 1, c + 1), ob = owned(2 * q, c);
                    acc_live += (oa ? (st & 1u) + ((st >> 2) & 1u) : 0u) + (oc ? (st >> 1) & 1u : 0u) + (ob ? (st >> 3) & 1u : 0u);
                }
                fprev[0] = f0[1], fprev[1] = f1[1];
                dr_prev = dr_own;
            }
            if (kWaveSets == 2) {
                // The pair delivered in step s is first read in step s + 1, and nothing it writes is touched by set X
                // before: the entry warps only join the END-of-step barrier, so they have the whole step for their work
                // (one entry warp per scheduler gets a fifth of the issue slots: ~760 instructions take about a phase).
                bar_sync(kBarXEnd, kWaveSetThreads + kWaveEntry);
            } else {
                __syncthreads();  // end of phase alpha
                __syncthreads();  // end of phase beta
            }
            eslot = eslot + 2 == RB ? 0 : eslot + 2;
            if (PACK) {
                pk_lr += 2, pk_rr += 2;
                if (pk_rr >= pl.rows_v) pk_rr -= pl.rows_v, ++pk_m;
            }
        }
        if (!kWaveTma) cp_async_wait<0>();  // nothing may still be landing in the staging slots when the next chunk starts
        if (st_up || st_dn) __threadfence_system();  // my stores into the neighbours' ghost rows, before the barrier below
    } else {
        // =============================== worker warps ==============================================================
        // A thread works for the same (at most two) sweep stages and the same column pair during the whole launch, so
        // everything that depends on them is computed once: column masks, row ranges, and the ring slots advance by two
        // rows per step.
        const int grp = tid >> 6, k64 = tid & 63;
        const int set = kWaveSets == 2 ? grp >> 2 : 0;           // 0 = X, 1 = Y
        const int Kx = kWaveSets == 2 ? (K + 1) >> 1 : K;        // stages [0, Kx) belong to X, [Kx, K) to Y
        const int set_j0 = set ? Kx : 0, set_nj = set ? K - Kx : Kx;
        const int k = (grp & 1) * 64 + k64, kr = (k + 1) & (HW - 1);
        const int cl = 2 * k + 1, cr = 2 * kr, ce = 2 * k, co = ce + 1;
        constexpr int kItems = 2;  // ceil(2 * 8 / kWaveGroups)
        int jj[kItems], e2[kItems], rloA[kItems], rhiA[kItems], rloB[kItems], rhiB[kItems];
        uint32_t maskA[kItems], maskB[kItems];
        bool on[kItems];
#pragma unroll
        for (int i = 0; i < kItems; ++i) {
            const int jl = kWaveSets == 2 ? ((grp & 3) >> 1) + 2 * i : (grp >> 1) + i * (kWaveGroups / 2);  // stage within the set
            const int j = set_j0 + jl;
            jj[i] = j;
            on[i] = jl < set_nj;
            e2[i] = j == 0 ? 0 : RB - 4 * j;  // slot of row 2p for p = -2j (step 0); 4j < RB
            const int c0 = halo_left ? 2 * j : 0, c1 = halo_right ? TW - 2 * j : cols;              // sweep j's columns
            const int d0 = halo_left ? 2 * j + 2 : 0, d1 = halo_right ? TW - 2 * j - 2 : cols;      // sweep j+1's columns
            maskA[i] = (cl >= c0 && cl + 1 < c1) ? 0xfu : kStSecond;                                // H-odd needs both columns
            maskB[i] = (j + 1 < K && ce >= d0 && co < d1) ? 0xfu : kStFirst;                        // H-even(j+1) likewise
            rloA[i] = halo_top ? 2 + 2 * j : 2, rhiA[i] = halo_bottom ? n_lr - 2 * j : n_lr;
            rloB[i] = halo_top ? 2 + 2 * j : 0, rhiB[i] = halo_bottom ? n_lr - 2 * j : n_lr + 1;
        }
        __syncthreads();  // prologue: the dummy pair is in place
        using pins_none_t = std::integral_constant<int, kPinsNone>;
        using pins_each_t = std::integral_constant<int, kPinsEach>;
        // One phase for NI of this thread's items at once (2 * NI sticks in flight per pass: the relaxation of a stick is
        // one long dependent binary64 chain, and with 4 worker warps per scheduler the pipe needs the extra chains).
        auto phase_alpha = [&](auto ni, int i0, int s) {
            constexpr int NI = decltype(ni)::value, NS = 2 * NI;
            V2 pa[NI], pb[NI], pc[NI], pd[NI];
            uint32_t st[NI];
            int sa[NI], pp[NI], jn[NI];
            uint32_t live1 = 0, live2 = 0, pins1 = 0, pins2 = 0, tear1 = 0xffffffffu, tear2 = 0xffffffffu;
            if (PACK && PER_CLOTH) tear1 = tear2 = 0u;
#pragma unroll
            for (int n = 0; n < NI; ++n) {
                const int i = i0 + n;
                jn[n] = jj[i], pp[n] = s - 2 * jj[i], sa[n] = e2[i];  // slots of rows 2p, 2p+1 = sa, sa+1; state slot sa/2
                const int sb = sa[n] + 1;
                pa[n] = xyO[sa[n] * HW + k], pb[n] = xyE[sa[n] * HW + kr], pc[n] = xyO[sb * HW + k], pd[n] = xyE[sb * HW + kr];
                st[n] = stA[(sa[n] >> 1) * HW + k];  // (derived by the entry warps when the pair arrived)
                const uint32_t sm = st[n] & maskA[i];
                live1 |= (sm & 3u) << (2 * n), live2 |= ((sm >> 2) & 3u) << (2 * n);
                pins1 |= ((st[n] >> 4) & 15u) << (4 * n), pins2 |= ((st[n] >> 8) & 15u) << (4 * n);
                if (PACK && PER_CLOTH) tear1 |= ((st[n] >> 12) & 3u) << (2 * n), tear2 |= ((st[n] >> 14) & 3u) << (2 * n);
            }
            bool touched = false;
            uint32_t tore1 = 0, tore2 = 0;
            auto passes = [&](auto pm) {
                constexpr int PM = decltype(pm)::value;
                {  // H-odd: a->b and c->d
                    V2* hh[NS];
                    V2* tt[NS];
#pragma unroll
                    for (int n = 0; n < NI; ++n) hh[2 * n] = &pa[n], tt[2 * n] = &pb[n], hh[2 * n + 1] = &pc[n], tt[2 * n + 1] = &pd[n];
                    tore1 = relax_n<T, DRAG, NS, PM>(hh, tt, live1, pins1, u, dragging, touched, tear1);
                }
                {  // V-even: a->c and b->d
                    V2* hh[NS];
                    V2* tt[NS];
#pragma unroll
                    for (int n = 0; n < NI; ++n) hh[2 * n] = &pa[n], tt[2 * n] = &pc[n], hh[2 * n + 1] = &pb[n], tt[2 * n + 1] = &pd[n];
                    tore2 = relax_n<T, DRAG, NS, PM>(hh, tt, live2, pins2, u, dragging, touched, tear2);
                }
            };
            // pinned particles are rare (the cloth's pinned rows): one vote per item picks the pass code without pin logic
            if (__any_sync(0xffffffffu, (pins1 | pins2) != 0u)) passes(pins_each_t{});
            else passes(pins_none_t{});
            if (DRAG && (tore1 | tore2)) {  // rare
#pragma unroll
                for (int n = 0; n < NI; ++n) {
                    const uint32_t t1 = (tore1 >> (2 * n)) & 3u, t2 = (tore2 >> (2 * n)) & 3u;
                    if (t1 | t2) {
                        stA[(sa[n] >> 1) * HW + k] = (uint16_t)(st[n] & ~(t1 | (t2 << 2)));
                        if (t1 & 1u) tear(2 * pp[n], cl, CLOTHGPU_FLAG_ALIVE_H, jn[n]);
                        if (t1 & 2u) tear(2 * pp[n] + 1, cl, CLOTHGPU_FLAG_ALIVE_H, jn[n]);
                        if (t2 & 1u) tear(2 * pp[n], cl, CLOTHGPU_FLAG_ALIVE_V, jn[n]);
                        if (t2 & 2u) tear(2 * pp[n], cr, CLOTHGPU_FLAG_ALIVE_V, jn[n]);
                    }
                }
            }
            {   // (written back even when no stick moved: a store costs less than the branch around it)
#pragma unroll
                for (int n = 0; n < NI; ++n) {
                    const int sb = sa[n] + 1;
                    xyO[sa[n] * HW + k] = pa[n], xyE[sa[n] * HW + kr] = pb[n], xyO[sb * HW + k] = pc[n], xyE[sb * HW + kr] = pd[n];
                }
            }
        };
        auto phase_beta = [&](auto ni, int i0, int s) {
            constexpr int NI = decltype(ni)::value, NS = 2 * NI;
            V2 pa[NI], pb[NI], pc[NI], pd[NI];
            uint32_t st[NI];
            int sa[NI], sb[NI], pp[NI], jn[NI];
            uint32_t live1 = 0, live2 = 0, pins1 = 0, pins2 = 0, tear1 = 0xffffffffu, tear2 = 0xffffffffu;
            if (PACK && PER_CLOTH) tear1 = tear2 = 0u;
            bool any_mid = false;  // some item still has an H-even pass to do (it is not the last sweep's)
#pragma unroll
            for (int n = 0; n < NI; ++n) {
                const int i = i0 + n;
                jn[n] = jj[i], pp[n] = s - 2 * jj[i] - 1;
                const int s2p = e2[i] == 0 ? RB - 2 : e2[i] - 2;  // slot of row 2p (this pair is one behind A's)
                sa[n] = s2p + 1, sb[n] = e2[i];
                pa[n] = xyE[sa[n] * HW + k], pc[n] = xyO[sa[n] * HW + k], pb[n] = xyE[sb[n] * HW + k], pd[n] = xyO[sb[n] * HW + k];
                st[n] = stB[(s2p >> 1) * HW + k];
                const uint32_t sm = st[n] & maskB[i];  // (the last sweep's mask has no second-pass bits)
                live1 |= (sm & 3u) << (2 * n), live2 |= ((sm >> 2) & 3u) << (2 * n);
                pins1 |= ((st[n] >> 4) & 15u) << (4 * n), pins2 |= ((st[n] >> 8) & 15u) << (4 * n);
                if (PACK && PER_CLOTH) tear1 |= ((st[n] >> 12) & 3u) << (2 * n), tear2 |= ((st[n] >> 14) & 3u) << (2 * n);
                any_mid |= jn[n] + 1 < K;
            }
            bool touched = false;
            uint32_t tore1 = 0, tore2 = 0;
            auto passes = [&](auto pm) {
                constexpr int PM = decltype(pm)::value;
                {  // V-odd: a->b and c->d
                    V2* hh[NS];
                    V2* tt[NS];
#pragma unroll
                    for (int n = 0; n < NI; ++n) hh[2 * n] = &pa[n], tt[2 * n] = &pb[n], hh[2 * n + 1] = &pc[n], tt[2 * n + 1] = &pd[n];
                    tore1 = relax_n<T, DRAG, NS, PM>(hh, tt, live1, pins1, u, dragging, touched, tear1);
                }
                if (any_mid) {  // H-even of the next sweep: a->c and b->d (warp-uniform)
                    V2* hh[NS];
                    V2* tt[NS];
#pragma unroll
                    for (int n = 0; n < NI; ++n) hh[2 * n] = &pa[n], tt[2 * n] = &pc[n], hh[2 * n + 1] = &pb[n], tt[2 * n + 1] = &pd[n];
                    tore2 = relax_n<T, DRAG, NS, PM>(hh, tt, live2, pins2, u, dragging, touched, tear2);
                }
            };
            if (__any_sync(0xffffffffu, (pins1 | pins2) != 0u)) passes(pins_each_t{});
            else passes(pins_none_t{});
            if (DRAG && (tore1 | tore2)) {  // rare
#pragma unroll
                for (int n = 0; n < NI; ++n) {
                    const uint32_t t1 = (tore1 >> (2 * n)) & 3u, t2 = (tore2 >> (2 * n)) & 3u;
                    if (t1 | t2) {
                        stB[((sa[n] - 1) >> 1) * HW + k] = (uint16_t)(st[n] & ~(t1 | (t2 << 2)));
                        if (t1 & 1u) tear(2 * pp[n] + 1, ce, CLOTHGPU_FLAG_ALIVE_V, jn[n]);
                        if (t1 & 2u) tear(2 * pp[n] + 1, co, CLOTHGPU_FLAG_ALIVE_V, jn[n]);
                        if (t2 & 1u) tear(2 * pp[n] + 1, ce, CLOTHGPU_FLAG_ALIVE_H, jn[n] + 1);
                        if (t2 & 2u) tear(2 * pp[n] + 2, ce, CLOTHGPU_FLAG_ALIVE_H, jn[n] + 1);
                    }
                }
            }
#pragma unroll
            for (int n = 0; n < NI; ++n) {
                if (kWaveEntryStores || jn[n] + 1 < K) {
                    xyE[sa[n] * HW + k] = pa[n], xyO[sa[n] * HW + k] = pc[n], xyE[sb[n] * HW + k] = pb[n], xyO[sb[n] * HW + k] = pd[n];
                } else if (ce >= lx0 && ce < lx1) {
                    // ---- the last colour pass is done: rows 2p+1 and 2p+2 leave through the registers --------------------------
                    const uint32_t fa2 = *reinterpret_cast<const uint16_t*>(fl8 + sa[n] * TW + ce);  // (after possible tears of this pass)
                    const uint32_t fb2 = *reinterpret_cast<const uint16_t*>(fl8 + sb[n] * TW + ce);
#pragma unroll
                    for (int w = 0; w < 2; ++w) {
                        const int lr = 2 * pp[n] + 1 + w;
                        if (lr >= ly0 && lr < ly1) {
                            const V2 pe = w ? pb[n] : pa[n], po = w ? pd[n] : pc[n];
                            const uint32_t ff = w ? fb2 : fa2;
                            const long long at = row_off(lr) + ce;
                            const int gy = y0 + lr - 2, gx = ox + ce;
                            V2 ox2, oy2;
                            ox2.x = pe.x, ox2.y = po.x, oy2.x = pe.y, oy2.y = po.y;
                            *reinterpret_cast<V2*>(out.x + at) = ox2;
                            *reinterpret_cast<V2*>(out.y + at) = oy2;
                            *reinterpret_cast<unsigned short*>(out.fl + at) = (unsigned short)ff;
                            if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {
                                const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + gx;
                                *reinterpret_cast<V2*>(peer_up.x + pat) = ox2;
                                *reinterpret_cast<V2*>(peer_up.y + pat) = oy2;
                                *reinterpret_cast<unsigned short*>(peer_up.fl + pat) = (unsigned short)ff;
                            }
                            if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
                                const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + gx;
                                *reinterpret_cast<V2*>(peer_dn.x + pat) = ox2;
                                *reinterpret_cast<V2*>(peer_dn.y + pat) = oy2;
                                *reinterpret_cast<unsigned short*>(peer_dn.fl + pat) = (unsigned short)ff;
                            }
                        }
                    }
                }
            }
        };
        using one_t = std::integral_constant<int, 1>;
        using two_t = std::integral_constant<int, 2>;
#pragma unroll 1
        for (int s = 0; s < n_steps; ++s) {
            // set Y's first stage reads what set X's last stage finished in step s - 1
            if (kWaveSets == 2 && set == 1 && s >= 1) bar_sync(kBarXDone + ((s - 1) & 3), kWaveWorkers);
            // ---- phase alpha: A_j(p = s - 2j) = H-odd(j) + V-even(j) on rows (2p, 2p+1) x columns (2k+1, 2k+2) ----------
            {
                bool v[kItems];
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    const int p = s - 2 * jj[i];
                    v[i] = on[i] && 2 * p >= rloA[i] && 2 * p < rhiA[i];  // warp-uniform (rhi is even with a bottom halo)
                }
                if (kWaveBoth && v[0] && v[1]) {
                    phase_alpha(two_t{}, 0, s);
                } else {
                    if (v[0]) phase_alpha(one_t{}, 0, s);
                    if (v[1]) phase_alpha(one_t{}, 1, s);
                }
            }
            if (kWaveSets == 2) {
                // B_j only needs A_j of the same stage: the alpha barrier spans the two half-band groups that share
                // their stages (4 warps, one per scheduler), not the whole set
                bar_sync(set ? 14 + ((grp >> 1) & 1) : ((grp & 2) ? 13 : kBarX), kWaveSetThreads / 2);  // ids 2, 13 (X), 14, 15 (Y)
                if (kWaveStagger && set == 0 && s >= 1) bar_arrive(kBarXDone + ((s - 1) & 3), kWaveWorkers);
            } else {
                __syncthreads();
            }
            // ---- phase beta: B_j(p = s - 2j - 1) = V-odd(j) + H-even(j+1) on rows (2p+1, 2p+2) x columns (2k, 2k+1) -------
            {
                // V-odd(j) needs both rows inside sweep j's range; below it the rows are stale for every later pass too.
                // Without a bottom halo the pair whose lower row does not exist still gives its upper row the H-even pass.
                bool v[kItems];
#pragma unroll
                for (int i = 0; i < kItems; ++i) {
                    const int p = s - 2 * jj[i] - 1;
                    v[i] = on[i] && p >= 0 && 2 * p + 1 >= rloB[i] && 2 * p + 2 < rhiB[i];  // warp-uniform
                }
                if (kWaveBoth && v[0] && v[1]) {
                    phase_beta(two_t{}, 0, s);
                } else {
                    if (v[0]) phase_beta(one_t{}, 0, s);
                    if (v[1]) phase_beta(one_t{}, 1, s);
                }
            }
            if (kWaveSets == 2) {
                if (set) bar_sync(kBarY, kWaveSetThreads);
                else bar_sync(kBarXEnd, kWaveSetThreads + kWaveEntry);  // with the entry warps: the next pair is in the ring
                if (set == 0) {
                    if (!kWaveStagger && s + 1 < n_steps) bar_arrive(kBarXDone + (s & 3), kWaveWorkers);  // X has finished step s
                } else if (s + 2 + kWaveSlack < n_steps) {
                    bar_arrive(kBarYDone + (s & 3), kWaveSetThreads + kWaveEntry);  // Y has finished step s: its slots may be refilled
                }
            } else {
                __syncthreads();
            }
#pragma unroll
            for (int i = 0; i < kItems; ++i) e2[i] = e2[i] + 2 == RB ? 0 : e2[i] + 2;
        }
    }

    __syncthreads();  // the ring and the staging slots are reused by the next chunk
    if (tid == kWaveWorkers) {
        if (st_up) ss.done(0, (unsigned)(min(rb, peer_up.row_hi) - max(ra, peer_up.row_lo)));
        if (st_dn) ss.done(1, (unsigned)(min(rb, peer_dn.row_hi) - max(ra, peer_dn.row_lo)));
    }
    }  // chunks of this CTA

    // ---- counters: a stick live when its block state was derived is visited K times unless it tears on the way (then
    // ---- sweep + 1 times, in acc_extra); sticks torn by the entry's H-even pass were visited once -------------------------
    // (per thread the difference may be negative — derivation and tears of one block happen in different threads — so
    //  the sum is formed in signed 64-bit; the grand total is non-negative)
    events.flush(ctr);
    int le = (int)acc_live - (int)acc_torn_post;  // live sticks left (signed per thread, see above)
    long long v = (long long)le * K + (long long)acc_extra;
    unsigned int t = acc_torn;
    for (int o = 16; o > 0; o >>= 1) {
        v += __shfl_down_sync(0xffffffffu, v, o);
        t += __shfl_down_sync(0xffffffffu, t, o);
        le += __shfl_down_sync(0xffffffffu, le, o);
    }
    if ((tid & 31) == 0) {
        if (le) atomicAdd(&ctr->live_last, (unsigned long long)(long long)le);  // (two's complement: negative partial sums cancel)
        if (v) atomicAdd(&ctr->relaxations, (unsigned long long)v);
        if (t) {
            atomicAdd(&ctr->torn_total, (unsigned long long)t);
            atomicAdd(&ctr->torn_last, (unsigned long long)t);
        }
    }
}

}  // namespace clothgpu_impl
