// cloth_stream1.cuh — the reference's own frame (ONE relaxation sweep, physics/cloth.go:96-100) as a single streaming
// pass over HBM: no shared memory, no block barriers, every warp independent.
//
// With one sweep per frame the launch is bound by HBM, not by arithmetic, so the job is to keep loads in flight.
// A warp owns a window of 64 columns (one (even, odd) column pair per lane, 128-bit accesses, 512 contiguous bytes per
// warp and plane) and marches down a chunk of rows two at a time, holding a sliding window of three rows in registers:
//
//     load rows (r, r+1)  ->  particle_update (cloth.go:92-94; previous positions go straight back to HBM)
//       H-even   the lane's own pair                                    in registers
//       H-odd    odd column of lane i  <->  even column of lane i+1     two warp shuffles each way
//       V-even   rows (r, r+1)                                          in registers
//       V-odd    rows (r-1, r), r-1 carried from the previous trip      in registers
//     store rows (r-1, r), carry row r+1
//
// so every particle is read once and written once.  The colour order H-even, H-odd, V-even, V-odd of the whole-cloth
// pass is respected because each stick only needs its two endpoints to have finished the previous colours, which the
// row pipeline guarantees (V-even of a pair after both H passes of both rows, V-odd after the V-even of both pairs it
// joins).  Column halo: lane 0's even column misses its H-odd stick to the left and lane 31's odd column the one to
// the right, so windows advance by 60 columns and lanes 0 / 31 are recomputed by the neighbouring warp (except at the
// cloth's edge, where those sticks do not exist).  Row halo: a chunk reads two rows above and below its owned rows.
// Loads run two trips ahead of the arithmetic: every lane owns a small ring of shared-memory slots that cp.async fills
// (a lane only reads back what it copied itself, so this needs no barrier and no registers for the data in flight).
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
constexpr int kS1WindowStep = 60;     // columns a window advances by (30 owned pairs + 1 halo pair per side)

struct Stream1Plan {
    int windows;     // column windows per cloth
    int chunk_rows;  // owned rows per chunk (even)
    int chunks;      // row chunks per cloth (strip)
    long long tasks; // windows * chunks * ensemble
};

template <typename T>
struct RowRaw {  // one row of a lane's column pair, as loaded
    typename Vec2<T>::type x, y, px, py;
    uint32_t fl;  // low byte = even column, next byte = odd column
};

// RING == 0: the next trip's loads wait in registers.  RING >= 2: every lane owns RING slots of 8 x 16 B in shared memory
// that cp.async fills RING trips ahead (a lane only ever reads back what it copied itself, so no barrier is involved):
// more loads in flight, 34 fewer live registers.
template <typename T, bool PER_CLOTH, bool DRAG, int RING>
__global__ void __launch_bounds__(kS1Threads, RING ? S1_RING_CTAS : S1_MIN_CTAS)
    stream1_frame_kernel(Planes<T> in, Planes<T> out, Geom g, Stream1Plan pl, const FrameU<T> u0,
                         const FrameU<T>* __restrict__ per_cloth, DevCounters* ctr, const PeerOut<T> peer_up,
                         const PeerOut<T> peer_dn, const StripSync ss) {
    using V2 = typename Vec2<T>::type;
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

    const int c0 = w * kS1WindowStep;
    const int ce = c0 + 2 * lane;  // my even column; the odd one is ce + 1
    const bool ex_e = ce < g.nx, ex_o = ce + 1 < g.nx;
    // pairs whose two columns come out right in this window (see the header): they are stored by this warp
    const bool own_cols = ex_e && (lane >= 1 || w == 0) && (lane <= 30 || c0 + 64 >= g.nx);
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
        r.x.x = r.x.y = r.y.x = r.y.y = r.px.x = r.px.y = r.py.x = r.py.y = T(0);
        r.fl = 0;
        if (gy < re && ex_e) {
            const long long at = cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce;
            r.x = *reinterpret_cast<const V2*>(in.x + at);
            r.y = *reinterpret_cast<const V2*>(in.y + at);
            r.px = *reinterpret_cast<const V2*>(in.px + at);
            r.py = *reinterpret_cast<const V2*>(in.py + at);
            r.fl = *reinterpret_cast<const unsigned short*>(in.fl + at);
            if (!ex_o) r.fl &= 0xffu;  // column nx (row padding) does not exist
        }
        return r;
    };
    auto store_xy = [&](int gy, const V2& pe, const V2& po, uint32_t fe, uint32_t fo) {
        if (!own_cols || gy < ra || gy >= rb) return;
        const long long at = cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce;
        V2 ox2, oy2;
        ox2.x = pe.x, ox2.y = po.x, oy2.x = pe.y, oy2.y = po.y;
        const unsigned short ff = (unsigned short)((fe & 0xffu) | ((fo & 0xffu) << 8));
#if S1_STCS  // streaming stores: the planes are far larger than L2 and are read again only by the NEXT frame
        __stcs(reinterpret_cast<V2*>(out.x + at), ox2);
        __stcs(reinterpret_cast<V2*>(out.y + at), oy2);
        __stcs(reinterpret_cast<unsigned short*>(out.fl + at), (unsigned short)ff);
#else
        *reinterpret_cast<V2*>(out.x + at) = ox2;
        *reinterpret_cast<V2*>(out.y + at) = oy2;
        *reinterpret_cast<unsigned short*>(out.fl + at) = ff;
#endif
        if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {  // the neighbouring strip's ghost copy of this row
            const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + ce;
            *reinterpret_cast<V2*>(peer_up.x + pat) = ox2;
            *reinterpret_cast<V2*>(peer_up.y + pat) = oy2;
            *reinterpret_cast<unsigned short*>(peer_up.fl + pat) = ff;
        }
        if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
            const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + ce;
            *reinterpret_cast<V2*>(peer_dn.x + pat) = ox2;
            *reinterpret_cast<V2*>(peer_dn.y + pat) = oy2;
            *reinterpret_cast<unsigned short*>(peer_dn.fl + pat) = ff;
        }
    };
    auto store_prev = [&](int gy, const V2& vpx, const V2& vpy) {
        if (!own_cols || gy < ra || gy >= rb) return;
        const long long at = cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce;
#if S1_STCS
        __stcs(reinterpret_cast<V2*>(out.px + at), vpx);
        __stcs(reinterpret_cast<V2*>(out.py + at), vpy);
#else
        *reinterpret_cast<V2*>(out.px + at) = vpx;
        *reinterpret_cast<V2*>(out.py + at) = vpy;
#endif
        if (gy >= peer_up.row_lo && gy < peer_up.row_hi) {
            const long long pat = (long long)e * peer_up.cloth_stride + (long long)(gy - peer_up.held_row0) * g.pitch + ce;
            *reinterpret_cast<V2*>(peer_up.px + pat) = vpx;
            *reinterpret_cast<V2*>(peer_up.py + pat) = vpy;
        }
        if (gy >= peer_dn.row_lo && gy < peer_dn.row_hi) {
            const long long pat = (long long)e * peer_dn.cloth_stride + (long long)(gy - peer_dn.held_row0) * g.pitch + ce;
            *reinterpret_cast<V2*>(peer_dn.px + pat) = vpx;
            *reinterpret_cast<V2*>(peer_dn.py + pat) = vpy;
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
    V2 cE, cO;
    cE.x = cE.y = cO.x = cO.y = T(0);
    uint32_t cfe = 0, cfo = 0;

    // ---- prefetch machinery -------------------------------------------------------------------------------------
    extern __shared__ __align__(16) unsigned char s1_smem[];
    // slot (stage, q, plane) of this lane: 32 lanes x sizeof(V2) contiguous, conflict-free
    auto slot = [&](int stage, int q, int plane) -> V2* {
        return reinterpret_cast<V2*>(s1_smem) + ((((threadIdx.x >> 5) * (RING ? RING : 1) + stage) * 8 + q * 4 + plane) * 32 + lane);
    };
    auto issue_trip = [&](int r, int stage) {  // rows r, r+1 -> stage (RING > 0 only)
#pragma unroll
        for (int q = 0; q < 2; ++q) {
            const int gy = r + q;
            const bool ok = gy < re && ex_e;
            const long long at = ok ? cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce : 0;
            cp_async_lane<sizeof(V2)>(slot(stage, q, 0), in.x + at, ok);
            cp_async_lane<sizeof(V2)>(slot(stage, q, 1), in.y + at, ok);
            cp_async_lane<sizeof(V2)>(slot(stage, q, 2), in.px + at, ok);
            cp_async_lane<sizeof(V2)>(slot(stage, q, 3), in.py + at, ok);
        }
        cp_async_commit();
    };
    // (the value is used RING trips later; nothing may touch it before — masking the non-existent odd column right here
    //  made every warp wait for the load at once: 17 % of all stall samples of the launch sat on that one instruction)
    auto load_flags = [&](int gy) -> uint32_t {
        uint32_t f = 0;
        if (gy < re && ex_e) f = *reinterpret_cast<const unsigned short*>(in.fl + cloth_base + (long long)(gy - g.held_row0) * g.pitch + ce);
        return f;
    };
    const uint32_t fl_keep = ex_o ? 0xffffu : 0xffu;  // column nx (row padding) does not exist

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
#pragma unroll 1
    for (int r = rs; r < re; r += 2) {
        RowRaw<T> raw0, raw1;
        if (RING == 0) {
            // (a ping-pong of two register sets instead of these copies measured slower: twice the code, worse schedule)
            raw0 = nxt0, raw1 = nxt1;
            nxt0 = load_row(r + 2);  // in flight during this trip's arithmetic (rows >= re load nothing)
            nxt1 = load_row(r + 3);
        } else {
            cp_async_wait<RING - 1>();  // the oldest group (this trip) has landed
            raw0.x = *slot(stage, 0, 0), raw0.y = *slot(stage, 0, 1), raw0.px = *slot(stage, 0, 2), raw0.py = *slot(stage, 0, 3);
            raw1.x = *slot(stage, 1, 0), raw1.y = *slot(stage, 1, 1), raw1.px = *slot(stage, 1, 2), raw1.py = *slot(stage, 1, 3);
            raw0.fl = fq[0] & fl_keep, raw1.fl = fq[1] & fl_keep;
#pragma unroll
            for (int k = 0; k + 1 < RING; ++k) fq[2 * k] = fq[2 * k + 2], fq[2 * k + 1] = fq[2 * k + 3];
            issue_trip(r + 2 * RING, stage);  // refill the slots just read (always commits, so group counting stays uniform)
            fq[2 * RING - 2] = load_flags(r + 2 * RING), fq[2 * RING - 1] = load_flags(r + 2 * RING + 1);
            stage = stage + 1 == RING ? 0 : stage + 1;
        }
        V2 E[2], O[2];
        uint32_t fe[2], fo[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {  // ---- loop over particles, cloth.go:92-94
            const RowRaw<T>& rw = q ? raw1 : raw0;
            E[q].x = rw.x.x, O[q].x = rw.x.y, E[q].y = rw.y.x, O[q].y = rw.y.y;
            fe[q] = rw.fl & 0xffu, fo[q] = rw.fl >> 8;
            T px0 = rw.px.x, px1 = rw.px.y, py0 = rw.py.x, py1 = rw.py.y;
            if (r + q < re && ex_e) {
                const uint32_t fe_in = fe[q], fo_in = fo[q];
                particle_update(E[q].x, E[q].y, px0, py0, fe[q], u);
                if (ex_o) particle_update(O[q].x, O[q].y, px1, py1, fo[q], u);
                if (own_cols && r + q >= ra && r + q < rb) events.note2(fe_in, fe[q], fo_in, fo[q]);
                V2 vpx, vpy;
                vpx.x = px0, vpx.y = px1, vpy.x = py0, vpy.y = py1;
                store_prev(r + q, vpx, vpy);
            }
        }
        const bool own_r[2] = {r >= ra && r < rb, r + 1 >= ra && r + 1 < rb};
        bool touched = false;
#if defined(S1_NO_MATH)  // timing experiment only: what the memory system gives this access pattern without the sticks
        if (u.iterations == 12345)
#endif
        {
        {  // ---- H-even: the lane's own pair, rows r and r+1
            V2* hh[2] = {&E[0], &E[1]};
            V2* tt[2] = {&O[0], &O[1]};
            uint32_t live = 0, pins = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                live |= live_of(fe[q], fo[q], CLOTHGPU_FLAG_ALIVE_H) << q;
                pins |= ((fe[q] & CLOTHGPU_FLAG_PIN) | ((fo[q] & CLOTHGPU_FLAG_PIN) << 1)) << (2 * q);
            }
            const uint32_t tore = relax_n<T, DRAG, 2>(hh, tt, live, pins, u, dragging, touched);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (own_cols && own_r[q]) visits += (live >> q) & 1u, torn += (tore >> q) & 1u;
                if (DRAG && (tore & (1u << q))) fe[q] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_H;
            }
        }
        {  // ---- H-odd: my odd column (head) and the even column of lane + 1 (tail)
            V2 Tn[2];
            uint32_t fn[2];
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                Tn[q] = shfl_v2(E[q], true);
                fn[q] = __shfl_down_sync(FULL, fe[q], 1);
            }
            V2* hh[2] = {&O[0], &O[1]};
            V2* tt[2] = {&Tn[0], &Tn[1]};
            uint32_t live = 0, pins = 0;
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (lane < 31) live |= live_of(fo[q], fn[q], CLOTHGPU_FLAG_ALIVE_H) << q;
                pins |= ((fo[q] & CLOTHGPU_FLAG_PIN) | ((fn[q] & CLOTHGPU_FLAG_PIN) << 1)) << (2 * q);
            }
            const uint32_t tore = relax_n<T, DRAG, 2>(hh, tt, live, pins, u, dragging, touched);
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (own_cols && own_r[q]) visits += (live >> q) & 1u, torn += (tore >> q) & 1u;
                if (DRAG && (tore & (1u << q))) fo[q] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_H;
                const V2 back = shfl_v2(Tn[q], false);  // the relaxed tail returns to its lane
                if (lane >= 1) E[q] = back;
            }
        }
        {  // ---- V-even: rows (r, r+1), both columns
            V2* hh[2] = {&E[0], &O[0]};
            V2* tt[2] = {&E[1], &O[1]};
            const uint32_t live = live_of(fe[0], fe[1], CLOTHGPU_FLAG_ALIVE_V) | (live_of(fo[0], fo[1], CLOTHGPU_FLAG_ALIVE_V) << 1);
            const uint32_t pins = (fe[0] & CLOTHGPU_FLAG_PIN) | ((fe[1] & CLOTHGPU_FLAG_PIN) << 1) |
                                  ((fo[0] & CLOTHGPU_FLAG_PIN) << 2) | ((fo[1] & CLOTHGPU_FLAG_PIN) << 3);
            const uint32_t tore = relax_n<T, DRAG, 2>(hh, tt, live, pins, u, dragging, touched);
            if (own_cols && own_r[0]) visits += __popc(live), torn += __popc(tore);
            if (DRAG && (tore & 1u)) fe[0] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
            if (DRAG && (tore & 2u)) fo[0] &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
        }
        {  // ---- V-odd: rows (r-1, r); r-1 is the carried row
            V2* hh[2] = {&cE, &cO};
            V2* tt[2] = {&E[0], &O[0]};
            const uint32_t live = live_of(cfe, fe[0], CLOTHGPU_FLAG_ALIVE_V) | (live_of(cfo, fo[0], CLOTHGPU_FLAG_ALIVE_V) << 1);
            const uint32_t pins = (cfe & CLOTHGPU_FLAG_PIN) | ((fe[0] & CLOTHGPU_FLAG_PIN) << 1) |
                                  ((cfo & CLOTHGPU_FLAG_PIN) << 2) | ((fo[0] & CLOTHGPU_FLAG_PIN) << 3);
            const uint32_t tore = relax_n<T, DRAG, 2>(hh, tt, live, pins, u, dragging, touched);
            if (own_cols && r - 1 >= ra && r - 1 < rb) visits += __popc(live), torn += __popc(tore);
            if (DRAG && (tore & 1u)) cfe &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
            if (DRAG && (tore & 2u)) cfo &= ~(uint32_t)CLOTHGPU_FLAG_ALIVE_V;
        }
        }
        store_xy(r - 1, cE, cO, cfe, cfo);  // both rows have now seen all four colours
        store_xy(r, E[0], O[0], fe[0], fo[0]);
        cE = E[1], cO = O[1], cfe = fe[1], cfo = fo[1];
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
