// cloth_small.cuh — GUI-sized cloths: ONE resident CTA steps n frames in ONE launch.
//
// The reference's own cloth is 171 x 36 particles (main.go:149-166): 0.4 MB of state.  One frame of it is a few
// microseconds of arithmetic, so a launch per frame (let alone one per colour pass) is bound by launch latency, and the
// frames of a headless loop (clothgpu_step_n) have no reason to leave the chip in between.  Here the whole cloth lives
// in the shared memory of one CTA — (x, y), (px, py) as pairs and the flag byte: 33 B per particle in binary64, 203 KB for
// the default cloth — for all n frames of the call: load once, per frame the particle loop (cloth.go:92-94) and the
// frame's `iterations` x four colour passes (cloth.go:96-100) separated by block barriers, store once.  The frame
// uniforms come from a device array, one folded FrameU per frame.  Same particle_update / relax_n as every other schedule,
// so the results are bit-identical; any order inside a colour is as good as any other (a colour is a perfect matching).
#pragma once
#include "cloth_fused.cuh"
#include "cloth_kernels.cuh"
#include "cloth_types.cuh"

namespace clothgpu_impl {

constexpr int kSmallThreads = 1024;

template <typename T>
__host__ __device__ constexpr size_t small_smem_bytes(long long particles) {
    return (size_t)particles * (2 * 2 * sizeof(T)) + (size_t)((particles + 15) / 16 * 16);
}

// ctr_last: the counter slot of the LAST frame of the call (cloth_types.cuh, DevCounters; the host advances its slot index
// by n): cumulative fields are added there, its per-frame fields are set from the last frame, the other slot's are cleared.
template <typename T>
__global__ void __launch_bounds__(kSmallThreads, 1)
    small_frames_kernel(Planes<T> P, Geom g, const FrameU<T>* __restrict__ frames, int n_frames, DevCounters* ctr_last) {
    using V2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char small_raw[];
    const int nx = g.nx, ny = g.own_rows, N = nx * ny;
    V2* xy = reinterpret_cast<V2*>(small_raw);
    V2* pxy = xy + N;
    uint8_t* fl = reinterpret_cast<uint8_t*>(pxy + N);
    __shared__ unsigned long long s_acc[8];  // relaxations, torn_total, pinned_new, inactive_new, torn_last, live_last, hl_last
    const int tid = threadIdx.x;
    if (tid < 8) s_acc[tid] = 0ull;

    for (int i = tid; i < N; i += kSmallThreads) {  // ---- load once
        const int r = i / nx, c = i - r * nx;
        const long long at = (long long)r * g.pitch + c;
        V2 a, b;
        a.x = P.x[at], a.y = P.y[at], b.x = P.px[at], b.y = P.py[at];
        xy[i] = a, pxy[i] = b, fl[i] = P.fl[at];
    }
    __syncthreads();

    unsigned int visits = 0, torn = 0, pinned = 0, holes = 0;   // whole call
    unsigned int torn_f = 0, live_f = 0, hl_f = 0;              // last frame only
#pragma unroll 1
    for (int f = 0; f < n_frames; ++f) {
        const FrameU<T> u = frames[f];
        const bool dragging = (u.buttons & CLOTHGPU_MOUSE_DRAGGING) != 0;
        const bool last_frame = f + 1 == n_frames;
        // ---- loop over particles, cloth.go:92-94 ------------------------------------------------------------------------
        for (int i = tid; i < N; i += kSmallThreads) {
            V2 a = xy[i], b = pxy[i];
            uint32_t fb = fl[i];
            const uint32_t before = fb;
            particle_update(a.x, a.y, b.x, b.y, fb, u);
            xy[i] = a, pxy[i] = b, fl[i] = (uint8_t)fb;
            const uint32_t ev = (before ^ fb) & (CLOTHGPU_FLAG_PIN | CLOTHGPU_FLAG_ACTIVE);
            pinned += ev & 1u, holes += (ev >> 1) & 1u;
            if (last_frame) hl_f += (fb >> 2) & 1u;
        }
        __syncthreads();
        // ---- loop over sticks, cloth.go:96-100: `iterations` sweeps of the four colours ---------------------------------
#pragma unroll 1
        for (int it = 0; it < u.iterations; ++it) {
            const bool last_sweep = last_frame && it + 1 == u.iterations;
#pragma unroll 1
            for (int colour = 0; colour < 4; ++colour) {
                const bool vertical = colour >= 2;
                const int odd = colour & 1;
                // horizontal: `per` sticks per row, ny rows; vertical: nx sticks per stick row, `per` such rows
                const int per = vertical ? (ny - odd) / 2 : (nx - odd) / 2;
                const int count = vertical ? per * nx : per * ny;
                const uint32_t alive_bit = vertical ? CLOTHGPU_FLAG_ALIVE_V : CLOTHGPU_FLAG_ALIVE_H;
                for (int s0 = 0; s0 < count; s0 += kSmallThreads) {  // (block-uniform trip count: relax_n votes per warp)
                    const int s = s0 + tid;
                    const bool valid = s < count;
                    int head = 0, tail = 0;
                    if (valid) {
                        if (vertical) {
                            const int j = s / nx, c = s - j * nx;
                            head = (2 * j + odd) * nx + c, tail = head + nx;
                        } else {
                            const int r = s / per, k = s - r * per;
                            head = r * nx + 2 * k + odd, tail = head + 1;
                        }
                    }
                    V2 hv = xy[head], tv = xy[tail];
                    const uint32_t fh = fl[head], ft = fl[tail];
                    const uint32_t live = (valid && (fh & alive_bit) && (fh & ft & CLOTHGPU_FLAG_ACTIVE)) ? 1u : 0u;  // cloth.go:97
                    const uint32_t pins = (fh & CLOTHGPU_FLAG_PIN) | ((ft & CLOTHGPU_FLAG_PIN) << 1);
                    V2* hh[1] = {&hv};
                    V2* tt[1] = {&tv};
                    bool touched = false;
                    const uint32_t tore = relax_n<T, true, 1>(hh, tt, live, pins, u, dragging, touched);
                    if (live) {
                        xy[head] = hv, xy[tail] = tv;
                        if (tore) fl[head] = (uint8_t)(fh & ~alive_bit);  // tear = clear the alive bit (constraint.go:35-39)
                    }
                    visits += live, torn += tore;
                    if (last_frame) torn_f += tore;
                    if (last_sweep) live_f += live - tore;  // visited and not torn: what the next frame starts with
                }
                __syncthreads();
            }
        }
    }

    for (int i = tid; i < N; i += kSmallThreads) {  // ---- store once (in place: nobody else reads these planes)
        const int r = i / nx, c = i - r * nx;
        const long long at = (long long)r * g.pitch + c;
        const V2 a = xy[i], b = pxy[i];
        P.x[at] = a.x, P.y[at] = a.y, P.px[at] = b.x, P.py[at] = b.y, P.fl[at] = fl[i];
    }
    // ---- counters ---------------------------------------------------------------------------------------------------------
    unsigned int v[7] = {visits, torn, pinned, holes, torn_f, live_f, hl_f};
#pragma unroll
    for (int k = 0; k < 7; ++k) {
        for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_down_sync(0xffffffffu, v[k], o);
        if ((tid & 31) == 0 && v[k]) atomicAdd(&s_acc[k], (unsigned long long)v[k]);
    }
    __syncthreads();
    if (tid == 0) {
        DevCounters* other = reinterpret_cast<DevCounters*>(reinterpret_cast<unsigned long long>(ctr_last) ^ 64ull);
        ctr_last->relaxations += s_acc[0];
        ctr_last->torn_total += s_acc[1];
        ctr_last->pinned_new += s_acc[2];
        ctr_last->inactive_new += s_acc[3];
        ctr_last->torn_last = s_acc[4];
        ctr_last->live_last = s_acc[5];
        ctr_last->hl_last = s_acc[6];
        other->torn_last = other->live_last = other->hl_last = 0ull;
    }
}

}  // namespace clothgpu_impl
