// cloth_types.cuh — device-side data layout and the two per-element update rules.
//
// State lives as structure-of-arrays planes in HBM: x, y, px, py (binary64, or binary32 in the
// optional F32 mode) plus one flag byte per particle (CLOTHGPU_FLAG_* of include/clothgpu.h).
// Rows are `pitch` elements apart (pitch = nx rounded up to 16 elements, so every row starts on a
// 128-byte line and even columns are 16-byte aligned for 128-bit vector access).
//
// The arithmetic below restates physics/particle.go:75-181 and physics/constraint.go:25-54 of the
// reference.  It is compiled with -fmad=false: every product and sum is rounded separately, exactly
// as the Go compiler emits for amd64.  CUDA's binary64 `/` and sqrt() are IEEE round-to-nearest.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/clothgpu.h"

namespace clothgpu_impl {

template <typename T>
struct Planes {
    T* x;
    T* y;
    T* px;
    T* py;
    uint8_t* fl;
};

// Row-strip decomposition over several GPUs: where a neighbouring strip keeps copies ("ghost rows") of some of
// my owned rows.  The planes are the neighbour's buffers this launch must fill (its `out` side), mapped into
// this process over NVLink (cudaIpcOpenMemHandle) or plain pointers when the neighbour lives in this process.
// rows [row_lo, row_hi) are GLOBAL rows; empty when there is no neighbour on that side.
template <typename T>
struct PeerOut {
    T* x;
    T* y;
    T* px;
    T* py;
    uint8_t* fl;
    long long cloth_stride;  // the neighbour's elements between consecutive cloths
    int held_row0;           // global index of the neighbour's first held row
    int row_lo, row_hi;
};

// Geometry of what one handle holds.  "Held" rows = owned rows plus ghost rows (strips only).
struct Geom {
    int nx;            // particles per row of the cloth
    int pitch;         // elements between rows
    int held_row0;     // global index of the first held row
    int held_rows;     // rows held
    int own_row0;      // global index of the first owned row
    int own_rows;      // rows owned
    long long cloth_stride;  // elements between consecutive cloths of an ensemble (per plane)
    int ensemble;
};

// Frame-uniform values, folded once per frame on the host in the reference's operation order
// (frame_fold.h).  The reference re-derives all of them per particle (physics/particle.go:78-83,
// 96-99, 109-122, 133-138, 152-155); they do not depend on the particle.
template <typename T>
struct FrameU {
    T mx, my;          // Mouse.x, Mouse.y
    T push_x, push_y;  // clamp(m - mprev, +-stiffness) * dragForce      particle.go:109-125
    T s_drag;          // dist < tearDistance  <=>  dist^2 sum < s_drag  (exact, see frame_fold.h)
    T s_pin;           // dist < ClothPinDist(4)
    T s_focus;         // dist < focusArea
    T s_near;          // max of the three thresholds above: a particle with dist^2 sum >= s_near is out of the mouse's reach
    T fric;            // slider "friction" widened from float32       particle.go:82
    T gx, gy;          // posX, posY = v * (dt*dt)                      particle.go:152-155
    T W, H;            // window size as REAL                           particle.go:164,172
    T off_x, off_y;    // hud.WinOffsetX/Y                              particle.go:89-90
    T L;               // rest length = REAL(spacing)                   cloth.go:63,68
    T s_rest;          // dist < L  <=>  dist^2 sum < s_rest            constraint.go:30
    T tear_len;        // 150                                           constraint.go:36
    T gain;            // 0.35                                          constraint.go:42
    uint32_t buttons;  // CLOTHGPU_MOUSE_*
    int iterations;
};

// Device-side counters accumulated by the frame kernels themselves (warp-reduced, one atomic per warp that saw an
// event), so that clothgpu_get_counters is one small copy instead of a census launch.
// A handle owns TWO of these, adjacent and 128-byte aligned; frame f accumulates into slot f & 1 (the host sums the two
// slots for the cumulative fields and reads the *_last fields of the slot of the last frame).  The first kernel of a frame
// zeroes the *_last fields of the OTHER slot — nobody touches that one until the next frame — so no memset node sits
// between two frame kernels on the stream.
struct alignas(64) DevCounters {
    unsigned long long relaxations;   // stick visits
    unsigned long long torn_total;
    unsigned long long torn_last;     // sticks torn by the frame
    unsigned long long live_last;     // sticks the next frame's loop would relax (cloth.go:97), after the frame's last launch
    unsigned long long hl_last;       // particles highlighted by the frame (particle.go:140-142)
    unsigned long long pinned_new;    // particles pinned by ctrl-click (particle.go:128-130) since the last census
    unsigned long long inactive_new;  // particles deactivated by the right button (particle.go:145-149) since then
    unsigned long long pad_;
};
static_assert(sizeof(DevCounters) == 64, "two slots are told apart by address bit 6");

// Called by every thread of the first kernel of a frame: one thread clears the per-frame fields of the other slot.
__device__ __forceinline__ void prepare_next_frame_counters(DevCounters* ctr) {
    if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
        DevCounters* other = reinterpret_cast<DevCounters*>(reinterpret_cast<unsigned long long>(ctr) ^ 64ull);
        other->torn_last = 0ull;
        other->live_last = 0ull;
        other->hl_last = 0ull;
    }
}

// What the particle loop changed on the particles a thread owns (events are rare: they need the mouse nearby).
struct ParticleEvents {
    unsigned int hl = 0, pinned = 0, holes = 0;
    __device__ __forceinline__ void note(uint32_t before, uint32_t after) {
#ifdef CLOTH_NO_EVENTS  // timing experiment only (counters wrong)
        return;
#endif
        const uint32_t ev = ((before ^ after) & (CLOTHGPU_FLAG_PIN | CLOTHGPU_FLAG_ACTIVE)) | (after & CLOTHGPU_FLAG_HIGHLIGHT);
        if (ev) {
            pinned += ev & 1u;
            holes += (ev >> 1) & 1u;
            hl += (ev >> 2) & 1u;
        }
    }
    // the (even, odd) column pair a lane owns: one combined test on the common path
    __device__ __forceinline__ void note2(uint32_t before0, uint32_t after0, uint32_t before1, uint32_t after1) {
#ifdef CLOTH_NO_EVENTS
        return;
#endif
        const uint32_t any = (((before0 ^ after0) | (before1 ^ after1)) & (CLOTHGPU_FLAG_PIN | CLOTHGPU_FLAG_ACTIVE)) |
                             ((after0 | after1) & CLOTHGPU_FLAG_HIGHLIGHT);
        if (any) {
            note(before0, after0);
            note(before1, after1);
        }
    }
    // all 32 lanes; one atomic per counter and warp, only when something happened
    __device__ __forceinline__ void flush(DevCounters* ctr) {
        if (!__any_sync(0xffffffffu, (hl | pinned | holes) != 0u)) return;
        unsigned int a = hl, b = pinned, c = holes;
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_down_sync(0xffffffffu, a, o);
            b += __shfl_down_sync(0xffffffffu, b, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
        if ((threadIdx.x & 31) == 0) {
            if (a) atomicAdd(&ctr->hl_last, (unsigned long long)a);
            if (b) atomicAdd(&ctr->pinned_new, (unsigned long long)b);
            if (c) atomicAdd(&ctr->inactive_new, (unsigned long long)c);
        }
    }
};

// ---- row strips over several GPUs: the handshake lives inside the frame kernels ------------------------------------
// Every strip owns two 32-bit epoch slots in its own memory: slot 0 is written by its upper neighbour, slot 1 by its
// lower neighbour; epoch k in a slot means "that neighbour has finished the edge work of its operation k-1 ... 0".
// A piece of work of MY operation k (a CTA's tile, a warp's row chunk) that reads ghost rows or stores into a
// neighbour's ghost rows first waits until that neighbour's epoch is >= k (`need`): the neighbour's stores of operation
// k-1 into my ghost rows have landed, and it no longer reads the buffer my stores are about to overwrite.  Pieces that
// touch neither run without waiting, so the interior of frame f+1 overlaps the neighbours' tail of frame f.
// When a piece has stored its share of the rows a neighbour keeps as ghosts it fences (system scope) and adds the
// number of (row, column block) units it covered to a ticket counter in my memory; the piece that completes the
// operation's quota publishes epoch k+1 in the neighbour's slot (st.release.sys).  Every piece that reads ghost rows
// also stores peer rows (the neighbour's ghost depth is at least the halo this launch reads), so "all peer stores
// done" implies "all ghost reads done".
__device__ __forceinline__ unsigned int ld_acquire_sys(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

struct StripSync {
    const unsigned int* my_slots;   // [0] <- upper neighbour, [1] <- lower neighbour; nullptr: not attached
    unsigned int* peer_slot[2];     // where I publish: the upper neighbour's slot 1, the lower neighbour's slot 0
    unsigned long long* ticket;     // [2] in my memory, cumulative units stored towards the upper / lower neighbour
    unsigned long long goal[2];     // value of ticket[side] when this operation's stores towards that side are complete
    unsigned int need;              // epoch the neighbours must have reached (= operations I have completed)

#ifdef CLOTH_NO_STRIPSYNC  // timing experiment only (attached strips unsynchronised)
    __device__ __forceinline__ bool on() const { return false; }
#else
    __device__ __forceinline__ bool on() const { return my_slots != nullptr; }
#endif
    // Called by every thread that is about to read ghost rows of / store into the neighbour on `side` (0 up, 1 down).
    __device__ __forceinline__ void wait(int side) const {
        // epochs only grow; the signed difference tolerates wrap-around
        while ((int)(ld_acquire_sys(my_slots + side) - need) < 0) __nanosleep(64);
    }
    // Called by ONE thread of a piece after all the piece's threads have fenced (__threadfence_system) their peer stores
    // and met at a barrier / __syncwarp: accounts `units` towards `side`, publishes the epoch when the quota is complete.
    __device__ __forceinline__ void done(int side, unsigned int units) const {
        __threadfence_system();
        const unsigned long long before = atomicAdd(ticket + side, (unsigned long long)units);
        if (before + units == goal[side]) {
            __threadfence_system();
            st_release_sys(peer_slot[side], need + 1u);
        }
    }
};

// Asynchronous global -> shared copy of one lane's 16, 8 or 4 bytes; pred == false zero-fills instead of reading.
#ifndef S1_STCS
#define S1_STCS 0
#endif
#ifndef S1_LDHINT
#define S1_LDHINT 0
#endif
template <int BYTES>
__device__ __forceinline__ void cp_async_lane(void* smem, const void* gmem, bool pred) {
    const unsigned s = (unsigned)__cvta_generic_to_shared(smem);
    const int n = pred ? BYTES : 0;
#if S1_LDHINT
    if (BYTES == 16) {
        unsigned long long pol;
        asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
        asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2, %3;" ::"r"(s), "l"(gmem), "r"(n), "l"(pol) : "memory");
        return;
    }
#endif
    if (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
    else if (BYTES == 8)
        asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(s), "l"(gmem), "r"(n) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

// Shared-memory mbarrier and the bulk (TMA, no tensor map) global -> shared copy that completes on it.
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n"
        "MBAR_WAIT:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra MBAR_DONE;\n\t"
        "bra MBAR_WAIT;\n"
        "MBAR_DONE:\n\t}" ::"r"(smem_u32(b)), "r"(parity)
        : "memory");
}
// global -> shared bulk copy (TMA, no tensor map: a contiguous run of bytes), completion counted on an mbarrier
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* b) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(b))
                 : "memory");
}

// ---- (*particle).update, physics/particle.go:75-181, in pieces; `fl` is the particle's flag byte ----
// The window clamp at the end of the rule, :164-178.
template <typename T>
__device__ __forceinline__ void particle_window_clamp(T& x, T& y, T& px, T& py, const FrameU<T>& u) {
    if (x >= u.W) {  // :164-170
        x = u.W;
        px = x;
    } else if (x < T(0)) {
        x = T(0);
        px = x;
    }
    if (y > u.H) {  // :172-178
        y = u.H;
        py = y;
    } else if (y < T(0)) {
        y = T(0);
        py = y;
    }
}

// The pieces of the rule after the pinned test, so that a caller can put votes between them (cloth_stream1.cuh):
// squared distance to the mouse (:104-106; the sqrt is folded into the thresholds s_*) ...
template <typename T>
__device__ __forceinline__ T particle_mouse_dist2(const T x, const T y, const FrameU<T>& u) {
    const T dx = x - u.mx;
    const T dy = y - u.my;
    return dx * dx + dy * dy;
}
// ... what the mouse does to a particle at squared distance s (:108-149; nothing when s >= u.s_near) ...
template <typename T>
__device__ __forceinline__ void particle_mouse(const T x, const T y, T& px, T& py, uint32_t& fl, const T s, const FrameU<T>& u) {
    if ((u.buttons & CLOTHGPU_MOUSE_DRAGGING) && s < u.s_drag) {  // :108-125
        px = x - u.push_x;
        py = y - u.push_y;
    }
    if ((u.buttons & CLOTHGPU_MOUSE_CTRL) && s < u.s_pin) fl |= CLOTHGPU_FLAG_PIN;  // :128-130
    if (s < u.s_focus) {                                                            // :140-149
        fl |= CLOTHGPU_FLAG_HIGHLIGHT;
        if (u.buttons & CLOTHGPU_MOUSE_RIGHT) fl &= ~(uint32_t)CLOTHGPU_FLAG_ACTIVE;
    }
}
// ... and the Verlet step (:151-162) with the window clamp (:164-178).  CLAMP = false leaves the clamp to the caller: the
// return value says whether particle_window_clamp still has to run on this particle (it left the window — an event; a
// caller can test a whole warp's particles with one vote and keep the clamp's selects and copies out of its common path).
template <typename T, bool CLAMP>
__device__ __forceinline__ bool particle_integrate(T& x, T& y, T& px, T& py, const FrameU<T>& u) {
    const T ox = x, oy = y;              // :151
    x = x + (x - px) * u.fric + u.gx;    // :159
    y = y + (y - py) * u.fric + u.gy;    // :160
    px = ox;                             // :162
    py = oy;
    if (!CLAMP) return !(x < u.W) || x < T(0) || !(y <= u.H) || y < T(0);
    particle_window_clamp(x, y, px, py, u);
    return false;
}

// (*particle).update, physics/particle.go:75-181.  CLAMP as in particle_integrate (true: the whole rule, returns false).
// MAY_PIN = false: the caller knows that the particle is not pinned (it has looked at the flag bytes of a whole warp's
// particles at once), so the early return — and the merge of two register sets behind it — is not compiled in.
template <typename T, bool CLAMP = true, bool MAY_PIN = true>
__device__ __forceinline__ bool particle_update(T& x, T& y, T& px, T& py, uint32_t& fl,
                                                const FrameU<T>& u) {
    fl &= ~(uint32_t)CLOTHGPU_FLAG_HIGHLIGHT;  // :76
    if (MAY_PIN && (fl & CLOTHGPU_FLAG_PIN)) {  // :85-92
        x += u.off_x;
        y += u.off_y;
        return false;
    }
    particle_mouse(x, y, px, py, fl, particle_mouse_dist2(x, y, u), u);
    return particle_integrate<T, CLAMP>(x, y, px, py, u);
}

// The common case of particle_update — not pinned, out of the mouse's reach, still inside the window after the step —
// without any of the rule's branches: the caller keeps the inputs, calls this, and when ANY lane of the warp reports
// true re-runs particle_update on the kept inputs of the lanes that do (the mouse touches a handful of particles per
// frame, pins are the cloth's top row, the window clamp an event).  Where it returns false the values written are
// exactly particle_update's: same operations in the same order, none of the branches taken.
template <typename T>
__device__ __forceinline__ bool particle_update_common(T& x, T& y, T& px, T& py, uint32_t& fl, const FrameU<T>& u) {
    const T dx = x - u.mx, dy = y - u.my;
    const T s = dx * dx + dy * dy;
    const T nx = x + (x - px) * u.fric + u.gx;  // :159
    const T ny = y + (y - py) * u.fric + u.gy;  // :160
    const bool rare = (fl & CLOTHGPU_FLAG_PIN) || s < u.s_near || !(nx < u.W) || nx < T(0) || !(ny <= u.H) || ny < T(0);
    px = x, py = y, x = nx, y = ny;
    fl &= ~(uint32_t)CLOTHGPU_FLAG_HIGHLIGHT;
    return rare;
}

// (*constraint).Update, physics/constraint.go:25-54, on a stick known to be live (still in the
// slice, both endpoints active: physics/cloth.go:97).  Returns true when the stick tears; it is still
// relaxed on that visit because the Go code falls through after removeConstraint (:37 -> :41).
template <typename T>
__device__ __forceinline__ bool stick_relax(T& x1, T& y1, T& x2, T& y2, bool pin1, bool pin2,
                                            const FrameU<T>& u) {
    const T dx = x1 - x2;  // :26-28
    const T dy = y1 - y2;
    const T s = dx * dx + dy * dy;
    if (s < u.s_rest) return false;  // :30-32  (dist < length, decided on s exactly)
    const T dist = sqrt(s);
    const bool tore = (u.buttons & CLOTHGPU_MOUSE_DRAGGING) && dist > u.tear_len;  // :35-39
    const T diff = (u.L - dist) / dist;                   // :41
    const T mul = diff * u.gain * (T(1) - u.L / dist);    // :42
    const T ox = dx * mul, oy = dy * mul;                 // :44
    if (!pin1) {                                          // :46-49
        x1 += ox;
        y1 += oy;
    }
    if (!pin2) {  // :50-53
        x2 -= ox;
        y2 -= oy;
    }
    return tore;
}

}  // namespace clothgpu_impl
