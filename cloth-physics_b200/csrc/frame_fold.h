// frame_fold.h — host side: fold one clothgpu_frame (Mouse + HUD sliders + window, as (*Cloth).Update
// receives them, physics/cloth.go:85) into the frame-uniform block the kernels read.
//
// Everything here is evaluated in the reference's operation order and type (float32 slider values
// widened to REAL at the point of use, physics/particle.go:78-83), so the folded values are the ones
// every particle would have computed for itself.
#pragma once
#include <algorithm>
#include <cmath>
#include <limits>

#include "cloth_types.cuh"

namespace clothgpu_impl {

// Smallest s >= 0 with RN(sqrt(s)) >= r, so that for every s >= 0:  (sqrt(s) < r) == (s < result).
// RN∘sqrt is monotone non-decreasing, hence the set {s : sqrt(s) < r} is a down-set and the
// comparison against the rounded square root can be made on s itself without computing the root.
// Uses the host's correctly rounded IEEE sqrt.  NaN s compares false on both sides.
template <typename T>
inline T sqrt_lt_threshold(T r) {
    if (!(r > T(0))) return T(0);  // r <= 0 or NaN: never true
    if (std::isinf(r)) return r;   // true for every finite s
    const T inf = std::numeric_limits<T>::infinity();
    T s = r * r;
    for (;;) {  // walk down while the predecessor still reaches r
        if (!(s > T(0))) break;
        T p = std::nextafter(s, T(0));
        if (std::sqrt(p) >= r)
            s = p;
        else
            break;
    }
    while (std::sqrt(s) < r) s = std::nextafter(s, inf);  // walk up to the first s that reaches r
    return s;
}

template <typename T>
inline FrameU<T> fold_frame(const clothgpu_frame& f, int spacing) {
    FrameU<T> u;
    // physics/particle.go:78-83
    T dragForce = (T)f.slider[CLOTHGPU_SLIDER_DRAG_FORCE];
    const T gravityForce = (T)f.slider[CLOTHGPU_SLIDER_GRAVITY];
    const T stiffness = (T)f.slider[CLOTHGPU_SLIDER_STIFFNESS];
    const T friction = (T)f.slider[CLOTHGPU_SLIDER_FRICTION];
    const T tearDistance = (T)f.slider[CLOTHGPU_SLIDER_TEAR_DISTANCE];
    // :96-99 + increaseForce :184-189
    if (f.buttons & CLOTHGPU_MOUSE_LEFT) {
        const T maxDragForce = (T)f.drag_force_max;
        dragForce += (T)f.mouse_force;
        if (dragForce > maxDragForce) dragForce = maxDragForce;
    }
    u.mx = (T)f.mouse_x;
    u.my = (T)f.mouse_y;
    // :109-122, same clamp order
    T mdx = u.mx - (T)f.mouse_px;
    T mdy = u.my - (T)f.mouse_py;
    if (mdx > stiffness) mdx = stiffness;
    if (mdy > stiffness) mdy = stiffness;
    if (mdx < -stiffness) mdx = -stiffness;
    if (mdy < -stiffness) mdy = -stiffness;
    u.push_x = mdx * dragForce;  // :123-124: p.px = p.x - dx*p.dragForce
    u.push_y = mdy * dragForce;
    // :133-138, float32 clamp then widened
    float focusArea = f.scroll_y;
    if (focusArea > 120.0f)
        focusArea = 120.0f;
    else if (focusArea < 30.0f)
        focusArea = 30.0f;
    u.s_drag = sqrt_lt_threshold<T>(tearDistance);
    u.s_pin = sqrt_lt_threshold<T>((T)4);  // consts.ClothPinDist
    u.s_focus = sqrt_lt_threshold<T>((T)focusArea);
    u.s_near = std::max(u.s_drag, std::max(u.s_pin, u.s_focus));
    u.fric = friction;
    // :152-155 with vx = vy = 0 at frame start (:180)
    const T dt = (T)f.dt;
    const T vx = T(0), vy = T(0) + gravityForce;
    u.gx = vx * (dt * dt);
    u.gy = vy * (dt * dt);
    u.W = (T)f.win_w;
    u.H = (T)f.win_h;
    u.off_x = (T)f.win_off_x;
    u.off_y = (T)f.win_off_y;
    u.L = (T)spacing;
    u.s_rest = sqrt_lt_threshold<T>(u.L);
    u.tear_len = (T)f.tear_length;
    u.gain = (T)f.relax_gain;
    u.buttons = f.buttons;
    u.iterations = f.iterations < 1 ? 1 : f.iterations;
    return u;
}

}  // namespace clothgpu_impl
