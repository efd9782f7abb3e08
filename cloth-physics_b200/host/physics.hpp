// physics.hpp — header-only C++ mirror of the reference's Go package `physics` over the C ABI of libclothgpu.so.
//
// The reference is Go; this image has no Go toolchain, so the compiled host-side mirror is C++ (the cgo shim a
// maintainer would use is under host/go/, source only).  Same names, argument meaning and error behaviour:
//
//   physics::NewCloth(width, height, spacing)      physics/cloth.go:31   (colour is a render attribute: omitted)
//   Cloth::Init(posX, posY, hud)                    physics/cloth.go:42   idempotent once initialised
//   Cloth::Update(win, mouse, hud, dt)              physics/cloth.go:85   physics half; the render half is Segments()
//   Cloth::Reset(startX, startY, hud)               physics/cloth.go:142
//   Cloth::Width / Cloth::Height                    physics/cloth.go:21-22 (public, mutable, used by the next Init/Reset)
//   physics::Mouse + its 20 accessors               physics/mouse.go:9-99
//   gui::Hud (sliders + WinOffsetX/Y only)          gui/hud.go:32-97
//
// Errors: the Go code has none besides a panic (integer divide by zero in Init when Width/spacing < 10,
// cloth.go:72).  The mirror throws std::runtime_error carrying clothgpu_last_error() in the same situations and
// when no B200 is present — there is no CPU fallback.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/clothgpu.h"

namespace gui {

// gui/hud.go:32-38
enum HudSliderType { HudSliderDragForce = 0, HudSliderGravityForce, HudSliderStiffness, HudSliderFriction, HudSliderTearDistance };

struct Slider {  // gui/hud.go:65-71 (Widget.Value is what particle.update reads)
    float Value, Min, Max;
};

struct Hud {
    Slider Sliders[5];
    double WinOffsetX = 0, WinOffsetY = 0;  // gui/hud.go:47-48
    Hud() {                                 // NewHud, gui/hud.go:86-97 (row i of the table feeds enum value i)
        Sliders[HudSliderDragForce] = {2.0f, 1.1f, 15.0f};
        Sliders[HudSliderGravityForce] = {250.0f, 100.0f, 500.0f};
        Sliders[HudSliderStiffness] = {30.0f, 10.0f, 50.0f};
        Sliders[HudSliderFriction] = {0.98f, 0.95f, 0.99f};
        Sliders[HudSliderTearDistance] = {15.0f, 5.0f, 50.0f};
    }
};

}  // namespace gui

namespace physics {

struct Window {  // gtx.Constraints.Max, physics/particle.go:102
    int X, Y;
};

// physics/mouse.go:9-99
class Mouse {
  public:
    void UpdatePosition(double x, double y) { px_ = x_, py_ = y_, x_ = x, y_ = y; }
    void SetLeftButton() { left_ = true; }
    bool GetLeftButton() const { return left_; }
    void ReleaseLeftButton() { left_ = false; }
    void SetRightButton() { right_ = true; }
    bool GetRightButton() const { return right_; }
    void ReleaseRightButton() { right_ = false; }
    void SetDragging(bool d) { dragging_ = d; }
    bool GetDragging() const { return dragging_; }
    void SetCtrlDown(bool s) { ctrl_ = s; }
    bool GetCtrlDown() const { return ctrl_; }
    void SetForce(double f) { force_ = f; }
    double GetForce() const { return force_; }
    void ResetForce() { force_ = 0; }
    void SetScrollY(float s) { scroll_ = s; }
    float GetScrollY() const { return scroll_; }
    void SetMaxScrollY(float s) { max_scroll_ = s; }
    float GetMaxScrollY() const { return max_scroll_; }
    double X() const { return x_; }
    double Y() const { return y_; }

    void Fill(clothgpu_frame& f) const {
        f.mouse_x = x_, f.mouse_y = y_, f.mouse_px = px_, f.mouse_py = py_;
        f.mouse_force = force_;
        f.scroll_y = scroll_;
        f.buttons = (left_ ? (uint32_t)CLOTHGPU_MOUSE_LEFT : 0u) | (right_ ? (uint32_t)CLOTHGPU_MOUSE_RIGHT : 0u) |
                    (dragging_ ? (uint32_t)CLOTHGPU_MOUSE_DRAGGING : 0u) | (ctrl_ ? (uint32_t)CLOTHGPU_MOUSE_CTRL : 0u);
    }

  private:
    double x_ = 0, y_ = 0, px_ = 0, py_ = 0, force_ = 0;
    float scroll_ = 0, max_scroll_ = 0;
    bool left_ = false, right_ = false, dragging_ = false, ctrl_ = false;
};

struct Segment {  // one live stick as the render loops see it, physics/cloth.go:108-114
    float x1, y1, x2, y2;
};

class Cloth {
  public:
    int Width, Height;  // physics/cloth.go:21-22

    Cloth(int width, int height, int spacing) : Width(width), Height(height), spacing_(spacing) {}
    Cloth(const Cloth&) = delete;
    Cloth& operator=(const Cloth&) = delete;
    ~Cloth() { Close(); }

    // physics/cloth.go:42-81
    void Init(int posX, int posY, const gui::Hud&) {
        if (gpu_) return;  // "Skip the cloth initialization when the window is resized."
        clothgpu_desc d;
        clothgpu_desc_default(&d);
        d.width = Width, d.height = Height, d.spacing = spacing_;
        d.pos_x = posX, d.pos_y = posY;
        d.pin_rule = CLOTHGPU_PIN_REFERENCE;
        Check(clothgpu_create(&d, &gpu_), "Init");
    }

    // physics/cloth.go:142-148
    void Reset(int startX, int startY, const gui::Hud& hud) {
        clothgpu_info info;
        if (gpu_ && clothgpu_get_info(gpu_, &info) == CLOTHGPU_OK && info.nx == Width / spacing_ + 1 &&
            info.ny == Height / spacing_ + 1) {
            Check(clothgpu_reset(gpu_, startX, startY), "Reset");
            return;
        }
        Close();  // Width/Height changed (main.go:190-191): a different grid
        Init(startX, startY, hud);
    }

    // physics half of physics/cloth.go:85-100
    void Update(Window win, const Mouse& mouse, const gui::Hud& hud, double dt) {
        clothgpu_frame f;
        clothgpu_frame_default(&f);
        for (int i = 0; i < 5; ++i) f.slider[i] = hud.Sliders[i].Value;  // physics/particle.go:78-83
        f.drag_force_max = hud.Sliders[gui::HudSliderDragForce].Max;    // physics/particle.go:97
        mouse.Fill(f);
        f.win_w = win.X, f.win_h = win.Y;
        f.win_off_x = hud.WinOffsetX, f.win_off_y = hud.WinOffsetY;
        f.dt = dt;
        f.iterations = Iterations;
        Check(clothgpu_step(gpu_, &f), "Update");
    }

    // render half, physics/cloth.go:108-114 and 126-134: live sticks in creation order; highlighted[i] != 0 when both
    // endpoints are highlighted
    size_t Segments(std::vector<Segment>& out, std::vector<uint8_t>& highlighted) {
        clothgpu_info info;
        Check(clothgpu_get_info(gpu_, &info), "Segments");
        out.resize((size_t)info.sticks);
        highlighted.resize((size_t)info.sticks);
        size_t n = 0;
        Check(clothgpu_read_segments_f32(gpu_, 0, &out[0].x1, highlighted.data(), nullptr, out.size(), &n), "Segments");
        out.resize(n);
        highlighted.resize(n);
        return n;
    }

    // the same loops one step further: addSegment / normal (physics/cloth.go:150-168) applied on the device; out gets 8
    // floats per live stick {a+n, b+n, b-n, a-n}
    size_t Quads(std::vector<float>& out, std::vector<uint8_t>& highlighted, float lineWidth = 0.6f) {
        clothgpu_info info;
        Check(clothgpu_get_info(gpu_, &info), "Quads");
        out.resize((size_t)info.sticks * 8);
        highlighted.resize((size_t)info.sticks);
        size_t n = 0;
        Check(clothgpu_read_quads_f32(gpu_, 0, lineWidth, out.data(), highlighted.data(), (size_t)info.sticks, &n), "Quads");
        out.resize(n * 8);
        highlighted.resize(n);
        return n;
    }

    // Quads into page-locked memory the handle keeps (clothgpu_host_alloc): for GUI-sized cloths the export kernel writes
    // the list there itself — one launch, one synchronisation, no copies.  The pointers stay valid until Close().
    size_t QuadsPinned(const float*& quads, const uint8_t*& highlighted, float lineWidth = 0.6f) {
        clothgpu_info info;
        Check(clothgpu_get_info(gpu_, &info), "QuadsPinned");
        const size_t cap = (size_t)info.sticks;
        if (cap > pinned_cap_) {
            FreePinned();
            Check(clothgpu_host_alloc(cap * 32, &pinned_quads_), "QuadsPinned");
            Check(clothgpu_host_alloc(cap, &pinned_hl_), "QuadsPinned");
            pinned_cap_ = cap;
        }
        size_t n = 0;
        if (cap) Check(clothgpu_read_quads_f32(gpu_, 0, lineWidth, (float*)pinned_quads_, (uint8_t*)pinned_hl_, cap, &n), "QuadsPinned");
        quads = (const float*)pinned_quads_, highlighted = (const uint8_t*)pinned_hl_;
        return n;
    }

    clothgpu_counters Counters() {
        clothgpu_counters c;
        Check(clothgpu_get_counters(gpu_, &c), "Counters");
        return c;
    }

    void Close() {
        if (gpu_) clothgpu_destroy(gpu_);
        gpu_ = nullptr;
        FreePinned();
    }

    clothgpu* Handle() { return gpu_; }
    int Iterations = 1;  // extension: relaxation sweeps per frame (the reference does 1)

  private:
    void FreePinned() {
        clothgpu_host_free(pinned_quads_);
        clothgpu_host_free(pinned_hl_);
        pinned_quads_ = pinned_hl_ = nullptr;
        pinned_cap_ = 0;
    }
    void *pinned_quads_ = nullptr, *pinned_hl_ = nullptr;
    size_t pinned_cap_ = 0;
    static void Check(int rc, const char* what) {
        if (rc != CLOTHGPU_OK) throw std::runtime_error(std::string(what) + ": " + clothgpu_last_error());
    }
    int spacing_;
    clothgpu* gpu_ = nullptr;
};

inline Cloth* NewCloth(int width, int height, int spacing) { return new Cloth(width, height, spacing); }

}  // namespace physics
