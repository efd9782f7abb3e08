// physics_selftest — drives the C++ mirror of package physics the way main.go drives the Go package
// (main.go:149-166 geometry, :174-177 force ramp, :260-309 pointer events, :310 Update) for a short scripted session
// and prints counters as one JSON line; tests/test_host_mirror.py compares them with an independent CPU run of the same
// script.  Exit code 3 = no usable GPU (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>

#include "physics.hpp"

int main(int argc, char** argv) {
    const int frames = argc > 1 ? atoi(argv[1]) : 120;
    gui::Hud hud;
    std::unique_ptr<physics::Cloth> cloth(physics::NewCloth(1024, 211, 6));  // main.go:149-158
    physics::Mouse mouse;
    mouse.SetScrollY(50);      // main.go:84
    mouse.SetMaxScrollY(120);  // main.go:85
    try {
        cloth->Init(128, 164, hud);  // main.go:160-166
        cloth->Init(128, 164, hud);  // idempotent (main.go:217 calls it again on every resize)
    } catch (const std::exception& e) {
        fprintf(stderr, "%s\n", e.what());
        return strstr(e.what(), "no CUDA device") || strstr(e.what(), "sm_100a") ? 3 : 1;
    }
    const physics::Window win{1280, 820};
    std::vector<physics::Segment> seg;
    std::vector<uint8_t> hl;
    for (int t = 0; t < frames; ++t) {
        if (t == 20) {  // right button down: punch holes along a line (main.go:303-306)
            mouse.SetRightButton();
        } else if (t == 60) {  // release (main.go:286-293)
            mouse.ResetForce(), mouse.ReleaseLeftButton(), mouse.ReleaseRightButton(), mouse.SetDragging(false), mouse.SetCtrlDown(false);
        } else if (t == 70) {  // left press then drag (main.go:279-285, 294-302)
            mouse.SetLeftButton();
            mouse.UpdatePosition(400, 230);
        }
        if (t >= 20 && t < 60) mouse.UpdatePosition(300.0f + 6.0f * (t - 20), 260.0f);
        if (t > 70) {
            mouse.SetForce((t - 70) / 60.0 * 5);  // main.go:174-177
            mouse.SetLeftButton();
            mouse.UpdatePosition(400.0f + 21.0f * (t - 70), 230.0f + ((t / 3) % 2 ? 40.0f : -40.0f));
            mouse.SetDragging(true);
        }
        cloth->Update(win, mouse, hud, 0.022);  // main.go:310, consts.Delta
    }
    const size_t n = cloth->Segments(seg, hl);
    size_t nh = 0;
    for (uint8_t h : hl) nh += h != 0;
    double sum = 0;
    for (const auto& s : seg) sum += (double)s.x1 + s.y1 + s.x2 + s.y2;
    // the outline quads through pageable vectors (copies) and through the handle's page-locked buffers (written by the
    // export kernel itself): the same bytes
    std::vector<float> quads;
    std::vector<uint8_t> qhl;
    const size_t nq = cloth->Quads(quads, qhl);
    const float* pq = nullptr;
    const uint8_t* ph = nullptr;
    const size_t np = cloth->QuadsPinned(pq, ph);
    if (nq != n || np != n || memcmp(pq, quads.data(), n * 32) != 0 || memcmp(ph, qhl.data(), n) != 0 || memcmp(ph, hl.data(), n) != 0) {
        fprintf(stderr, "pinned and pageable quad lists differ\n");
        return 1;
    }
    const clothgpu_counters c = cloth->Counters();
    printf("{\"frames\": %llu, \"live_sticks\": %llu, \"alive_sticks\": %llu, \"inactive\": %llu, \"pinned\": %llu, "
           "\"highlighted\": %llu, \"relaxations\": %llu, \"segments\": %zu, \"highlighted_segments\": %zu, \"segment_sum\": %.17g}\n",
           (unsigned long long)c.frames, (unsigned long long)c.live_sticks, (unsigned long long)c.alive_sticks,
           (unsigned long long)c.inactive, (unsigned long long)c.pinned, (unsigned long long)c.highlighted,
           (unsigned long long)c.relaxations, n, nh, sum);
    cloth->Reset(128, 164, hud);  // main.go:193
    const clothgpu_counters r = cloth->Counters();
    if (r.frames != 0 || r.alive_sticks != 12105 || r.inactive != 0 || r.pinned != 11) {
        fprintf(stderr, "Reset did not restore the Init state\n");
        return 1;
    }
    // the reference panics in Init when Width/spacing < 10 (cloth.go:72)
    physics::Cloth tiny(50, 50, 6);
    try {
        tiny.Init(0, 0, hud);
        fprintf(stderr, "Init of a 9-column cloth did not fail\n");
        return 1;
    } catch (const std::exception&) {
    }
    return 0;
}
