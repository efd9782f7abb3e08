// harness — drives package `physics` headless through a recorded mouse trace and dumps CLSTATE1 snapshots; also the
// headless timing arm for BASELINE.md.  It compiles against EITHER implementation of the package:
//
//   - the unmodified reference plus physics/dump_harness.go (copy cloth-physics_b200/host/go/physics/dump_harness.go into
//     a checkout of esimov/cloth-physics, build with `-tags refharness`): regenerates the golden states from the REAL
//     reference — the one thing that would pin the CPU checker the parity tests use (DESIGN.md §6), and
//   - the cgo shim (clothgpu_cgo.go), whose Dump is clothgpu_save_state: the drop-in under the very same driver.
//
// SOURCE ONLY: this image has no Go toolchain (and no Gio module cache), so it has never been built here.  Everything it
// needs is constructible without a window: layout.Context{Ops, Constraints}, gui.NewHud(), the exported Mouse setters.
//
//	go run -tags refharness ./harness -trace tests/golden/c0_trace.jsonl -checkpoints 0,3,99,140,399,449,549,599,799,800,999 -out /tmp/ref
//	python tests/golden/compare_states.py /tmp/ref tests/golden/states      # bit for bit against the CPU checker's dumps
//
// Trace: JSON lines written by `python tests/golden/make_golden.py --dump-trace` (format: clothgpu/trace.py; SURVEY §3.4):
// per frame EVERY Mouse.UpdatePosition argument pair in order, then the four button bools, force and scrollY.
package main

import (
	"bufio"
	"encoding/json"
	"flag"
	"fmt"
	"image"
	"image/color"
	"log"
	"os"
	"strconv"
	"strings"
	"time"

	"gioui.org/layout"
	"gioui.org/op"
	"gioui.org/unit"

	"github.com/esimov/cloth-physics/gui"
	"github.com/esimov/cloth-physics/physics"
)

type record struct {
	Moves      [][2]float64 `json:"moves"`
	Left       bool         `json:"left"`
	Right      bool         `json:"right"`
	Dragging   bool         `json:"dragging"`
	Ctrl       bool         `json:"ctrl"`
	Force      float64      `json:"force"`
	ScrollY    float32      `json:"scrollY"`
	Win        *[2]int      `json:"win"`
	WinOff     *[2]float64  `json:"win_off"`
	Dt         *float64     `json:"dt"`
	Iterations *int         `json:"iterations"`
}

func main() {
	trace := flag.String("trace", "", "JSON-lines trace (make_golden.py --dump-trace)")
	out := flag.String("out", "ref", "prefix of the state dumps: <out>_<frame>.clstate")
	cps := flag.String("checkpoints", "", "comma-separated frame numbers to dump after")
	width := flag.Int("width", 1024, "Cloth.Width  (main.go:150: Dp(1024))")
	height := flag.Int("height", 211, "Cloth.Height (main.go:151: Dp(640*0.33))")
	spacing := flag.Int("spacing", 6, "cloth spacing (main.go:55)")
	posX := flag.Int("posx", 128, "Init posX (main.go:160-163)")
	posY := flag.Int("posy", 164, "Init posY (main.go:164)")
	flag.Parse()

	checkpoint := map[int]bool{}
	for _, s := range strings.Split(*cps, ",") {
		if s = strings.TrimSpace(s); s != "" {
			n, err := strconv.Atoi(s)
			if err != nil {
				log.Fatalf("bad checkpoint %q", s)
			}
			checkpoint[n] = true
		}
	}

	hud := gui.NewHud()
	cloth := physics.NewCloth(*width, *height, *spacing, color.NRGBA{R: 0x55, A: 0xff}) // main.go:149-158
	cloth.Init(*posX, *posY, hud)                                                        // main.go:160-166
	mouse := &physics.Mouse{}
	mouse.SetScrollY(50)     // consts.DefaultFocusArea, main.go:84
	mouse.SetMaxScrollY(120) // consts.MaxFocusArea,     main.go:85

	in, err := os.Open(*trace)
	if err != nil {
		log.Fatal(err)
	}
	defer in.Close()
	sc := bufio.NewScanner(in)
	sc.Buffer(make([]byte, 1<<20), 1<<24)

	var physicsTime time.Duration
	frames := 0
	for ; sc.Scan(); frames++ {
		var r record
		if err := json.Unmarshal(sc.Bytes(), &r); err != nil {
			log.Fatalf("frame %d: %v", frames, err)
		}
		if r.Iterations != nil && *r.Iterations != 1 {
			log.Fatalf("frame %d: the reference relaxes once per frame (cloth.go:96-100), trace asks for %d", frames, *r.Iterations)
		}
		// every UpdatePosition of the frame, in order: px,py are whatever x,y were before the LAST call (mouse.go:21-27)
		for _, m := range r.Moves {
			mouse.UpdatePosition(m[0], m[1])
		}
		if r.Left {
			mouse.SetLeftButton()
		} else {
			mouse.ReleaseLeftButton()
		}
		if r.Right {
			mouse.SetRightButton()
		} else {
			mouse.ReleaseRightButton()
		}
		mouse.SetDragging(r.Dragging)
		mouse.SetCtrlDown(r.Ctrl)
		mouse.SetForce(r.Force)
		mouse.SetScrollY(unit.Dp(r.ScrollY))
		win := [2]int{1280, 820} // consts.WindowSizeX/Y
		if r.Win != nil {
			win = *r.Win
		}
		hud.WinOffsetX, hud.WinOffsetY = 0, 0
		if r.WinOff != nil {
			hud.WinOffsetX, hud.WinOffsetY = r.WinOff[0], r.WinOff[1]
		}
		dt := 0.022 // consts.Delta
		if r.Dt != nil {
			dt = *r.Dt
		}
		gtx := layout.Context{Ops: new(op.Ops), Constraints: layout.Exact(image.Pt(win[0], win[1]))}
		t0 := time.Now()
		cloth.Update(gtx, mouse, hud, dt) // physics + the two render loops (cloth.go:85-138): what the GUI pays per frame
		physicsTime += time.Since(t0)
		if checkpoint[frames] {
			path := fmt.Sprintf("%s_%04d.clstate", *out, frames)
			if err := cloth.Dump(path, *posX, *posY, uint64(frames+1)); err != nil {
				log.Fatal(err)
			}
		}
	}
	if err := sc.Err(); err != nil {
		log.Fatal(err)
	}
	fmt.Printf("{\"frames\": %d, \"update_ns_per_frame\": %.1f, \"slice_len\": %d}\n",
		frames, float64(physicsTime.Nanoseconds())/float64(max(frames, 1)), cloth.SliceLen())
}
