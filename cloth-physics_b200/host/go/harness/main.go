// harness — regenerates tests/golden/*.npz-equivalent dumps from the REAL reference once a Go toolchain
// (and Gio v0.3.1 in the module cache) is available; also the headless timing arm for BASELINE.md.
// SOURCE ONLY: cannot be built in this image.  It drives the unmodified reference package headless:
// a layout.Context with an empty op list and exact constraints, gui.NewHud(), and the Mouse setters
// are all constructible without a window (SURVEY.md §8c).
//
//	go run ./harness -frames 1000 -trace ../../../tests/golden/c0_trace.jsonl -out c0_ref.bin
//
// Input: one JSON object per frame with the clothgpu_frame fields (written by
// tests/golden/make_golden.py --dump-trace).  Output: after the listed checkpoint frames, the raw
// little-endian float64 planes x,y,px,py and one flag byte per particle in the CLOTHGPU_FLAG_* layout,
// directly comparable with the fixtures under tests/golden/.
package main

import (
	"bufio"
	"encoding/binary"
	"encoding/json"
	"flag"
	"image"
	"os"
	"time"

	"gioui.org/layout"
	"gioui.org/op"
	"gioui.org/unit"

	"github.com/esimov/cloth-physics/gui"
	"github.com/esimov/cloth-physics/physics"
)

type frame struct {
	MouseX, MouseY   float64
	Left, Right      bool
	Dragging, Ctrl   bool
	Force            float64
	ScrollY          float32
	WinW, WinH       int
	Dt               float64
	Checkpoint       bool
}

func main() {
	frames := flag.Int("frames", 1000, "frames to step")
	trace := flag.String("trace", "", "JSON-lines trace")
	out := flag.String("out", "ref.bin", "state dump")
	flag.Parse()

	hud := gui.NewHud()
	cloth := physics.NewCloth(1024, 211, 6, [4]uint8{}) // main.go:149-158
	cloth.Init(128, 164, hud)                           // main.go:160-166
	mouse := &physics.Mouse{}
	mouse.SetScrollY(50)

	in, _ := os.Open(*trace)
	defer in.Close()
	sc := bufio.NewScanner(in)
	w, _ := os.Create(*out)
	_ = binary.LittleEndian
	defer w.Close()

	t0 := time.Now()
	for i := 0; i < *frames && sc.Scan(); i++ {
		var f frame
		json.Unmarshal(sc.Bytes(), &f)
		mouse.UpdatePosition(f.MouseX, f.MouseY)
		if f.Left {
			mouse.SetLeftButton()
		} else {
			mouse.ReleaseLeftButton()
		}
		if f.Right {
			mouse.SetRightButton()
		} else {
			mouse.ReleaseRightButton()
		}
		mouse.SetDragging(f.Dragging)
		mouse.SetCtrlDown(f.Ctrl)
		mouse.SetForce(f.Force)
		mouse.SetScrollY(unit.Dp(f.ScrollY))
		gtx := layout.Context{Ops: new(op.Ops), Constraints: layout.Exact(image.Pt(f.WinW, f.WinH))}
		cloth.Update(gtx, mouse, hud, f.Dt)
		if f.Checkpoint {
			// particle and constraint are unexported: the dump itself lives in a file added to
			// package physics for the harness build (physics/dump_harness.go: func (c *Cloth)
			// Dump(w io.Writer) walking c.particles / c.constraints in slice order).
			cloth.Dump(w)
		}
	}
	binary.Write(os.Stderr, binary.LittleEndian, time.Since(t0).Nanoseconds())
}
