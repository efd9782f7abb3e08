//go:build cgo && !refharness

// Package physics — cgo binding of libclothgpu.so behind the reference's exported surface.
//
// SOURCE ONLY: this image has no Go toolchain, so this file is not compiled or tested here; the
// verifiable boundary is the C ABI (include/clothgpu.h), exercised through the identical entry points by
// cloth-physics_b200/python/clothgpu (ctypes) and cloth-physics_b200/host/physics.hpp (C++).
//
// Drop-in recipe (INTEGRATION.md): in a checkout of esimov/cloth-physics delete physics/particle.go and
// physics/constraint.go, replace physics/cloth.go by this file, keep physics/mouse.go unchanged and add
// mouse_frame.go next to it.  main.go keeps calling
//
//	cloth := physics.NewCloth(w, h, spacing, col)   // physics/cloth.go:31  (main.go:158)
//	cloth.Init(startX, startY, hud)                  // physics/cloth.go:42  (main.go:166,217)
//	cloth.Update(gtx, mouse, hud, delta)             // physics/cloth.go:85  (main.go:310)
//	cloth.Reset(startX, startY, hud)                 // physics/cloth.go:142 (main.go:193)
//
// and reads/writes cloth.Width / cloth.Height exactly as before.
package physics

/*
#cgo CFLAGS: -I${SRCDIR}/../../../../include
#cgo LDFLAGS: -L${SRCDIR}/../../../csrc -lclothgpu -Wl,-rpath,${SRCDIR}/../../../csrc
#include <stdlib.h>
#include "clothgpu.h"
*/
import "C"

import (
	"fmt"
	"image/color"
	"runtime"
	"unsafe"

	"gioui.org/f32"
	"gioui.org/layout"
	"gioui.org/op/clip"
	"gioui.org/op/paint"

	"github.com/esimov/cloth-physics/gui"
	"github.com/esimov/cloth-physics/utils"
)

const strokeWidth = 0.6

// Cloth keeps the exported fields of the reference type (physics/cloth.go:18-27); particles and sticks
// live in HBM behind the handle.
type Cloth struct {
	Width  int
	Height int

	spacing int
	tint    color.NRGBA
	gpu     *C.clothgpu
	// grid the handle was created for and its stick count at Init (constant per handle: read once, not per frame)
	nx, ny    int
	capSticks int
	posX      int
	posY      int
	frames    uint64
	// host staging for the render export, reused across frames: page-locked memory of clothgpu_host_alloc (the export
	// kernel stores the GUI cloth's list straight into it — no copies), viewed as Go slices
	quadsMem, hlMem unsafe.Pointer
	quads           []float32
	hl              []uint8
}

// NewCloth mirrors physics/cloth.go:31-38.  No device work happens until Init.
func NewCloth(width, height, spacing int, col color.NRGBA) *Cloth {
	return &Cloth{Width: width, Height: height, spacing: spacing, tint: col}
}

// Init mirrors physics/cloth.go:42-81: idempotent once initialised.  The reference panics with an
// integer divide by zero when Width/spacing < 10 (cloth.go:72); clothgpu_create reports the same
// condition as CLOTHGPU_ERR_INVALID and we turn it back into a panic.
func (c *Cloth) Init(posX, posY int, hud *gui.Hud) {
	if c.gpu != nil {
		return
	}
	var d C.clothgpu_desc
	C.clothgpu_desc_default(&d)
	d.width, d.height, d.spacing = C.int32_t(c.Width), C.int32_t(c.Height), C.int32_t(c.spacing)
	d.pos_x, d.pos_y = C.int32_t(posX), C.int32_t(posY)
	d.pin_rule = C.CLOTHGPU_PIN_REFERENCE
	d.precision = C.CLOTHGPU_F64
	if rc := C.clothgpu_create(&d, &c.gpu); rc != C.CLOTHGPU_OK {
		panic(fmt.Sprintf("clothgpu_create: %s", C.GoString(C.clothgpu_last_error())))
	}
	var info C.clothgpu_info
	C.clothgpu_get_info(c.gpu, &info)
	c.nx, c.ny, c.capSticks = int(info.nx), int(info.ny), int(info.sticks)
	c.posX, c.posY, c.frames = posX, posY, 0
	// Close clears the finalizer again, so Reset may re-Init the same object after a resize (main.go:190-193) without
	// "runtime.SetFinalizer: finalizer already set".
	runtime.SetFinalizer(c, (*Cloth).Close)
}

// Reset mirrors physics/cloth.go:142-148.
func (c *Cloth) Reset(startX, startY int, hud *gui.Hud) {
	if c.gpu == nil {
		c.Init(startX, startY, hud)
		return
	}
	// Width/Height may have been changed by main.go:190-191; a different grid needs a new handle.
	if c.nx != c.Width/c.spacing+1 || c.ny != c.Height/c.spacing+1 {
		c.Close()
		c.Init(startX, startY, hud)
		return
	}
	c.check(C.clothgpu_reset(c.gpu, C.int32_t(startX), C.int32_t(startY)))
	c.posX, c.posY, c.frames = startX, startY, 0
}

// Close releases the device memory (the reference relies on the GC; there is no counterpart).
func (c *Cloth) Close() {
	if c.gpu != nil {
		runtime.SetFinalizer(c, nil)
		C.clothgpu_destroy(c.gpu)
		c.gpu = nil
		C.clothgpu_host_free(c.quadsMem)
		C.clothgpu_host_free(c.hlMem)
		c.quadsMem, c.hlMem, c.quads, c.hl = nil, nil, nil, nil
	}
}

// Dump writes a CLSTATE1 snapshot (include/clothgpu.h) — same signature as the Dump that dump_harness.go adds to the
// reference's package, so host/go/harness drives either implementation.
func (c *Cloth) Dump(path string, posX, posY int, frames uint64) error {
	cs := C.CString(path)
	defer C.free(unsafe.Pointer(cs))
	if rc := C.clothgpu_save_state(c.gpu, cs); rc != C.CLOTHGPU_OK {
		return fmt.Errorf("clothgpu_save_state: %s", C.GoString(C.clothgpu_last_error()))
	}
	return nil
}

// SliceLen is len(cloth.constraints) of the reference: the sticks not torn yet.
func (c *Cloth) SliceLen() int {
	var k C.clothgpu_counters
	c.check(C.clothgpu_get_counters(c.gpu, &k))
	return int(k.alive_sticks)
}

// Update mirrors physics/cloth.go:85-138: the physics half runs on the GPU, the render half walks the
// compacted live-stick list the device exports (same order and same float32 endpoints as the two
// loops of cloth.go:108-114 and 126-134).
func (c *Cloth) Update(gtx layout.Context, mouse *Mouse, hud *gui.Hud, dt float64) {
	// cloth.go:86-90: the highlight path's colour is the cloth's red lightened by the drag force
	dragForce := float32(mouse.GetForce() * 0.1)
	clothColor := color.NRGBA{R: 0x55, A: 0xff}
	col := utils.LinearFromSRGB(clothColor).HSLA().Lighten(dragForce).RGBA().SRGB()

	var f C.clothgpu_frame
	C.clothgpu_frame_default(&f)
	// particle.go:78-83: the five sliders are read as float32 and widened at the point of use
	f.slider[C.CLOTHGPU_SLIDER_DRAG_FORCE] = C.float(hud.Sliders[gui.HudSliderDragForce].Widget.Value)
	f.slider[C.CLOTHGPU_SLIDER_GRAVITY] = C.float(hud.Sliders[gui.HudSliderGravityForce].Widget.Value)
	f.slider[C.CLOTHGPU_SLIDER_STIFFNESS] = C.float(hud.Sliders[gui.HudSliderStiffness].Widget.Value)
	f.slider[C.CLOTHGPU_SLIDER_FRICTION] = C.float(hud.Sliders[gui.HudSliderFriction].Widget.Value)
	f.slider[C.CLOTHGPU_SLIDER_TEAR_DISTANCE] = C.float(hud.Sliders[gui.HudSliderTearDistance].Widget.Value)
	f.drag_force_max = C.float(hud.Sliders[gui.HudSliderDragForce].Max) // particle.go:97
	mouse.fillFrame(unsafe.Pointer(&f))                                  // mouse_frame.go
	f.win_w = C.int32_t(gtx.Constraints.Max.X)                          // particle.go:102
	f.win_h = C.int32_t(gtx.Constraints.Max.Y)
	f.win_off_x, f.win_off_y = C.double(hud.WinOffsetX), C.double(hud.WinOffsetY) // particle.go:89-90
	f.dt = C.double(dt)
	f.iterations = 1 // the reference relaxes once per frame (cloth.go:96-100)
	c.check(C.clothgpu_step(c.gpu, &f))
	c.frames++

	// render half, cloth.go:102-138: the device applies addSegment / normal (cloth.go:150-168) to every live stick and
	// hands back the four corners of its outline quad, in the order the two loops of cloth.go:108-114 / 126-134 walk
	var n C.size_t
	if c.capSticks > 0 { // (a 1 x 1 "cloth" has no sticks: nothing to read back, and &c.quads[0] would panic)
		if len(c.quads) < 8*c.capSticks {
			C.clothgpu_host_free(c.quadsMem)
			C.clothgpu_host_free(c.hlMem)
			c.quadsMem, c.hlMem = nil, nil
			c.check(C.clothgpu_host_alloc(C.size_t(32*c.capSticks), &c.quadsMem))
			c.check(C.clothgpu_host_alloc(C.size_t(c.capSticks), &c.hlMem))
			c.quads = unsafe.Slice((*float32)(c.quadsMem), 8*c.capSticks)
			c.hl = unsafe.Slice((*uint8)(c.hlMem), c.capSticks)
		}
		c.check(C.clothgpu_read_quads_f32(c.gpu, 0, C.float(strokeWidth), (*C.float)(unsafe.Pointer(&c.quads[0])),
			(*C.uint8_t)(unsafe.Pointer(&c.hl[0])), C.size_t(c.capSticks), &n))
	}

	var path clip.Path
	path.Begin(gtx.Ops)
	for i := 0; i < int(n); i++ {
		outline(&path, c.quads[8*i:8*i+8])
	}
	paint.FillShape(gtx.Ops, c.tint, clip.Outline{Path: path.End()}.Op())
	path.Begin(gtx.Ops)
	for i := 0; i < int(n); i++ {
		if c.hl[i] != 0 {
			outline(&path, c.quads[8*i:8*i+8])
		}
	}
	paint.FillShape(gtx.Ops, color.NRGBA{R: col.R, A: col.A}, clip.Outline{Path: path.End()}.Op()) // cloth.go:136
}

// outline adds one stick's quad {a+n, b+n, b-n, a-n} to the path, as addSegment does (cloth.go:150-157).
func outline(p *clip.Path, q []float32) {
	p.MoveTo(f32.Pt(q[0], q[1]))
	p.LineTo(f32.Pt(q[2], q[3]))
	p.LineTo(f32.Pt(q[4], q[5]))
	p.LineTo(f32.Pt(q[6], q[7]))
	p.Close()
}

func (c *Cloth) check(rc C.int) {
	if rc != C.CLOTHGPU_OK {
		panic(fmt.Sprintf("clothgpu: %s", C.GoString(C.clothgpu_last_error())))
	}
}
