//go:build cgo && !refharness

// mouse_frame.go — the one addition next to the reference's unchanged physics/mouse.go: copy the
// Mouse (physics/mouse.go:9-19) into the fields of a clothgpu_frame, i.e. everything
// particle.update / constraint.Update read from it (particle.go:96-149, constraint.go:35).
// SOURCE ONLY (no Go toolchain in this image).
package physics

/*
#include "clothgpu.h"
*/
import "C"

import "unsafe"

func (m *Mouse) fillFrame(frame unsafe.Pointer) {
	f := (*C.clothgpu_frame)(frame)
	f.mouse_x, f.mouse_y = C.double(m.x), C.double(m.y)
	f.mouse_px, f.mouse_py = C.double(m.px), C.double(m.py)
	f.mouse_force = C.double(m.force)
	f.scroll_y = C.float(m.scrollY)
	var b C.uint32_t
	if m.leftDown {
		b |= C.CLOTHGPU_MOUSE_LEFT
	}
	if m.rightDown {
		b |= C.CLOTHGPU_MOUSE_RIGHT
	}
	if m.isDragging {
		b |= C.CLOTHGPU_MOUSE_DRAGGING
	}
	if m.ctrlDown {
		b |= C.CLOTHGPU_MOUSE_CTRL
	}
	f.buttons = b
}
