//go:build refharness

// dump_harness.go — the one file the parity harness ADDS to the reference's own `physics` package (copy it next to the
// reference's unmodified cloth.go / particle.go / constraint.go and build with `-tags refharness`; the tag keeps it out of
// the cgo shim's build, whose Cloth has no particle slice — the shim's Dump is in clothgpu_cgo.go).  particle and
// constraint are unexported (physics/particle.go:16, physics/constraint.go:10), so only a file inside the package can walk
// them.  It changes nothing: it reads cloth.particles / cloth.constraints in slice order and writes a CLSTATE1 snapshot
// (include/clothgpu.h: clothgpu_save_state writes the same file), so a state of the REAL reference can be diffed against
// the tests' CPU checker's and the GPU library's with tests/golden/compare_states.py.
//
// SOURCE ONLY: this image has no Go toolchain, so the file has never been compiled here.
package physics

import (
	"encoding/binary"
	"fmt"
	"io"
	"math"
	"os"
)

const (
	flagPin       = 1 << 0 // particle.pinX        physics/particle.go:23
	flagActive    = 1 << 1 // particle.isActive    physics/particle.go:24
	flagHighlight = 1 << 2 // particle.highlighted physics/particle.go:25
	flagAliveH    = 1 << 3 // the stick to the right neighbour is still in Cloth.constraints
	flagAliveV    = 1 << 4 // the stick to the particle below is still in Cloth.constraints
)

type fnvWriter struct {
	w   io.Writer
	sum uint64
	err error
}

func (f *fnvWriter) Write(p []byte) (int, error) {
	for _, b := range p {
		f.sum = (f.sum ^ uint64(b)) * 0x100000001b3
	}
	n, err := f.w.Write(p)
	if err != nil && f.err == nil {
		f.err = err
	}
	return n, err
}

// Dump writes the cloth as a CLSTATE1 file: header, the planes x, y, px, py (float64, row-major = the order of
// cloth.particles, physics/cloth.go:51-77), one flag byte per particle, FNV-1a 64 checksum.  posX / posY are the Init
// origin (the reference does not keep them), frames the number of Update calls so far.
func (c *Cloth) Dump(path string, posX, posY int, frames uint64) error {
	nx, ny := c.Width/c.spacing+1, c.Height/c.spacing+1 // physics/cloth.go:43-44, 51-52 (inclusive loops)
	if len(c.particles) != nx*ny {
		return fmt.Errorf("dump: %d particles, expected %d x %d", len(c.particles), nx, ny)
	}
	index := make(map[*particle]int, len(c.particles))
	flags := make([]byte, (len(c.particles)+7)/8*8)
	for i, p := range c.particles {
		index[p] = i
		if p.pinX {
			flags[i] |= flagPin
		}
		if p.isActive {
			flags[i] |= flagActive
		}
		if p.highlighted {
			flags[i] |= flagHighlight
		}
	}
	// stick membership: a constraint's p1 is the top / left particle (physics/cloth.go:61-70)
	for _, s := range c.constraints {
		i1, i2 := index[s.p1], index[s.p2]
		switch i2 - i1 {
		case 1:
			flags[i1] |= flagAliveH
		case nx:
			flags[i1] |= flagAliveV
		default:
			return fmt.Errorf("dump: stick between particles %d and %d is not a grid stick", i1, i2)
		}
	}
	f, err := os.Create(path)
	if err != nil {
		return err
	}
	defer f.Close()
	w := &fnvWriter{w: f, sum: 0xcbf29ce484222325}
	le := binary.LittleEndian
	hdr := make([]byte, 64)
	copy(hdr, "CLSTATE1")
	le.PutUint32(hdr[8:], 64)
	le.PutUint32(hdr[12:], 1)
	for k, v := range []int{nx, ny, 0, ny, 1, c.spacing, posX, posY} {
		le.PutUint32(hdr[16+4*k:], uint32(int32(v)))
	}
	le.PutUint64(hdr[48:], frames)
	w.Write(hdr)
	plane := make([]byte, 8*len(c.particles))
	for k := 0; k < 4; k++ {
		for i, p := range c.particles {
			v := [4]float64{p.x, p.y, p.px, p.py}[k]
			le.PutUint64(plane[8*i:], math.Float64bits(v))
		}
		w.Write(plane)
	}
	w.Write(flags)
	var tail [8]byte
	le.PutUint64(tail[:], w.sum)
	if _, err := f.Write(tail[:]); err != nil {
		return err
	}
	return w.err
}

// SliceLen is len(cloth.constraints) — the tear artefact of physics/constraint.go:57-64 shows up here.
func (c *Cloth) SliceLen() int { return len(c.constraints) }
