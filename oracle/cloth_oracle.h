/*
 * cloth_oracle.h — CPU oracle for the per-frame cloth update of esimov/cloth-physics.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.  The product
 * (libclothgpu.so) never links, loads or calls it and has no CPU fallback.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or fixtures, and it cannot be built here
 * (no Go toolchain, Gio not vendored), so this restatement is pinned only by known-answer tests derived
 * from the reference's source text and by bit-for-bit agreement with a second, independent restatement (a
 * line-by-line Python transliteration of the Go files that keeps their data structures and slice semantics,
 * tests/go_transliteration.py; both in tests/test_oracle_kat.py) — not by outputs of the reference itself.
 *
 * The only thing shared with the product is the POD input types of include/clothgpu.h
 * (clothgpu_desc / clothgpu_frame and the CLOTHGPU_FLAG_* bit layout), so one harness drives both.
 */
#ifndef CLOTH_ORACLE_H
#define CLOTH_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#include "../include/clothgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct oracle_cloth oracle_cloth;

/* How the stick loop of Cloth.Update (physics/cloth.go:96-100) is ordered and how tearing is modelled. */
enum {
    /* What the Go binary does: sticks in Cloth.Init creation order, one in-place Gauss-Seidel sweep,
     * tearing removes the stick from the slice WHILE it is being ranged over (constraint.go:57-64),
     * which skips the following stick and re-visits the slice's stale tail. */
    ORACLE_SEQ_FAITHFUL = 0,
    /* Same order, but tearing only marks the stick dead (no skip / re-visit artefact). */
    ORACLE_SEQ_MASKED = 1,
    /* Deterministic mode the GPU must match bit for bit: per iteration four conflict-free colour
     * passes H-even, H-odd, V-even, V-odd; tearing marks the stick dead. */
    ORACLE_COLOUR = 2
};

/* NewCloth + Init (physics/cloth.go:31-81).  Returns NULL on a bad descriptor (e.g. the reference pin
 * rule with width/spacing < 10, where the Go code panics with an integer divide by zero,
 * cloth.go:72).  desc->precision selects binary64 or binary32 arithmetic; desc->strip_* select a
 * row strip with ghost rows exactly like the product. */
oracle_cloth* oracle_create(const clothgpu_desc* desc, int mode);
void oracle_destroy(oracle_cloth* c);
/* (*Cloth).Reset, physics/cloth.go:142-148 */
void oracle_reset(oracle_cloth* c, int32_t pos_x, int32_t pos_y);

/* The physics half of (*Cloth).Update, physics/cloth.go:92-100.  `threads` > 1 parallelises the
 * particle loop and (ORACLE_COLOUR only) each colour pass with OpenMP; results do not depend on it. */
void oracle_step(oracle_cloth* c, const clothgpu_frame* f, int threads);

int32_t oracle_nx(const oracle_cloth* c);
int32_t oracle_ny(const oracle_cloth* c);
/* rows held (owned + ghosts) and the global index of the first held row */
int32_t oracle_held_rows(const oracle_cloth* c);
int32_t oracle_held_row0(const oracle_cloth* c);
/* len(cloth.constraints) */
int64_t oracle_slice_len(const oracle_cloth* c);
/* stick visits (constraint.Update calls) so far */
uint64_t oracle_relaxations(const oracle_cloth* c);
/* sticks removed by the last oracle_step */
uint64_t oracle_torn_last(const oracle_cloth* c);

/* Copy `rows` global rows starting at `row0` (must be held) to/from row-major host arrays, nx per
 * row.  flags use the CLOTHGPU_FLAG_* layout.  NULL planes are skipped. */
int oracle_read_rows(const oracle_cloth* c, int32_t row0, int32_t rows, double* x, double* y,
                     double* px, double* py, uint8_t* flags);
int oracle_write_rows(oracle_cloth* c, int32_t row0, int32_t rows, const double* x, const double* y,
                      const double* px, const double* py, const uint8_t* flags);

/* Visit order of the last ORACLE_SEQ_FAITHFUL step: creation ids of the sticks that the range loop
 * looked at (before the isActive test), up to cap entries; returns the number of visits. */
int64_t oracle_last_visit_order(const oracle_cloth* c, int64_t* ids, int64_t cap);

/* addSegment (physics/cloth.go:150-157) with normal (:160-168): the four float32 corners {a+n, b+n, b-n, a-n} of the
 * outline quad the reference's render loops emit for a stick with float32 endpoints ab = {ax, ay, bx, by}. */
void oracle_segment_quad(const float* ab, float line_width, float* quad8);

/* Stand-alone kernels for known-answer tests. */
/* constraint.Update on one stick (physics/constraint.go:25-54).  Returns 1 if it tore. */
int oracle_stick_update(double* x1, double* y1, double* x2, double* y2, int pin1, int pin2,
                        double length, int dragging, double tear_length, double relax_gain);

#ifdef __cplusplus
}
#endif
#endif
