/*
 * cloth_oracle.c — CPU oracle for the per-frame cloth update of esimov/cloth-physics
 * (physics/cloth.go:42-100, physics/particle.go:75-189, physics/constraint.go:25-64).
 *
 * TEST INFRASTRUCTURE ONLY — never linked, loaded or called by the product (libclothgpu.so).
 * PARITY UNPINNED — see cloth_oracle.h: the reference has no tests or fixtures and cannot be built
 * here, so this file is pinned by source-derived known-answer tests only.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off -fno-fast-math: IEEE binary64, no FMA
 * contraction, the arithmetic Go emits on amd64).
 */
#include "cloth_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>

#define CAT_(a, b) a##b
#define CAT(a, b) CAT_(a, b)

#define REAL double
#define SQRT sqrt
#define P(name) CAT(name, _f64)
#include "cloth_oracle_impl.inc"
#undef REAL
#undef SQRT
#undef P

#define REAL float
#define SQRT sqrtf
#define P(name) CAT(name, _f32)
#include "cloth_oracle_impl.inc"
#undef REAL
#undef SQRT
#undef P

struct oracle_cloth {
    clothgpu_desc desc;
    int mode;
    int nx, ny;
    int held_row0, held_rows;
    cloth_f64* c64;
    cloth_f32* c32;
    uint64_t relaxations;
    uint64_t torn_last;
    int64_t* visit;
    int64_t n_visit, cap_visit;
};

static void build(oracle_cloth* o, int pos_x, int pos_y) {
    const clothgpu_desc* d = &o->desc;
    if (d->precision == CLOTHGPU_F32)
        o->c32 = cloth_init_f32(o->nx, o->held_row0, o->held_rows, d->spacing, pos_x, pos_y, d->pin_rule);
    else
        o->c64 = cloth_init_f64(o->nx, o->held_row0, o->held_rows, d->spacing, pos_x, pos_y, d->pin_rule);
    o->relaxations = 0;
    o->torn_last = 0;
    o->n_visit = 0;
}

oracle_cloth* oracle_create(const clothgpu_desc* desc, int mode) {
    if (!desc || desc->spacing <= 0 || desc->width < 0 || desc->height < 0) return NULL;
    if (mode < ORACLE_SEQ_FAITHFUL || mode > ORACLE_COLOUR) return NULL;
    int clothX = desc->width / desc->spacing;  /* physics/cloth.go:43 */
    int clothY = desc->height / desc->spacing; /* physics/cloth.go:44 */
    /* physics/cloth.go:72: x % (clothX / 10) panics (integer divide by zero) when clothX < 10 */
    if (desc->pin_rule == CLOTHGPU_PIN_REFERENCE && clothX / 10 == 0) return NULL;
    oracle_cloth* o = (oracle_cloth*)calloc(1, sizeof(*o));
    o->desc = *desc;
    o->mode = mode;
    o->nx = clothX + 1;
    o->ny = clothY + 1;
    if (desc->strip_rows > 0) {
        int kmax = desc->max_iterations < 1 ? 1 : desc->max_iterations;
        int ghost = 2 * kmax;
        int lo = desc->strip_row0 - ghost, hi = desc->strip_row0 + desc->strip_rows + ghost;
        if (desc->strip_row0 < 0 || desc->strip_row0 + desc->strip_rows > o->ny || mode != ORACLE_COLOUR) {
            free(o);
            return NULL;
        }
        if (lo < 0) lo = 0;
        if (hi > o->ny) hi = o->ny;
        o->held_row0 = lo;
        o->held_rows = hi - lo;
    } else {
        o->held_row0 = 0;
        o->held_rows = o->ny;
    }
    build(o, desc->pos_x, desc->pos_y);
    return o;
}

static void drop(oracle_cloth* o) {
    cloth_free_f64(o->c64);
    cloth_free_f32(o->c32);
    o->c64 = NULL;
    o->c32 = NULL;
}

void oracle_destroy(oracle_cloth* o) {
    if (!o) return;
    drop(o);
    free(o->visit);
    free(o);
}

/* (*Cloth).Reset, physics/cloth.go:142-148: drop both slices and Init again. */
void oracle_reset(oracle_cloth* o, int32_t pos_x, int32_t pos_y) {
    drop(o);
    build(o, pos_x, pos_y);
}

void oracle_step(oracle_cloth* o, const clothgpu_frame* f, int threads) {
    if (threads < 1) threads = 1;
    if (o->c64)
        cloth_step_f64(o->c64, f, o->mode, threads, &o->relaxations, &o->torn_last, &o->visit,
                       &o->n_visit, &o->cap_visit);
    else
        cloth_step_f32(o->c32, f, o->mode, threads, &o->relaxations, &o->torn_last, &o->visit,
                       &o->n_visit, &o->cap_visit);
}

int32_t oracle_nx(const oracle_cloth* o) { return o->nx; }
int32_t oracle_ny(const oracle_cloth* o) { return o->ny; }
int32_t oracle_held_rows(const oracle_cloth* o) { return o->held_rows; }
int32_t oracle_held_row0(const oracle_cloth* o) { return o->held_row0; }
int64_t oracle_slice_len(const oracle_cloth* o) { return o->c64 ? o->c64->slice_len : o->c32->slice_len; }
uint64_t oracle_relaxations(const oracle_cloth* o) { return o->relaxations; }
uint64_t oracle_torn_last(const oracle_cloth* o) { return o->torn_last; }

int oracle_read_rows(const oracle_cloth* o, int32_t row0, int32_t rows, double* x, double* y,
                     double* px, double* py, uint8_t* flags) {
    if (row0 < o->held_row0 || row0 + rows > o->held_row0 + o->held_rows || rows < 0) return -1;
    int64_t first = (int64_t)(row0 - o->held_row0) * o->nx, count = (int64_t)rows * o->nx;
    if (o->c64)
        cloth_read_f64(o->c64, o->nx, first, count, x, y, px, py, flags);
    else
        cloth_read_f32(o->c32, o->nx, first, count, x, y, px, py, flags);
    return 0;
}

int oracle_write_rows(oracle_cloth* o, int32_t row0, int32_t rows, const double* x, const double* y,
                      const double* px, const double* py, const uint8_t* flags) {
    if (row0 < o->held_row0 || row0 + rows > o->held_row0 + o->held_rows || rows < 0) return -1;
    int64_t first = (int64_t)(row0 - o->held_row0) * o->nx, count = (int64_t)rows * o->nx;
    if (o->c64)
        cloth_write_f64(o->c64, first, count, x, y, px, py, flags);
    else
        cloth_write_f32(o->c32, first, count, x, y, px, py, flags);
    return 0;
}

int64_t oracle_last_visit_order(const oracle_cloth* o, int64_t* ids, int64_t cap) {
    int64_t n = o->n_visit < cap ? o->n_visit : cap;
    if (ids && n > 0) memcpy(ids, o->visit, (size_t)n * sizeof(int64_t));
    return o->n_visit;
}

/* math.Hypot as the amd64 assembly (and the portable Go version) computes it: max * sqrt(1 + (min/max)^2),
 * finite arguments only. */
static double go_hypot(double p, double q) {
    p = fabs(p), q = fabs(q);
    if (p < q) {
        double t = p;
        p = q, q = t;
    }
    if (p == 0) return 0;
    q = q / p;
    return p * sqrt(1 + q * q);
}

void oracle_segment_quad(const float* ab, float line_width, float* quad8) {
    const float ax = ab[0], ay = ab[1], bx = ab[2], by = ab[3];
    float dirx = bx - ax, diry = by - ay;            /* dir := b.Sub(a)                      cloth.go:161 */
    const float rx = +diry, ry = -dirx;              /* dir.X, dir.Y = +dir.Y, -dir.X        :162 */
    dirx = rx, diry = ry;
    const double d = go_hypot((double)dirx, (double)diry);  /* :163 */
    float nx = 0.0f, ny = 0.0f;                      /* f32.Point{} when |d| < 1e-5          :164-166 */
    if (!(fabs(d) < 1e-5)) {
        const float s = line_width / (float)d;       /* dir.Mul(w / float32(d))              :167 */
        nx = dirx * s, ny = diry * s;
    }
    quad8[0] = ax + nx, quad8[1] = ay + ny;          /* p.MoveTo(a.Add(n))                   :152 */
    quad8[2] = bx + nx, quad8[3] = by + ny;          /* p.LineTo(b.Add(n))                   :153 */
    quad8[4] = bx - nx, quad8[5] = by - ny;          /* p.LineTo(b.Sub(n))                   :154 */
    quad8[6] = ax - nx, quad8[7] = ay - ny;          /* p.LineTo(a.Sub(n))                   :155 */
}

int oracle_stick_update(double* x1, double* y1, double* x2, double* y2, int pin1, int pin2,
                        double length, int dragging, double tear_length, double relax_gain) {
    particle_f64 a = {0}, b = {0};
    a.x = *x1, a.y = *y1, a.pinX = (uint8_t)pin1, a.isActive = 1;
    b.x = *x2, b.y = *y2, b.pinX = (uint8_t)pin2, b.isActive = 1;
    constraint_f64 c = {0};
    c.p1 = &a, c.p2 = &b, c.length = length, c.alive = 1;
    clothgpu_frame f;
    memset(&f, 0, sizeof f);
    f.buttons = dragging ? CLOTHGPU_MOUSE_DRAGGING : 0;
    f.tear_length = tear_length;
    f.relax_gain = relax_gain;
    int tore = constraint_update_f64(NULL, &c, &f, 0);
    *x1 = a.x, *y1 = a.y, *x2 = b.x, *y2 = b.y;
    return tore;
}
