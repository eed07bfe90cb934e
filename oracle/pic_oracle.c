/*
 * pic_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of the OSIRIS per-timestep PIC
 * hot path.  It is the checker for the CUDA path (tests/, __graft_entry__.smoke(), and the
 * cpu_baseline leg of bench.py); nothing in the product path (osiris_b200/) may call it.
 *
 * Parity status: PINNED.  tests/test_oracle_vs_ref.py runs every function below that has a
 * compilable reference counterpart against the reference's own C kernels, built from
 * /root/reference/source/simd by oracle/Makefile into oracle/_ref/ (AVX2 fp32/fp64, AVX fp32,
 * SSE): x, ix, p bit-exact and J bit-exact against the AVX2 build (same FMA statement, same
 * summation order).  The Fortran-only parts (Yee, guard cells, sort, particle wrap, energies)
 * cannot be compiled here (no Fortran compiler); they are restated from the cited lines and
 * pinned by the invariants in tests/test_oracle_invariants.py.
 *
 * Plain scalar C, one particle at a time.  Data layouts are the reference's own Fortran
 * layouts: x(D,np), ix(D,np), p(3,np), q(np) and grids f(3, S1, S2[, S3]) with the
 * component fastest.
 *
 * Build: gcc -std=gnu99 -O2 -ffp-contract=off -mfma -DPREC_SINGLE|-DPREC_DOUBLE (see Makefile).
 * -ffp-contract=off: every fused multiply-add below is explicit, placed where the reference's
 * AVX2 statement (source/simd/os-spec-push-avx2.c, os-spec-current-avx2.c, splines-avx2.h)
 * places one; the scalar splitter (split-vec.h) has none.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#if defined(PREC_SINGLE)
typedef float REAL;
#define FMA(a, b, c) fmaf((a), (b), (c))
#define SQRT(a) sqrtf(a)
#define VECW 8 /* lanes of the reference's AVX2 fp32 kernels */
#elif defined(PREC_DOUBLE)
typedef double REAL;
#define FMA(a, b, c) fma((a), (b), (c))
#define SQRT(a) sqrt(a)
#define VECW 4 /* lanes of the reference's AVX2 fp64 kernels */
#else
#error define PREC_SINGLE or PREC_DOUBLE
#endif
#define FNMA(a, b, c) FMA(-(a), (b), (c)) /* -(a*b)+c, one rounding (vfnmadd) */
#define FMS(a, b, c) FMA((a), (b), -(c))  /*  (a*b)-c, one rounding (vfmsub)  */

#define R(v) ((REAL)(v))
#define MAXNP 5

int oracle_real_bytes(void) { return (int)sizeof(REAL); }

/* ------------------------------------------------------------------------------------
 * Splines.  simd/splines-avx2.h:27-193 (fp32), same statements in the fp64 half;
 * Fortran emf/os-emf-interpolate.f03:171-352.
 * ---------------------------------------------------------------------------------- */
static void spline_t(int order, REAL t0, REAL t1, REAL *w) {
  switch (order) {
  case 1:
    w[0] = t0;
    w[1] = t1;
    break;
  case 2:
    w[0] = (t0 * t0) * R(0.5);
    w[1] = FMA(t0, t1, R(0.5));
    w[2] = (t1 * t1) * R(0.5);
    break;
  case 3: {
    const REAL c1_2 = R(1.0) / R(2.0), c1_6 = R(1.0) / R(6.0), c2_3 = R(2.0) / R(3.0);
    REAL t2 = t0 * t0, t3 = t1 * t1;
    REAL a0 = t0 * t2, a1 = t1 * t3;
    w[0] = c1_6 * a0;
    w[1] = FMA(c1_2, a1, c2_3 - t3);
    w[2] = FMA(c1_2, a0, c2_3 - t2);
    w[3] = c1_6 * a1;
  } break;
  case 4: {
    const REAL c1_2 = R(1.0) / R(2.0), c1_24 = R(1.0) / R(24.0), c1_4 = R(1.0) / R(4.0);
    const REAL c1 = R(1.0), c3_2 = R(3.0) / R(2.0), c1_6 = R(1.0) / R(6.0);
    const REAL c11_24 = R(11.0) / R(24.0);
    REAL t2 = t0 * t1, t02 = t0 * t0, t12 = t1 * t1;
    w[0] = c1_24 * (t02 * t02);
    w[1] = c1_6 * FMA(t0, FMA(t0, c3_2 + t2, c1), c1_4);
    w[2] = FMA(t2, FMA(c1_4, t2, c1_2), c11_24);
    w[3] = c1_6 * FMA(t1, FMA(t1, c3_2 + t2, c1), c1_4);
    w[4] = c1_24 * (t12 * t12);
  } break;
  }
}
/* vspline_sN: t0 = 1/2 - x, t1 = 1/2 + x */
static void spline(int order, REAL x, REAL *w) { spline_t(order, R(0.5) - x, R(0.5) + x, w); }
/* vsplineh_sN: t0 = (1-h) - x, t1 = h + x */
static void splineh(int order, REAL x, REAL h, REAL *w) {
  spline_t(order, (R(1.0) - h) - x, h + x, w);
}
void oracle_spline(int order, REAL x, REAL *w) { spline(order, x, w); }
void oracle_splineh(int order, REAL x, REAL *w) {
  REAL h = (x < R(0.0)) ? R(1.0) : R(0.0);
  splineh(order, x, h, w);
}

/* vmntrim, simd/os-spec-push-avx2.c:44-58; Fortran ntrim spec/os-spec-push.f03:1332-1359 */
static int ntrim(REAL x) { return (x >= R(0.5)) - (x < R(-0.5)); }
int oracle_ntrim(REAL x) { return ntrim(x); }

/* ------------------------------------------------------------------------------------
 * Longitudinal deposit weights.  simd/os-spec-current-avx2.c:20-116;
 * Fortran spec/os-spec-current.f03:383-479.
 * ---------------------------------------------------------------------------------- */
static void wl(int order, REAL qn, REAL x0, REAL x1, REAL *w) {
  switch (order) {
  case 1:
    w[0] = qn * (x1 - x0);
    break;
  case 2: {
    REAL d = x1 - x0, s1_2 = R(0.5) - x0, p1_2 = R(0.5) + x0, n = qn * d;
    w[0] = n * FNMA(R(0.5), d, s1_2);
    w[1] = n * FMA(R(0.5), d, p1_2);
  } break;
  case 3: {
    const REAL c1_2 = R(0.5), c3_4 = R(0.75), c1_3 = R(1.0) / R(3.0);
    REAL d = x1 - x0, s = c1_2 - x0, p = c1_2 + x0, d_3 = c1_3 * d;
    w[0] = c1_2 * FMS(s, s, d * (s - d_3));
    w[1] = FNMA(d, x0 + d_3, FNMA(x0, x0, c3_4));
    w[2] = c1_2 * FMA(p, p, d * (p + d_3));
    REAL n = qn * d;
    w[0] = n * w[0];
    w[1] = n * w[1];
    w[2] = n * w[2];
  } break;
  case 4: {
    const REAL c1_2 = R(0.5), c1_4 = R(0.25), c1_6 = R(1.0) / R(6.0), c3_2 = R(1.5);
    const REAL c3_4 = R(0.75), c1_3 = R(1.0) / R(3.0), c2_3 = R(2.0) / R(3.0);
    REAL d = x1 - x0, s = c1_2 - x0, p = c1_2 + x0;
    REAL t = x0 - c1_6, u = x0 + c1_6;
    REAL s2 = s * s, s3 = s2 * s, p2 = p * p, p3 = p2 * p;
    REAL d_2 = d * c1_2, d_4 = d * c1_4;
    w[0] = c1_6 * FMA(d, FMS(d, s - d_4, c3_2 * s2), s3);
    w[1] = FMA(d, FMA(d_2, t + d_4, FMS(c3_4, t * t, c1_3)), FMA(c1_2, p3, c2_3 - p2));
    w[2] = FNMA(d, FMA(d_2, u + d_4, FMS(c3_4, u * u, c1_3)), FMA(c1_2, s3, c2_3 - s2));
    w[3] = c1_6 * FMA(d, FMA(d, p + d_4, c3_2 * p2), p3);
    REAL n = qn * d;
    w[0] = n * w[0];
    w[1] = n * w[1];
    w[2] = n * w[2];
    w[3] = n * w[3];
  } break;
  }
}

/* ------------------------------------------------------------------------------------
 * Virtual particles (trajectory segments).  simd/os-spec-push-avx2.h:59-126.
 * ---------------------------------------------------------------------------------- */
typedef struct {
  REAL x0, x1, y0, y1, z0, z1, q, vz;
  int ix, iy, iz;
} vp_t;

/* linear interpolation a + (b - a) * delta of split-vec.h: plain multiply and add (the
 * splitter is scalar C; the parity build of the reference uses -ffp-contract=off) */
#define LERP(a, b, delta) ((a) + ((b) - (a)) * (delta))

/* vsplit2D, simd/split-vec.h:105-254 (Fortran split_2d spec/os-spec-current.f03:659-888).
 * v[0..7] are the group's particles, extras are appended at v[*np]. */
static void split2d(vp_t *v, int k, int *np, int cross, int dix, int diy) {
  REAL delta, xint, yint, vzint, xint2, yint2, vzint2;
  vp_t *a = &v[k], *n;
  switch (cross) {
  case 0:
    break;
  case 1:
    xint = R(0.5) * dix;
    delta = (xint - a->x0) / (a->x1 - a->x0);
    yint = LERP(a->y0, a->y1, delta);
    n = &v[(*np)++];
    n->x0 = -xint; n->x1 = a->x1 - dix;
    n->y0 = yint;  n->y1 = a->y1;
    n->ix = a->ix + dix; n->iy = a->iy;
    n->q = a->q;
    n->vz = a->vz * (1 - delta);
    a->x1 = xint; a->y1 = yint;
    a->vz *= delta;
    break;
  case 2:
    yint = R(0.5) * diy;
    delta = (yint - a->y0) / (a->y1 - a->y0);
    xint = LERP(a->x0, a->x1, delta);
    n = &v[(*np)++];
    n->x0 = xint;  n->x1 = a->x1;
    n->y0 = -yint; n->y1 = a->y1 - diy;
    n->ix = a->ix; n->iy = a->iy + diy;
    n->q = a->q;
    n->vz = a->vz * (1 - delta);
    a->x1 = xint; a->y1 = yint;
    a->vz *= delta;
    break;
  case 3:
    xint = R(0.5) * dix;
    delta = (xint - a->x0) / (a->x1 - a->x0);
    yint = LERP(a->y0, a->y1, delta);
    if ((yint >= R(-0.5)) && (yint < R(0.5))) {
      vzint = a->vz * delta;
      vzint2 = a->vz * (1 - delta);
      yint2 = R(0.5) * diy;
      delta = (yint2 - yint) / (a->y1 - yint);
      xint2 = -xint + (a->x1 - xint) * delta;
      n = &v[(*np)++];
      n->x0 = -xint; n->x1 = xint2;
      n->y0 = yint;  n->y1 = yint2;
      n->ix = a->ix + dix; n->iy = a->iy;
      n->q = a->q;
      n->vz = vzint2 * delta;
      n = &v[(*np)++];
      n->x0 = xint2;  n->x1 = a->x1 - dix;
      n->y0 = -yint2; n->y1 = a->y1 - diy;
      n->ix = a->ix + dix; n->iy = a->iy + diy;
      n->q = a->q;
      n->vz = vzint2 * (1 - delta);
      a->x1 = xint; a->y1 = yint;
      a->vz = vzint;
    } else {
      vzint = a->vz * delta;
      yint2 = R(0.5) * diy;
      delta = (yint2 - a->y0) / (yint - a->y0);
      xint2 = LERP(a->x0, xint, delta);
      n = &v[(*np)++];
      n->x0 = xint2;  n->x1 = xint;
      n->y0 = -yint2; n->y1 = yint - diy;
      n->ix = a->ix; n->iy = a->iy + diy;
      n->q = a->q;
      n->vz = vzint * (1 - delta);
      n = &v[(*np)++];
      n->x0 = -xint;       n->x1 = a->x1 - dix;
      n->y0 = yint - diy;  n->y1 = a->y1 - diy;
      n->ix = a->ix + dix; n->iy = a->iy + diy;
      n->q = a->q;
      n->vz = a->vz - vzint;
      a->x1 = xint2; a->y1 = yint2;
      a->vz = vzint * delta;
    }
    break;
  }
}

/* helpers for the 3D splitter: a generic "cut along axis c" expressed on arrays so the
 * eight cases of split-vec.h:266-716 can be written once per axis pair. */
typedef struct {
  REAL p0[3], p1[3];
  REAL q;
  int i[3];
} vp3_t;

/* case with a single crossing on axis c (split-vec.h cases 1, 2, 4) */
static void split3d_one(vp3_t *v, int k, int *np, int c, const int *di) {
  vp3_t *a = &v[k], *n = &v[(*np)++];
  int o1 = (c + 1) % 3, o2 = (c + 2) % 3;
  REAL cint = R(0.5) * di[c];
  REAL delta = (cint - a->p0[c]) / (a->p1[c] - a->p0[c]);
  REAL i1 = LERP(a->p0[o1], a->p1[o1], delta);
  REAL i2 = LERP(a->p0[o2], a->p1[o2], delta);
  *n = *a;
  n->p0[c] = -cint; n->p1[c] = a->p1[c] - di[c];
  n->p0[o1] = i1;   n->p1[o1] = a->p1[o1];
  n->p0[o2] = i2;   n->p1[o2] = a->p1[o2];
  n->i[c] = a->i[c] + di[c];
  a->p1[c] = cint; a->p1[o1] = i1; a->p1[o2] = i2;
}

/* two crossings: first cut on axis c1, then axis c2 in whichever piece crosses it; the third
 * axis o is interpolated (split-vec.h cases 3 (x,y), 5 (x,z), 6 (y,z) and the first half of 7) */
static void split3d_two(vp3_t *v, int k, int *np, int c1, int c2, const int *di) {
  vp3_t *a = &v[k], *n;
  int o = 3 - c1 - c2;
  REAL aint = R(0.5) * di[c1];
  REAL delta = (aint - a->p0[c1]) / (a->p1[c1] - a->p0[c1]);
  REAL bint = LERP(a->p0[c2], a->p1[c2], delta);
  REAL oint = LERP(a->p0[o], a->p1[o], delta);
  if ((bint >= R(-0.5)) && (bint < R(0.5))) {
    /* c2 is crossed in the 2nd piece */
    REAL bint2 = R(0.5) * di[c2];
    delta = (bint2 - bint) / (a->p1[c2] - bint);
    REAL aint2 = -aint + (a->p1[c1] - aint) * delta;
    REAL oint2 = LERP(oint, a->p1[o], delta);
    n = &v[(*np)++];
    *n = *a;
    n->p0[c1] = -aint; n->p1[c1] = aint2;
    n->p0[c2] = bint;  n->p1[c2] = bint2;
    n->p0[o] = oint;   n->p1[o] = oint2;
    n->i[c1] = a->i[c1] + di[c1];
    n = &v[(*np)++];
    *n = *a;
    n->p0[c1] = aint2;  n->p1[c1] = a->p1[c1] - di[c1];
    n->p0[c2] = -bint2; n->p1[c2] = a->p1[c2] - di[c2];
    n->p0[o] = oint2;   n->p1[o] = a->p1[o];
    n->i[c1] = a->i[c1] + di[c1];
    n->i[c2] = a->i[c2] + di[c2];
    a->p1[c1] = aint; a->p1[c2] = bint; a->p1[o] = oint;
  } else {
    /* c2 is crossed in the 1st piece */
    REAL bint2 = R(0.5) * di[c2];
    delta = (bint2 - a->p0[c2]) / (bint - a->p0[c2]);
    REAL aint2 = LERP(a->p0[c1], aint, delta);
    REAL oint2 = LERP(a->p0[o], oint, delta);
    n = &v[(*np)++];
    *n = *a;
    n->p0[c1] = aint2;  n->p1[c1] = aint;
    n->p0[c2] = -bint2; n->p1[c2] = bint - di[c2];
    n->p0[o] = oint2;   n->p1[o] = oint;
    n->i[c2] = a->i[c2] + di[c2];
    n = &v[(*np)++];
    *n = *a;
    n->p0[c1] = -aint;           n->p1[c1] = a->p1[c1] - di[c1];
    n->p0[c2] = bint - di[c2];   n->p1[c2] = a->p1[c2] - di[c2];
    n->p0[o] = oint;             n->p1[o] = a->p1[o];
    n->i[c1] = a->i[c1] + di[c1];
    n->i[c2] = a->i[c2] + di[c2];
    a->p1[c1] = aint2; a->p1[c2] = bint2; a->p1[o] = oint2;
  }
}

/* vsplit3D, simd/split-vec.h:266-716 (Fortran split_3d spec/os-spec-current.f03:895-1601) */
static void split3d(vp3_t *v, int k, int *np, int cross, const int *di) {
  switch (cross) {
  case 0: break;
  case 1: split3d_one(v, k, np, 0, di); break;
  case 2: split3d_one(v, k, np, 1, di); break;
  case 4: split3d_one(v, k, np, 2, di); break;
  case 3: split3d_two(v, k, np, 0, 1, di); break;
  case 5: split3d_two(v, k, np, 0, 2, di); break;
  case 6: split3d_two(v, k, np, 1, 2, di); break;
  case 7: {
    split3d_two(v, k, np, 0, 1, di);
    int vpidx[3] = {k, *np - 2, *np - 1};
    REAL zint = R(0.5) * di[2];
    for (int i = 0; i < 3; i++) {
      vp3_t *a = &v[vpidx[i]];
      if ((a->p1[2] < R(-0.5)) || (a->p1[2] >= R(0.5))) {
        REAL delta = (zint - a->p0[2]) / (a->p1[2] - a->p0[2]);
        REAL xint = LERP(a->p0[0], a->p1[0], delta);
        REAL yint = LERP(a->p0[1], a->p1[1], delta);
        vp3_t *n = &v[(*np)++];
        *n = *a;
        n->p0[0] = xint;  n->p1[0] = a->p1[0];
        n->p0[1] = yint;  n->p1[1] = a->p1[1];
        n->p0[2] = -zint; n->p1[2] = a->p1[2] - di[2];
        n->i[2] = a->i[2] + di[2];
        a->p1[0] = xint; a->p1[1] = yint; a->p1[2] = zint;
        for (int j = i + 1; j < 3; j++) {
          vp3_t *b = &v[vpidx[j]];
          b->p0[2] -= di[2];
          b->p1[2] -= di[2];
          b->i[2] += di[2];
        }
        break;
      }
    }
  } break;
  }
}

/* ------------------------------------------------------------------------------------
 * Boris, simd/os-spec-push-avx2.c:90-146 (VDUDT_BORIS) and :154-159 (VRGAMMA);
 * Fortran dudt_boris spec/os-spec-push.f03:356-632.
 * ---------------------------------------------------------------------------------- */
static void boris(REAL tem, REAL *e, REAL *b, REAL *u, REAL *ene) {
  const REAL c1 = R(1.0);
  REAL e1 = e[0] * tem, e2 = e[1] * tem, e3 = e[2] * tem;
  REAL ut1 = u[0] + e1, ut2 = u[1] + e2, ut3 = u[2] + e3;
  REAL utsq = FMA(ut3, ut3, FMA(ut2, ut2, ut1 * ut1));
  REAL gamma = SQRT(utsq + c1);
  *ene = utsq / (gamma + c1);
  REAL tem_gamma = tem / gamma;
  REAL b1 = b[0] * tem_gamma, b2 = b[1] * tem_gamma, b3 = b[2] * tem_gamma;
  REAL u1 = FMA(b3, ut2, ut1), u2 = FMA(b1, ut3, ut2), u3 = FMA(b2, ut1, ut3);
  u1 = FNMA(b2, ut3, u1);
  u2 = FNMA(b3, ut1, u2);
  u3 = FNMA(b1, ut2, u3);
  REAL otsq = FMA(b1, b1, c1);
  otsq = FMA(b2, b2, otsq);
  otsq = FMA(b3, b3, otsq);
  otsq = R(2.0) / otsq;
  b1 *= otsq; b2 *= otsq; b3 *= otsq;
  ut1 = FMA(b3, u2, ut1);
  ut2 = FMA(b1, u3, ut2);
  ut3 = FMA(b2, u1, ut3);
  ut1 = FNMA(b2, u3, ut1);
  ut2 = FNMA(b3, u1, ut2);
  ut3 = FNMA(b1, u2, ut3);
  u[0] = ut1 + e1;
  u[1] = ut2 + e2;
  u[2] = ut3 + e3;
}
static REAL rgamma(const REAL *u) {
  return R(1.0) / SQRT(FMA(u[2], u[2], FMA(u[1], u[1], FMA(u[0], u[0], R(1.0)))));
}

static int lp_of(int order) { return (order == 1) ? 0 : (order == 4 ? -2 : -1); }

/* ------------------------------------------------------------------------------------
 * 2D deposit of one group of segments, simd/os-spec-current-avx2.c:730-865
 * (Fortran DEP_CURRENT_2D spec/os-spec-current.f03:2691-2796).
 * j points at component 1 of cell (1,1); sj = row stride in REALs (3*S1).
 * ---------------------------------------------------------------------------------- */
static void dep2d(int order, REAL *j, long sj, const REAL *jnorm, const vp_t *v, int n) {
  int NP = order + 1, lp = lp_of(order);
  for (int s = 0; s < n; s++) {
    const vp_t *a = &v[s];
    REAL s0x[MAXNP], s1x[MAXNP], s0y[MAXNP], s1y[MAXNP], wl1[MAXNP], wl2[MAXNP];
    REAL wp1[MAXNP], wp2[MAXNP];
    spline(order, a->x0, s0x);
    spline(order, a->x1, s1x);
    spline(order, a->y0, s0y);
    spline(order, a->y1, s1y);
    REAL qnx = a->q * jnorm[0], qny = a->q * jnorm[1];
    REAL qvz = (a->q * a->vz) * (R(1.0) / R(3.0));
    wl(order, qnx, a->x0, a->x1, wl1); wl1[NP - 1] = R(0.0);
    wl(order, qny, a->y0, a->y1, wl2); wl2[NP - 1] = R(0.0);
    for (int k = 0; k < NP; k++) {
      wp1[k] = s0y[k] + s1y[k];
      wp2[k] = s0x[k] + s1x[k];
    }
    REAL *pj = j + 3 * (long)(a->ix - 1 + lp) + sj * (long)(a->iy - 1 + lp);
    for (int k2 = 0; k2 < NP; k2++)
      for (int k1 = 0; k1 < NP; k1++) {
        REAL j1 = wl1[k1] * wp1[k2];
        REAL j2 = wl2[k2] * wp2[k1];
        REAL s00 = s0x[k1] * s0y[k2];
        REAL tmp1 = FMA(s1y[k2], s1x[k1], s00);
        REAL s10 = s0x[k1] * s1y[k2];
        REAL tmp2 = FMA(s0y[k2], s1x[k1], s10);
        tmp1 = FMA(tmp2, R(0.5), tmp1);
        REAL j3 = qvz * tmp1;
        REAL *c = pj + 3 * k1 + sj * k2;
        c[0] += j1;
        c[1] += j2;
        c[2] += j3;
      }
  }
}

/* 3D deposit of one vector group (<= VECW segments), simd/os-spec-current-avx2.c:1147-1318
 * (Fortran DEP_CURRENT_3D spec/os-spec-current.f03:2800-2905).  Accumulation order is the
 * reference's: plane k3 outermost, then lane, then the plane's points. */
static void dep3d_group(int order, REAL *j, long sjy, long sjz, const REAL *jnorm,
                        const vp3_t *v, int n) {
  int NP = order + 1, lp = lp_of(order);
  REAL wl1[VECW][MAXNP], wl2[VECW][MAXNP], wl3[VECW][MAXNP];
  REAL wp1[VECW][MAXNP][MAXNP], wp2[VECW][MAXNP][MAXNP], wp3[VECW][MAXNP][MAXNP];
  for (int s = 0; s < n; s++) {
    const vp3_t *a = &v[s];
    REAL s0x[MAXNP], s1x[MAXNP], s0y[MAXNP], s1y[MAXNP], s0z[MAXNP], s1z[MAXNP];
    spline(order, a->p0[0], s0x); spline(order, a->p1[0], s1x);
    spline(order, a->p0[1], s0y); spline(order, a->p1[1], s1y);
    spline(order, a->p0[2], s0z); spline(order, a->p1[2], s1z);
    REAL qnx = a->q * jnorm[0], qny = a->q * jnorm[1], qnz = a->q * jnorm[2];
    wl(order, qnx, a->p0[0], a->p1[0], wl1[s]); wl1[s][NP - 1] = R(0.0);
    wl(order, qny, a->p0[1], a->p1[1], wl2[s]); wl2[s][NP - 1] = R(0.0);
    wl(order, qnz, a->p0[2], a->p1[2], wl3[s]); wl3[s][NP - 1] = R(0.0);
    for (int k2 = 0; k2 < NP; k2++)
      for (int k1 = 0; k1 < NP; k1++) {
        REAL t1, t2;
        t1 = s0y[k1] * s0z[k2]; t2 = s0y[k1] * s1z[k2];
        t1 = FMA(s1y[k1], s1z[k2], t1); t2 = FMA(s1y[k1], s0z[k2], t2);
        wp1[s][k2][k1] = FMA(R(0.5), t2, t1);
        t1 = s0x[k1] * s0z[k2]; t2 = s0x[k1] * s1z[k2];
        t1 = FMA(s1x[k1], s1z[k2], t1); t2 = FMA(s1x[k1], s0z[k2], t2);
        wp2[s][k2][k1] = FMA(R(0.5), t2, t1);
        t1 = s0x[k1] * s0y[k2]; t2 = s0x[k1] * s1y[k2];
        t1 = FMA(s1x[k1], s1y[k2], t1); t2 = FMA(s1x[k1], s0y[k2], t2);
        wp3[s][k2][k1] = FMA(R(0.5), t2, t1);
      }
  }
  for (int k3 = 0; k3 < NP; k3++)
    for (int s = 0; s < n; s++) {
      const vp3_t *a = &v[s];
      REAL *pj = j + 3 * (long)(a->i[0] - 1 + lp) + sjy * (long)(a->i[1] - 1 + lp) +
                 sjz * (long)(a->i[2] - 1 + lp + k3);
      for (int k2 = 0; k2 < NP; k2++)
        for (int k1 = 0; k1 < NP; k1++) {
          REAL *c = pj + 3 * k1 + sjy * k2;
          c[0] += wl1[s][k1] * wp1[s][k3][k2];
          c[1] += wl2[s][k2] * wp2[s][k3][k1];
          c[2] += wl3[s][k3] * wp3[s][k2][k1];
        }
    }
}

/* ------------------------------------------------------------------------------------
 * oracle_advance_deposit: interpolate + Boris + move + split + deposit.
 * simd/os-spec-push-avx2.c:471-760 (2D), :1141-1472 (3D); marshalling as
 * spec/os-spec-vpush.f03:117-178, :267-327.  Arguments mirror the reference's C ABI
 * (1-based inclusive i0..i1, raw Fortran arrays, sizes = padded extents, offsets = lower
 * guard cells) with scalars by value.  *ene accumulates sum(q*ene_p) in double.
 * Grouping into vectors of VECW and chunks of p_cache_size reproduces the reference's
 * summation order into J exactly (os-spec-push-avx2.h:6-8).
 * ---------------------------------------------------------------------------------- */
static const REAL *fld_at(const REAL *f, const int *size, const int *off, int dim, int i1,
                          int i2, int i3) {
  /* pointer to component 1 of cell (i1,i2,i3), 1-based cell indices */
  long s1 = size[0], s2 = size[1];
  long p = (i1 - 1 + off[0]) + s1 * (long)(i2 - 1 + off[1]);
  if (dim == 3) p += s1 * s2 * (long)(i3 - 1 + off[2]);
  return f + 3 * p;
}

void oracle_advance_deposit(int dim, int order, int *ix, REAL *x, REAL *u, const REAL *q,
                            int i0, int i1, REAL rqm, const REAL *efield, const REAL *bfield,
                            const int *emf_size, const int *emf_offset, REAL *j,
                            const int *j_size, const int *j_offset, const double *dx, double dt,
                            double *ene) {
  const int NP = order + 1, lp = lp_of(order);
  const REAL tem = (REAL)(0.5 * dt / rqm);
  REAL dt_dx[3], jnorm[3];
  for (int d = 0; d < dim; d++) {
    dt_dx[d] = (REAL)(dt / dx[d]);
    jnorm[d] = (REAL)(dx[d] / dt / (dim == 2 ? 2 : 3));
  }
  const int cache = (dim == 2) ? 464 : 160;
  const long np_total = (long)i1 - i0 + 1;
  const long esy = 3L * emf_size[0], esz = esy * emf_size[1];
  const long jsy = 3L * j_size[0], jsz = jsy * j_size[1];
  REAL *j11 = (REAL *)fld_at(j, j_size, j_offset, dim, 1, 1, 1);
  vp_t *v2 = NULL;
  vp3_t *v3 = NULL;
  if (dim == 2) v2 = (vp_t *)malloc(sizeof(vp_t) * 3 * (cache + VECW));
  else v3 = (vp3_t *)malloc(sizeof(vp3_t) * 4 * (cache + VECW));
  double etot = 0.0;

  for (long kc = 0; kc < np_total; kc += cache) {
    long npc = (np_total - kc < cache) ? np_total - kc : cache;
    int nv = 0; /* virtual particles in the chunk buffer */
    for (long g = 0; g < npc; g += VECW) {
      int ng = (npc - g < VECW) ? (int)(npc - g) : VECW;
      int base = nv; /* the group's originals sit at base..base+VECW-1, extras after */
      int cross[VECW], di[VECW][3];
      for (int l = 0; l < ng; l++) {
        long p = (i0 - 1) + kc + g + l;
        REAL *xp = x + dim * p, *up = u + 3 * p;
        int *ip = ix + dim * p;
        REAL w[3][MAXNP], wh[3][MAXNP];
        int ih[3];
        for (int d = 0; d < dim; d++) {
          spline(order, xp[d], w[d]);
          REAL h = (xp[d] < R(0.0)) ? R(1.0) : R(0.0);
          ih[d] = ip[d] - ((xp[d] < R(0.0)) ? 1 : 0);
          splineh(order, xp[d], h, wh[d]);
        }
        REAL e[3], b[3];
        if (dim == 2) {
          /* which axes use the staggered weights: E1 (h,.), E2 (.,h), E3 (.,.);
           * B1 (.,h), B2 (h,.), B3 (h,h)   -- os-spec-push-avx2.c:592-657 */
          static const int eh[3][2] = {{1, 0}, {0, 1}, {0, 0}};
          static const int bh[3][2] = {{0, 1}, {1, 0}, {1, 1}};
          for (int c = 0; c < 3; c++) {
            for (int fb = 0; fb < 2; fb++) {
              const int *hh = fb ? bh[c] : eh[c];
              const REAL *f = fb ? bfield : efield;
              const REAL *wx = hh[0] ? wh[0] : w[0], *wy = hh[1] ? wh[1] : w[1];
              int bi = (hh[0] ? ih[0] : ip[0]) + lp, bj = (hh[1] ? ih[1] : ip[1]) + lp;
              const REAL *f0 = fld_at(f, emf_size, emf_offset, 2, bi, bj, 1) + c;
              REAL acc = R(0.0);
              for (int k2 = 0; k2 < NP; k2++) {
                REAL line = R(0.0);
                for (int k1 = 0; k1 < NP; k1++) line = FMA(f0[3 * k1 + esy * k2], wx[k1], line);
                acc = FMA(line, wy[k2], acc);
              }
              if (fb) b[c] = acc; else e[c] = acc;
            }
          }
        } else {
          static const int eh[3][3] = {{1, 0, 0}, {0, 1, 0}, {0, 0, 1}};
          static const int bh[3][3] = {{0, 1, 1}, {1, 0, 1}, {1, 1, 0}};
          for (int c = 0; c < 3; c++) {
            for (int fb = 0; fb < 2; fb++) {
              const int *hh = fb ? bh[c] : eh[c];
              const REAL *f = fb ? bfield : efield;
              const REAL *wx = hh[0] ? wh[0] : w[0], *wy = hh[1] ? wh[1] : w[1];
              const REAL *wz = hh[2] ? wh[2] : w[2];
              int bi = (hh[0] ? ih[0] : ip[0]) + lp, bj = (hh[1] ? ih[1] : ip[1]) + lp;
              int bk = (hh[2] ? ih[2] : ip[2]) + lp;
              const REAL *f0 = fld_at(f, emf_size, emf_offset, 3, bi, bj, bk) + c;
              REAL acc = R(0.0);
              for (int k3 = 0; k3 < NP; k3++) {
                REAL plane = R(0.0);
                for (int k2 = 0; k2 < NP; k2++) {
                  REAL line = R(0.0);
                  for (int k1 = 0; k1 < NP; k1++)
                    line = FMA(f0[3 * k1 + esy * k2 + esz * k3], wx[k1], line);
                  plane = FMA(line, wy[k2], plane);
                }
                acc = FMA(plane, wz[k3], acc);
              }
              if (fb) b[c] = acc; else e[c] = acc;
            }
          }
        }
        REAL enep;
        boris(tem, e, b, up, &enep);
        REAL rg = rgamma(up);
        etot += (double)q[p] * (double)enep;
        REAL x1[3];
        int tr[3];
        for (int d = 0; d < dim; d++) {
          x1[d] = FMA(up[d] * rg, dt_dx[d], xp[d]);
          tr[d] = ntrim(x1[d]);
        }
        cross[l] = 0;
        for (int d = 0; d < dim; d++) {
          di[l][d] = tr[d];
          if (tr[d]) cross[l] |= (1 << d);
        }
        if (dim == 2) {
          vp_t *a = &v2[base + l];
          a->x0 = xp[0]; a->x1 = x1[0]; a->y0 = xp[1]; a->y1 = x1[1];
          a->q = q[p]; a->vz = up[2] * rg;
          a->ix = ip[0]; a->iy = ip[1];
        } else {
          vp3_t *a = &v3[base + l];
          for (int d = 0; d < 3; d++) { a->p0[d] = xp[d]; a->p1[d] = x1[d]; a->i[d] = ip[d]; }
          a->q = q[p];
        }
        for (int d = 0; d < dim; d++) {
          xp[d] = x1[d] - (REAL)tr[d];
          ip[d] += tr[d];
        }
      }
      /* the reference pads a partial last group with dummy particles (q = 0, cell 1) that
       * occupy vector slots (os-spec-push-avx2.c:1197-1210); extras are appended after the
       * group's VECW slots */
      for (int l = ng; l < VECW; l++) {
        if (dim == 2) {
          vp_t z; memset(&z, 0, sizeof z); z.ix = z.iy = z.iz = 1;
          v2[base + l] = z;
        } else {
          vp3_t z; memset(&z, 0, sizeof z); z.i[0] = z.i[1] = z.i[2] = 1;
          v3[base + l] = z;
        }
      }
      int np2 = base + VECW;
      for (int l = 0; l < ng; l++) {
        if (dim == 2) split2d(v2, base + l, &np2, cross[l], di[l][0], di[l][1]);
        else split3d(v3, base + l, &np2, cross[l], di[l]);
      }
      nv = np2;
    }
    /* deposit all virtual particles of the chunk, in buffer order, vector by vector */
    if (dim == 2) {
      dep2d(order, j11, jsy, jnorm, v2, nv);
    } else {
      for (int s = 0; s < nv; s += VECW) {
        int n = (nv - s < VECW) ? nv - s : VECW;
        dep3d_group(order, j11, jsy, jsz, jnorm, v3 + s, n);
      }
    }
  }
  *ene += etot;
  free(v2);
  free(v3);
}

/* ------------------------------------------------------------------------------------
 * oracle_accelerate_deposit: the beam-initialisation pusher, spec/os-spec-accel.f90:41-397
 * (accelerate_deposit -> accelerate_deposit_2d :230-313, accelerate_deposit_3d :316-397).
 * No fields: x1 advances by vdt_dx1 = u_accel dt/dx1 / sqrt(1 + u_accel^2) with
 * u_accel = u_frac * ufl(1) (use_particle_uacc: u_frac * p(1)), the other coordinates and
 * the momentum stay; the current is deposited with q * q_frac, dxi(2:3) = 0 and -- in 2D --
 * rg = 0 (no out-of-plane current).  u_frac, q_frac, ufl1 arrive in particle precision
 * like the Fortran locals.  The Fortran deposit (dep_current_2d/3d) is restated by the same
 * splitter and stencil sums as the SIMD path (equal up to summation order).  *ene
 * accumulates sum q u^2/(gamma+1) of the unchanged momenta (what get_energy recomputes after
 * p_ene_recalc, spec/os-spec-define.f03:1040-1052).
 * ---------------------------------------------------------------------------------- */
void oracle_accelerate_deposit(int dim, int order, int *ix, REAL *x, const REAL *u, const REAL *q,
                               long np, REAL u_frac, REAL q_frac, REAL ufl1, int use_particle_uacc,
                               REAL *j, const int *j_size, const int *j_offset, const double *dx,
                               double dt, double *ene) {
  REAL jnorm[3];
  for (int d = 0; d < dim; d++) jnorm[d] = (REAL)(dx[d] / dt / (dim == 2 ? 2 : 3));
  const REAL dt_dx1 = (REAL)(dt / dx[0]);
  const long jsy = 3L * j_size[0], jsz = jsy * j_size[1];
  REAL *j11 = (REAL *)fld_at(j, j_size, j_offset, dim, 1, 1, 1);
  double etot = 0.0;
  for (long p = 0; p < np; p++) {
    REAL *xp = x + dim * p;
    const REAL *up = u + 3 * p;
    int *ip = ix + dim * p;
    const REAL u_accel = use_particle_uacc ? u_frac * up[0] : u_frac * ufl1;
    const REAL vdt_dx1 = (u_accel * dt_dx1) / SQRT(R(1.0) + u_accel * u_accel);
    const REAL x1 = xp[0] + vdt_dx1;
    const int tr = ntrim(x1);
    const REAL u2 = up[0] * up[0] + up[1] * up[1] + up[2] * up[2];
    etot += (double)q[p] * (double)(u2 / (SQRT(u2 + R(1.0)) + R(1.0)));
    if (dim == 2) {
      vp_t v[3 + VECW];
      memset(v, 0, sizeof v);
      v[0].x0 = xp[0]; v[0].x1 = x1; v[0].y0 = xp[1]; v[0].y1 = xp[1];
      v[0].q = q[p] * q_frac; v[0].vz = R(0.0);
      v[0].ix = ip[0]; v[0].iy = ip[1];
      int n = 1;
      split2d(v, 0, &n, tr ? 1 : 0, tr, 0);
      dep2d(order, j11, jsy, jnorm, v, n);
    } else {
      vp3_t v[4 + VECW];
      memset(v, 0, sizeof v);
      for (int d = 0; d < 3; d++) { v[0].p0[d] = xp[d]; v[0].p1[d] = xp[d]; v[0].i[d] = ip[d]; }
      v[0].p1[0] = x1;
      v[0].q = q[p] * q_frac;
      int n = 1;
      const int di[3] = {tr, 0, 0};
      split3d(v, 0, &n, tr ? 1 : 0, di);
      dep3d_group(order, j11, jsy, jsz, jnorm, v, n);
    }
    xp[0] = x1 - (REAL)tr;
    ip[0] += tr;
  }
  *ene += etot;
}

/* ------------------------------------------------------------------------------------
 * Yee solver.  emf/solver/os-emf-solver-yee.f03:102-138 (dbdt_2d), :211-251 (dbdt_3d),
 * :287-326 (dedt_2d), :385-428 (dedt_3d), :542-568 (advance_yee).
 * Grids: f(3, S1, S2[, S3]), cell (1,1,1) at padded position gcl.
 * ---------------------------------------------------------------------------------- */
#define F2(f, c, i1, i2) (f)[(c) + 3 * ((long)((i1)-1 + gcl[0]) + S1 * (long)((i2)-1 + gcl[1]))]
#define F3(f, c, i1, i2, i3) \
  (f)[(c) + 3 * ((long)((i1)-1 + gcl[0]) + S1 * ((long)((i2)-1 + gcl[1]) + S2 * (long)((i3)-1 + gcl[2])))]

void oracle_dbdt(int dim, REAL *b, const REAL *e, const int *nx, const int *gcl, const int *gcu,
                 const double *dx, double dt) {
  const long S1 = nx[0] + gcl[0] + gcu[0], S2 = nx[1] + gcl[1] + gcu[1];
  const REAL dtdx1 = (REAL)(dt / dx[0]), dtdx2 = (REAL)(dt / dx[1]);
  if (dim == 2) {
    for (int i2 = 0; i2 <= nx[1] + 1; i2++)
      for (int i1 = 0; i1 <= nx[0] + 1; i1++) {
        F2(b, 0, i1, i2) = F2(b, 0, i1, i2) - dtdx2 * (F2(e, 2, i1, i2 + 1) - F2(e, 2, i1, i2));
        F2(b, 1, i1, i2) = F2(b, 1, i1, i2) + dtdx1 * (F2(e, 2, i1 + 1, i2) - F2(e, 2, i1, i2));
        F2(b, 2, i1, i2) = F2(b, 2, i1, i2) - dtdx1 * (F2(e, 1, i1 + 1, i2) - F2(e, 1, i1, i2)) +
                           dtdx2 * (F2(e, 0, i1, i2 + 1) - F2(e, 0, i1, i2));
      }
  } else {
    const REAL dtdx3 = (REAL)(dt / dx[2]);
    for (int i3 = 0; i3 <= nx[2] + 1; i3++)
      for (int i2 = 0; i2 <= nx[1] + 1; i2++)
        for (int i1 = 0; i1 <= nx[0] + 1; i1++) {
          F3(b, 0, i1, i2, i3) = F3(b, 0, i1, i2, i3) -
                                 dtdx2 * (F3(e, 2, i1, i2 + 1, i3) - F3(e, 2, i1, i2, i3)) +
                                 dtdx3 * (F3(e, 1, i1, i2, i3 + 1) - F3(e, 1, i1, i2, i3));
          F3(b, 1, i1, i2, i3) = F3(b, 1, i1, i2, i3) +
                                 dtdx1 * (F3(e, 2, i1 + 1, i2, i3) - F3(e, 2, i1, i2, i3)) -
                                 dtdx3 * (F3(e, 0, i1, i2, i3 + 1) - F3(e, 0, i1, i2, i3));
          F3(b, 2, i1, i2, i3) = F3(b, 2, i1, i2, i3) -
                                 dtdx1 * (F3(e, 1, i1 + 1, i2, i3) - F3(e, 1, i1, i2, i3)) +
                                 dtdx2 * (F3(e, 0, i1, i2 + 1, i3) - F3(e, 0, i1, i2, i3));
        }
  }
}

void oracle_dedt(int dim, REAL *e, const REAL *b, const REAL *jay, const int *nx, const int *gcl,
                 const int *gcu, const double *dx, double dt) {
  const long S1 = nx[0] + gcl[0] + gcu[0], S2 = nx[1] + gcl[1] + gcu[1];
  const REAL dtdx1 = (REAL)(dt / dx[0]), dtdx2 = (REAL)(dt / dx[1]), dtif = (REAL)dt;
  if (dim == 2) {
    for (int i2 = 1; i2 <= nx[1] + 1; i2++)
      for (int i1 = 1; i1 <= nx[0] + 1; i1++) {
        F2(e, 0, i1, i2) = F2(e, 0, i1, i2) - dtif * F2(jay, 0, i1, i2) +
                           dtdx2 * (F2(b, 2, i1, i2) - F2(b, 2, i1, i2 - 1));
        F2(e, 1, i1, i2) = F2(e, 1, i1, i2) - dtif * F2(jay, 1, i1, i2) -
                           dtdx1 * (F2(b, 2, i1, i2) - F2(b, 2, i1 - 1, i2));
        F2(e, 2, i1, i2) = F2(e, 2, i1, i2) - dtif * F2(jay, 2, i1, i2) +
                           dtdx1 * (F2(b, 1, i1, i2) - F2(b, 1, i1 - 1, i2)) -
                           dtdx2 * (F2(b, 0, i1, i2) - F2(b, 0, i1, i2 - 1));
      }
  } else {
    const REAL dtdx3 = (REAL)(dt / dx[2]);
    for (int i3 = 1; i3 <= nx[2] + 1; i3++)
      for (int i2 = 1; i2 <= nx[1] + 1; i2++)
        for (int i1 = 1; i1 <= nx[0] + 1; i1++) {
          F3(e, 0, i1, i2, i3) = F3(e, 0, i1, i2, i3) - dtif * F3(jay, 0, i1, i2, i3) +
                                 dtdx2 * (F3(b, 2, i1, i2, i3) - F3(b, 2, i1, i2 - 1, i3)) -
                                 dtdx3 * (F3(b, 1, i1, i2, i3) - F3(b, 1, i1, i2, i3 - 1));
          F3(e, 1, i1, i2, i3) = F3(e, 1, i1, i2, i3) - dtif * F3(jay, 1, i1, i2, i3) +
                                 dtdx3 * (F3(b, 0, i1, i2, i3) - F3(b, 0, i1, i2, i3 - 1)) -
                                 dtdx1 * (F3(b, 2, i1, i2, i3) - F3(b, 2, i1 - 1, i2, i3));
          F3(e, 2, i1, i2, i3) = F3(e, 2, i1, i2, i3) - dtif * F3(jay, 2, i1, i2, i3) +
                                 dtdx1 * (F3(b, 1, i1, i2, i3) - F3(b, 1, i1 - 1, i2, i3)) -
                                 dtdx2 * (F3(b, 0, i1, i2, i3) - F3(b, 0, i1, i2 - 1, i3));
        }
  }
}

/* advance_yee (no boundary hooks: periodic / inter-node boundaries do nothing inside the
 * solver, emf/boundary/os-emf-boundary.f90:840-1112 only act on physical walls) */
void oracle_yee_advance(int dim, REAL *e, REAL *b, const REAL *jay, const int *nx, const int *gcl,
                        const int *gcu, const double *dx, double dt) {
  oracle_dbdt(dim, b, e, nx, gcl, gcu, dx, dt / 2.0);
  oracle_dedt(dim, e, b, jay, nx, gcl, gcu, dx, dt);
  oracle_dbdt(dim, b, e, nx, gcl, gcu, dx, dt / 2.0);
}

/* ------------------------------------------------------------------------------------
 * Customized-stencil ("Fei") solver, emf/solver/os-emf-solver-fei.f03: dbdt_2d_fei :108-251,
 * dedt_2d_fei :253-404, dbdt_3d_fei :406-565, dedt_3d_fei :567-730, advance_fei :2331-2363;
 * no boundary tapering, no current filter.  coef_e(1:n) enters dB/dt, coef_b(1:n) dE/dt.
 * ---------------------------------------------------------------------------------- */
void oracle_fei_dbdt(int dim, REAL *b, const REAL *e, const int *nx, const int *gcl, const int *gcu,
                     const double *dx, double dt, int n, const REAL *coef_e) {
  const long S1 = nx[0] + gcl[0] + gcu[0], S2 = nx[1] + gcl[1] + gcu[1];
  const REAL dtdx1 = (REAL)(dt / dx[0]), dtdx2 = (REAL)(dt / dx[1]);
  const int lb = 1 - n, ub = nx[0] + n;
  if (dim == 2) {
    for (int i2 = 0; i2 <= nx[1] + 1; i2++)
      for (int i1 = lb; i1 <= ub; i1++) {
        REAL tmp;
        F2(b, 0, i1, i2) = F2(b, 0, i1, i2) - dtdx2 * (F2(e, 2, i1, i2 + 1) - F2(e, 2, i1, i2));
        tmp = R(0.0);
        for (int j = 1; j <= n; j++) tmp = tmp + coef_e[j - 1] * (F2(e, 2, i1 + j, i2) - F2(e, 2, i1 - j + 1, i2));
        F2(b, 1, i1, i2) = F2(b, 1, i1, i2) + dtdx1 * tmp;
        tmp = R(0.0);
        for (int j = 1; j <= n; j++) tmp = tmp + coef_e[j - 1] * (F2(e, 1, i1 + j, i2) - F2(e, 1, i1 - j + 1, i2));
        F2(b, 2, i1, i2) = F2(b, 2, i1, i2) - dtdx1 * tmp + dtdx2 * (F2(e, 0, i1, i2 + 1) - F2(e, 0, i1, i2));
      }
  } else {
    const REAL dtdx3 = (REAL)(dt / dx[2]);
    for (int i3 = 0; i3 <= nx[2] + 1; i3++)
      for (int i2 = 0; i2 <= nx[1] + 1; i2++)
        for (int i1 = lb; i1 <= ub; i1++) {
          REAL tmp;
          F3(b, 0, i1, i2, i3) = F3(b, 0, i1, i2, i3) - dtdx2 * (F3(e, 2, i1, i2 + 1, i3) - F3(e, 2, i1, i2, i3)) +
                                 dtdx3 * (F3(e, 1, i1, i2, i3 + 1) - F3(e, 1, i1, i2, i3));
          tmp = R(0.0);
          for (int j = 1; j <= n; j++)
            tmp = tmp + coef_e[j - 1] * (F3(e, 2, i1 + j, i2, i3) - F3(e, 2, i1 - j + 1, i2, i3));
          F3(b, 1, i1, i2, i3) = F3(b, 1, i1, i2, i3) + dtdx1 * tmp -
                                 dtdx3 * (F3(e, 0, i1, i2, i3 + 1) - F3(e, 0, i1, i2, i3));
          tmp = R(0.0);
          for (int j = 1; j <= n; j++)
            tmp = tmp + coef_e[j - 1] * (F3(e, 1, i1 + j, i2, i3) - F3(e, 1, i1 - j + 1, i2, i3));
          F3(b, 2, i1, i2, i3) = F3(b, 2, i1, i2, i3) - dtdx1 * tmp +
                                 dtdx2 * (F3(e, 0, i1, i2 + 1, i3) - F3(e, 0, i1, i2, i3));
        }
  }
}

void oracle_fei_dedt(int dim, REAL *e, const REAL *b, const REAL *jay, const int *nx, const int *gcl,
                     const int *gcu, const double *dx, double dt, int n, const REAL *coef_b) {
  const long S1 = nx[0] + gcl[0] + gcu[0], S2 = nx[1] + gcl[1] + gcu[1];
  const REAL dtdx1 = (REAL)(dt / dx[0]), dtdx2 = (REAL)(dt / dx[1]), dtif = (REAL)dt;
  const int lb = 2 - n, ub = nx[0] + n;
  if (dim == 2) {
    for (int i2 = 1; i2 <= nx[1] + 1; i2++)
      for (int i1 = lb; i1 <= ub; i1++) {
        REAL tmp;
        F2(e, 0, i1, i2) = F2(e, 0, i1, i2) - dtif * F2(jay, 0, i1, i2) + dtdx2 * (F2(b, 2, i1, i2) - F2(b, 2, i1, i2 - 1));
        tmp = R(0.0);
        for (int j = 1; j <= n; j++) tmp = tmp + coef_b[j - 1] * (F2(b, 2, i1 + j - 1, i2) - F2(b, 2, i1 - j, i2));
        F2(e, 1, i1, i2) = F2(e, 1, i1, i2) - dtif * F2(jay, 1, i1, i2) - dtdx1 * tmp;
        tmp = R(0.0);
        for (int j = 1; j <= n; j++) tmp = tmp + coef_b[j - 1] * (F2(b, 1, i1 + j - 1, i2) - F2(b, 1, i1 - j, i2));
        F2(e, 2, i1, i2) = F2(e, 2, i1, i2) - dtif * F2(jay, 2, i1, i2) - dtdx2 * (F2(b, 0, i1, i2) - F2(b, 0, i1, i2 - 1)) +
                           dtdx1 * tmp;
      }
  } else {
    const REAL dtdx3 = (REAL)(dt / dx[2]);
    for (int i3 = 1; i3 <= nx[2] + 1; i3++)
      for (int i2 = 1; i2 <= nx[1] + 1; i2++)
        for (int i1 = lb; i1 <= ub; i1++) {
          REAL tmp;
          F3(e, 0, i1, i2, i3) = F3(e, 0, i1, i2, i3) - dtif * F3(jay, 0, i1, i2, i3) +
                                 dtdx2 * (F3(b, 2, i1, i2, i3) - F3(b, 2, i1, i2 - 1, i3)) -
                                 dtdx3 * (F3(b, 1, i1, i2, i3) - F3(b, 1, i1, i2, i3 - 1));
          tmp = R(0.0);
          for (int j = 1; j <= n; j++)
            tmp = tmp + coef_b[j - 1] * (F3(b, 2, i1 + j - 1, i2, i3) - F3(b, 2, i1 - j, i2, i3));
          F3(e, 1, i1, i2, i3) = F3(e, 1, i1, i2, i3) - dtif * F3(jay, 1, i1, i2, i3) +
                                 dtdx3 * (F3(b, 0, i1, i2, i3) - F3(b, 0, i1, i2, i3 - 1)) - dtdx1 * tmp;
          tmp = R(0.0);
          for (int j = 1; j <= n; j++)
            tmp = tmp + coef_b[j - 1] * (F3(b, 1, i1 + j - 1, i2, i3) - F3(b, 1, i1 - j, i2, i3));
          F3(e, 2, i1, i2, i3) = F3(e, 2, i1, i2, i3) - dtif * F3(jay, 2, i1, i2, i3) -
                                 dtdx2 * (F3(b, 0, i1, i2, i3) - F3(b, 0, i1, i2 - 1, i3)) + dtdx1 * tmp;
        }
  }
}

void oracle_fei_advance(int dim, REAL *e, REAL *b, const REAL *jay, const int *nx, const int *gcl,
                        const int *gcu, const double *dx, double dt, int n, const REAL *coef_e,
                        const REAL *coef_b) {
  oracle_fei_dbdt(dim, b, e, nx, gcl, gcu, dx, dt / 2.0, n, coef_e);
  oracle_fei_dedt(dim, e, b, jay, nx, gcl, gcu, dx, dt, n, coef_b);
  oracle_fei_dbdt(dim, b, e, nx, gcl, gcu, dx, dt / 2.0, n, coef_e);
}

/* ------------------------------------------------------------------------------------
 * Single-rank periodic guard-cell update along one axis.
 * vdf/os-vdf-comm.f90:142-326 (update_periodic_vdf).  mode 0 = replace, 1 = add.
 * ---------------------------------------------------------------------------------- */
void oracle_update_periodic(int dim, REAL *f, int fdim, const int *nx, const int *gcl,
                            const int *gcu, int axis, int mode) {
  long S[3] = {1, 1, 1}, st[3];
  for (int d = 0; d < dim; d++) S[d] = nx[d] + gcl[d] + gcu[d];
  st[0] = fdim; st[1] = st[0] * S[0]; st[2] = st[1] * S[1];
  long shift = (long)nx[axis] * st[axis];
  /* ranges along the update axis, in padded coordinates (cell i -> i-1+gcl) */
  int lo[2] = {0, gcl[axis] - 1};                               /* cells 1-gcl .. 0 */
  int up[2] = {nx[axis] + gcl[axis], (int)S[axis] - 1};        /* cells nx+1 .. nx+gcu */
  for (int side = 0; side < 2; side++) {
    int a0 = side ? up[0] : lo[0], a1 = side ? up[1] : lo[1];
    long sh = side ? -shift : shift;
    long r[3][2] = {{0, S[0] - 1}, {0, S[1] - 1}, {0, S[2] - 1}};
    r[axis][0] = a0; r[axis][1] = a1;
    for (long i3 = r[2][0]; i3 <= r[2][1]; i3++)
      for (long i2 = r[1][0]; i2 <= r[1][1]; i2++)
        for (long i1 = r[0][0]; i1 <= r[0][1]; i1++) {
          REAL *p = f + i1 * st[0] + i2 * st[1] + i3 * st[2];
          for (int c = 0; c < fdim; c++) {
            if (mode == 0) p[c] = p[c + sh];
            else { REAL val = p[c] + p[c + sh]; p[c] = val; p[c + sh] = val; }
          }
        }
  }
}

/* ------------------------------------------------------------------------------------
 * Particle periodic wrap on a single rank: spec/os-spec-boundary.f03:1280-1290
 * (ix < 1 -> ix + nx ; ix > nx -> ix - nx), axis by axis.
 * ---------------------------------------------------------------------------------- */
void oracle_particles_periodic(int dim, int *ix, long np, const int *nx) {
  for (long p = 0; p < np; p++)
    for (int d = 0; d < dim; d++) {
      int *i = &ix[dim * p + d];
      if (*i < 1) *i += nx[d];
      else if (*i > nx[d]) *i -= nx[d];
    }
}

/* ------------------------------------------------------------------------------------
 * Counting sort key and stable permutation: spec/os-species.f03:445-505
 * (generate_sort_idx_{2,3}d): key = ix1 + nx1*(ix2-1) + nx1*nx2*(ix3-1), stable.
 * idx[p] = new (0-based) position of particle p.
 * ---------------------------------------------------------------------------------- */
void oracle_sort_idx(int dim, const int *ix, long np, const int *nx, int64_t *idx) {
  long ncell = 1;
  for (int d = 0; d < dim; d++) ncell *= nx[d];
  int64_t *count = (int64_t *)calloc(ncell + 1, sizeof(int64_t));
  int64_t *key = (int64_t *)malloc(sizeof(int64_t) * (np ? np : 1));
  for (long p = 0; p < np; p++) {
    int64_t k = (ix[dim * p] - 1) + (int64_t)nx[0] * (ix[dim * p + 1] - 1);
    if (dim == 3) k += (int64_t)nx[0] * nx[1] * (ix[dim * p + 2] - 1);
    key[p] = k;
    count[k + 1]++;
  }
  for (long c = 0; c < ncell; c++) count[c + 1] += count[c];
  for (long p = 0; p < np; p++) idx[p] = count[key[p]]++;
  free(count);
  free(key);
}

/* ------------------------------------------------------------------------------------
 * Field energy: 1/2 * prod(dx) * sum_interior f_c^2 per component, double accumulation.
 * emf/diagnostics/os-emf-diagnostics.f03:43-58, vdf/os-vdf-math.f90:625-736.
 * ---------------------------------------------------------------------------------- */
void oracle_field_energy(int dim, const REAL *f, const int *nx, const int *gcl, const int *gcu,
                         const double *dx, double *out3) {
  const long S1 = nx[0] + gcl[0] + gcu[0], S2 = nx[1] + gcl[1] + gcu[1];
  double acc[3] = {0, 0, 0}, dvol = 1.0;
  for (int d = 0; d < dim; d++) dvol *= dx[d];
  int n3 = (dim == 3) ? nx[2] : 1;
  for (int i3 = 1; i3 <= n3; i3++)
    for (int i2 = 1; i2 <= nx[1]; i2++)
      for (int i1 = 1; i1 <= nx[0]; i1++)
        for (int c = 0; c < 3; c++) {
          double v = (dim == 3) ? (double)F3(f, c, i1, i2, i3) : (double)F2(f, c, i1, i2);
          acc[c] += v * v;
        }
  for (int c = 0; c < 3; c++) out3[c] = 0.5 * dvol * acc[c];
}

/* ------------------------------------------------------------------------------------
 * Charge density with the same-order spline (used by the continuity invariant test:
 * (rho_new - rho_old)/dt + div J = 0).  Follows the reference's charge deposit
 * spec/os-spec-charge.f03 in statement: rho(i+k) += q * prod_d S_d(k_d).
 * ---------------------------------------------------------------------------------- */
void oracle_deposit_rho(int dim, int order, const int *ix, const REAL *x, const REAL *q, long np,
                        double *rho, const int *size, const int *off) {
  int NP = order + 1, lp = lp_of(order);
  long s1 = size[0], s2 = size[1];
  for (long p = 0; p < np; p++) {
    REAL w[3][MAXNP];
    for (int d = 0; d < dim; d++) spline(order, x[dim * p + d], w[d]);
    int n3 = (dim == 3) ? NP : 1;
    for (int k3 = 0; k3 < n3; k3++)
      for (int k2 = 0; k2 < NP; k2++)
        for (int k1 = 0; k1 < NP; k1++) {
          long c = (ix[dim * p] - 1 + off[0] + lp + k1) +
                   s1 * (long)(ix[dim * p + 1] - 1 + off[1] + lp + k2);
          double wv = (double)w[0][k1] * (double)w[1][k2];
          if (dim == 3) {
            c += s1 * s2 * (long)(ix[dim * p + 2] - 1 + off[2] + lp + k3);
            wv *= (double)w[2][k3];
          }
          rho[c] += (double)q[p] * wv;
        }
  }
}

/* ------------------------------------------------------------------------------------
 * Current smoothing (digital filter), next row 1 of SURVEY.md 8f.
 *
 * Kernel set-up: setup_smooth (vdf/os-vdf-smooth.f90:373-467), combine_kernel3 (:155-179),
 * digital_filter (:117-148), type/order rules of read_nml_smooth (:247-353).  All in double;
 * the result is stored in REAL (scoef is real(p_k_fld), :49) and THEN normalised in REAL
 * storage with a double running total (:451-463).
 * type: 0 none, 1 custom (swfj given), 2 binomial, 3 compensated, 4 5pass, 5 tri6, 6 tri8,
 * 11 digital.  scoef is [dim][2*max_order+1], centre at index max_order.
 * ---------------------------------------------------------------------------------- */
#define SM_ORDER_MAX 15
static void combine_kernel3(const double (*kern3)[3], int n, int nmax, double *out) {
  /* out has 2*nmax+1 entries, Fortran index i <-> out[i-1] */
  for (int i = 0; i < 2 * nmax + 1; i++) out[i] = 0.0;
  out[nmax] = 1.0;
  for (int k = 1; k <= n; k++) {
    double f1 = 0.0, f2;
    for (int i = (nmax + 1) - (k - 1); i <= (nmax + 1) + k; i++) {
      const double ci = out[i - 1];
      out[i - 2] = kern3[k - 1][2] * ci + out[i - 2];
      f2 = kern3[k - 1][0] * ci;
      out[i - 1] = kern3[k - 1][1] * ci + f1;
      f1 = f2;
    }
  }
}
static double besseli0(double x) { /* besseli0_rdoub, os-math.f90:450-475: polynomial fit; the
                                       literals without a d exponent are default (single) reals */
  const double ax = fabs(x);
  if (ax < 3.75) {
    const double y = (x / 3.75) * (x / 3.75);
    return 1.0 + y * ((double)3.5156229f + y * ((double)3.0899424f + y * ((double)1.2067492f +
           y * ((double)0.2659732f + y * (0.360768e-1 + y * 0.45813e-2)))));
  }
  const double y = 3.75 / ax;
  return (exp(ax) / sqrt(ax)) * ((double)0.39894228f + y * (0.1328592e-1 + y * (0.225319e-2 +
         y * (-0.157565e-2 + y * (0.916281e-2 + y * (-0.2057706e-1 + y * (0.2635537e-1 +
         y * (-0.1647633e-1 + y * 0.392377e-2))))))));
}
int oracle_smooth_order(int type, int order) { /* read_nml_smooth :280-338 */
  switch (type) {
  case 0: return 0;
  case 3: return order < 2 ? 2 : order;
  case 4: return 5;
  case 5: return 3;
  case 6: return 4;
  default: return order;
  }
}
void oracle_smooth_setup(int dim, const int *type, const int *order_in, const double *swfj,
                         const double *digital_fc, const double *digital_A, int *order_out,
                         int *max_order_out, REAL *scoef) {
  int max_order = 0;
  for (int d = 0; d < dim; d++) {
    order_out[d] = oracle_smooth_order(type[d], order_in[d]);
    if (order_out[d] > max_order) max_order = order_out[d];
  }
  *max_order_out = max_order;
  const int W = 2 * max_order + 1;
  for (int i = 0; i < dim * W; i++) scoef[i] = R(0.0);
  for (int d = 0; d < dim; d++) {
    const int n = order_out[d];
    if (n == 0) continue;
    double kern3[SM_ORDER_MAX][3], comb[2 * SM_ORDER_MAX + 3];
    REAL *s = scoef + d * W;
    int have = 1;
    switch (type[d]) {
    case 1: /* custom: swfj(3, p_order_max, p_x_dim) */
      for (int j = 0; j < n; j++)
        for (int c = 0; c < 3; c++) kern3[j][c] = swfj[(d * SM_ORDER_MAX + j) * 3 + c];
      break;
    case 2:
      for (int j = 0; j < n; j++) { kern3[j][0] = 1.0; kern3[j][1] = 2.0; kern3[j][2] = 1.0; }
      break;
    case 3: {
      for (int j = 0; j < n - 1; j++) { kern3[j][0] = 1.0; kern3[j][1] = 2.0; kern3[j][2] = 1.0; }
      /* comp = ( 4.0 + 2.0 * ( order - 1 ) ) / ( order - 1 ): default-real arithmetic (:418) */
      const double comp = (double)((4.0f + 2.0f * (float)(n - 1)) / (float)(n - 1));
      kern3[n - 1][0] = -1.0; kern3[n - 1][1] = comp; kern3[n - 1][2] = -1.0;
      break;
    }
    case 4:
      for (int j = 0; j < 4; j++) { kern3[j][0] = 1.0; kern3[j][1] = 2.0; kern3[j][2] = 1.0; }
      kern3[4][0] = -5.0; kern3[4][1] = 15.0; kern3[4][2] = -5.0;
      break;
    case 5: {
      static const double t6[7] = {0.015625, -0.09375, 0.234375, 0.6875, 0.234375, -0.09375, 0.015625};
      for (int i = 0; i < 7; i++) s[max_order - 3 + i] = R(t6[i]);
      have = 0;
      break;
    }
    case 6: {
      static const double t8[9] = {-0.00390625, 0.03125, -0.109375, 0.21875, 0.7265625,
                                   0.21875, -0.109375, 0.03125, -0.00390625};
      for (int i = 0; i < 9; i++) s[max_order - 4 + i] = R(t8[i]);
      have = 0;
      break;
    }
    case 11: {
      const double A = digital_A[d], fc = digital_fc[d], pi = 3.14159265358979323846264;
      double alpha;
      if (A <= 21.) alpha = 0.;
      else if (A >= 50.) alpha = 0.1102 * (A - 8.7);
      else alpha = 0.5842 * pow(A - 21., 0.4) + 0.07886 * (A - 21.);
      s[max_order] = R(fc);
      for (int i = 1; i <= n; i++) {
        const float r = (float)i / (float)n; /* real(i)/n is default real (:139) */
        const double v = besseli0(alpha * sqrt(1.0 - (double)(r * r))) / besseli0(alpha) *
                         sin(i * pi * fc) / (i * pi);
        s[max_order + i] = R(v);
        s[max_order - i] = R(v);
      }
      have = 0;
      break;
    }
    default: have = 0; break;
    }
    if (have) {
      combine_kernel3((const double(*)[3])kern3, n, max_order, comb);
      for (int i = 0; i < W; i++) s[i] = R(comb[i]);
    }
    /* normalise (:451-463) */
    double total = 0.0;
    for (int j = max_order - n; j <= max_order + n; j++) total = total + (double)s[j];
    for (int j = max_order - n; j <= max_order + n; j++) s[j] = R((double)s[j] / total);
  }
}

/* filter_2d / filter_3d (vdf/os-vdf-smooth.f90:639-828, 834-1105): separable, in place, axis by
 * axis, over the WHOLE extent of the other axes (guard cells included) and ext_lb+order ..
 * ext_ub-order along the axis; the sliding window o() holds the unfiltered values.  Sum order:
 * f = o(-n)*s(-n); f = f + o(k)*s(k), k = -n+1..n, separate multiply and add in REAL.
 * (The reference's z pass loads the new window value into o(:,nsm2) instead of o(:,nsm3),
 * :1032-1034 -- identical when order(2) == order(3), as in every shipped deck; this restatement
 * uses nsm3.)  f: (3, S1, S2[, S3]) with S = nx + gcl + gcu. */
void oracle_smooth(int dim, REAL *f, const int *nx, const int *gcl, const int *gcu,
                   const int *order, int max_order, const REAL *scoef) {
  long S[3] = {1, 1, 1}, st[3];
  for (int d = 0; d < dim; d++) S[d] = nx[d] + gcl[d] + gcu[d];
  st[0] = 3; st[1] = 3 * S[0]; st[2] = 3 * S[0] * S[1];
  const int W = 2 * max_order + 1;
  for (int ax = 0; ax < dim; ax++) {
    const int n = order[ax];
    if (n <= 0) continue;
    const REAL *s = scoef + ax * W + max_order; /* s[k], k = -n..n */
    const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
    REAL o[3][2 * SM_ORDER_MAX + 1];
    for (long j2 = 0; j2 < S[a2]; j2++)
      for (long j1 = 0; j1 < S[a1]; j1++) {
        REAL *line = f + j1 * st[a1] + j2 * st[a2];
        const long lb = n, ub = S[ax] - 1 - n; /* int_lb, int_ub in padded coordinates */
        for (int k = -n; k < n; k++)
          for (int c = 0; c < 3; c++) o[c][k + n] = line[(lb + k) * st[ax] + c];
        for (long i = lb; i <= ub; i++) {
          for (int c = 0; c < 3; c++) o[c][2 * n] = line[(i + n) * st[ax] + c];
          REAL v[3];
          for (int c = 0; c < 3; c++) v[c] = o[c][0] * s[-n];
          for (int k = -n + 1; k <= n; k++)
            for (int c = 0; c < 3; c++) v[c] = v[c] + o[c][k + n] * s[k];
          for (int k = 0; k < 2 * n; k++)
            for (int c = 0; c < 3; c++) o[c][k] = o[c][k + 1];
          for (int c = 0; c < 3; c++) line[i * st[ax] + c] = v[c];
        }
      }
  }
}

/* ------------------------------------------------------------------------------------
 * Moving window.
 * move_window_vdf (vdf/os-vdf.f03:56-205): along every axis with nmove > 0 the data slide down,
 * f(i) = f(i + nmove) for i = lbound .. ubound - nmove, and the uncovered cells are set to 0.
 * f: (fdim, S1, S2[, S3]) with S = nx + gcl + gcu.
 * ---------------------------------------------------------------------------------- */
void oracle_move_window_vdf(int dim, REAL *f, int fdim, const int *nx, const int *gcl, const int *gcu,
                            const int *nmove) {
  long S[3] = {1, 1, 1}, st[3];
  for (int d = 0; d < dim; d++) S[d] = nx[d] + gcl[d] + gcu[d];
  st[0] = fdim; st[1] = st[0] * S[0]; st[2] = st[1] * S[1];
  for (int ax = 0; ax < dim; ax++) {
    const int n = nmove[ax];
    if (n <= 0) continue;
    const int a1 = (ax + 1) % 3, a2 = (ax + 2) % 3;
    for (long j2 = 0; j2 < S[a2]; j2++)
      for (long j1 = 0; j1 < S[a1]; j1++) {
        REAL *line = f + j1 * st[a1] + j2 * st[a2];
        for (long i = 0; i < S[ax]; i++)
          for (int c = 0; c < fdim; c++)
            line[i * st[ax] + c] = (i + n < S[ax]) ? line[(i + n) * st[ax] + c] : R(0.0);
      }
  }
}

/* Species boundaries of a single rank, axis by axis, lower side then upper side
 * (update_boundary_species -> process_phys_boundary, spec/os-spec-boundary.f03:1099-1232, 1238-1325).
 * bc[side*3 + axis]: 1 periodic (ix -+ nx, :1280-1290), 2 absorbing (particle removed, :1293-1300),
 * 3 specular (p_d -> -p_d, ix -> 1 | nx, x -> -x, :1302-1325; the upper wall of a moving window,
 * :1276).  keep[k] = 0 marks a removed particle. */
void oracle_particles_bc(int dim, int *ix, REAL *x, REAL *p, long np, const int *nx, const int *bc,
                         unsigned char *keep) {
  for (long k = 0; k < np; k++) keep[k] = 1;
  for (int d = 0; d < dim; d++)
    for (long k = 0; k < np; k++) {
      if (!keep[k]) continue;
      int *i = &ix[dim * k + d];
      const int side = (*i < 1) ? 0 : (*i > nx[d] ? 1 : -1);
      if (side < 0) continue;
      switch (bc[side * 3 + d]) {
      case 1: *i += side ? -nx[d] : nx[d]; break;
      case 2: keep[k] = 0; break;
      case 3:
        p[3 * k + d] = -p[3 * k + d];
        *i = side ? nx[d] : 1;
        x[dim * k + d] = -x[dim * k + d];
        break;
      default: break;
      }
    }
}
