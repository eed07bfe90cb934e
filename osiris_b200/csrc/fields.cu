// fields.cu -- grid kernels: Yee FDTD advance, guard-cell updates (periodic in-rank copy/fold and
// slab pack/unpack for the NCCL exchange), host<->device layout conversion, field energy.
//
// Reference: source/emf/solver/os-emf-solver-yee.f03:102-138, 211-251, 287-326, 385-428, 542-568;
// source/vdf/os-vdf-comm.f90:142-326 (periodic), :332-473 (pack/unpack), :756-784 (ranges);
// source/emf/diagnostics/os-emf-diagnostics.f03:43-58 (energy).
//
// All arithmetic is written as separate multiplies and adds in the reference's order (the unit is
// compiled with -fmad=false), so E and B are bit-identical to the Fortran statement.
#include <utility>

#include "common.cuh"
#include <limits.h>

#include "fields.cuh"

namespace osgpu {

// ---------------------------------------------------------------------------------------------
// Yee: B -= dt * curl E over cells 0..nx+1 (dbdt_2d/3d_yee)
// ---------------------------------------------------------------------------------------------
template <int DIM, typename T>
__global__ void __launch_bounds__(128) yee_b_kernel(Field3<T> b, Field3<const T> e, Grid g, T dtdx1,
                                                    T dtdx2, T dtdx3) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;  // cell index 0..nx+1
  const int j = blockIdx.y, k = blockIdx.z;
  if (i > g.nx[0] + 1) return;
  const long long sy = g.P0, sz = (long long)g.P0 * g.S[1];
  const long long c =
      g.at(i - 1 + g.gl[0], j - 1 + g.gl[1], (DIM == 3) ? k - 1 + g.gl[2] : 0);
  const T e1 = e.c[0][c], e2 = e.c[1][c], e3 = e.c[2][c];
  if (DIM == 2) {
    b.c[0][c] = b.c[0][c] - dtdx2 * (e.c[2][c + sy] - e3);
    b.c[1][c] = b.c[1][c] + dtdx1 * (e.c[2][c + 1] - e3);
    b.c[2][c] = b.c[2][c] - dtdx1 * (e.c[1][c + 1] - e2) + dtdx2 * (e.c[0][c + sy] - e1);
  } else {
    b.c[0][c] = b.c[0][c] - dtdx2 * (e.c[2][c + sy] - e3) + dtdx3 * (e.c[1][c + sz] - e2);
    b.c[1][c] = b.c[1][c] + dtdx1 * (e.c[2][c + 1] - e3) - dtdx3 * (e.c[0][c + sz] - e1);
    b.c[2][c] = b.c[2][c] - dtdx1 * (e.c[1][c + 1] - e2) + dtdx2 * (e.c[0][c + sy] - e1);
  }
}

// Yee: E += dt * (curl B - J) over cells 1..nx+1 (dedt_2d/3d_yee)
template <int DIM, typename T>
__global__ void __launch_bounds__(128)
    yee_e_kernel(Field3<T> e, Field3<const T> b, Field3<const T> jay, Grid g, T dtdx1, T dtdx2,
                 T dtdx3, T dtif) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1;  // cell index 1..nx+1
  const int j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.nx[0] + 1) return;
  const long long sy = g.P0, sz = (long long)g.P0 * g.S[1];
  const long long c =
      g.at(i - 1 + g.gl[0], j - 1 + g.gl[1], (DIM == 3) ? k - 1 + g.gl[2] : 0);
  const T b1 = b.c[0][c], b2 = b.c[1][c], b3 = b.c[2][c];
  if (DIM == 2) {
    e.c[0][c] = e.c[0][c] - dtif * jay.c[0][c] + dtdx2 * (b3 - b.c[2][c - sy]);
    e.c[1][c] = e.c[1][c] - dtif * jay.c[1][c] - dtdx1 * (b3 - b.c[2][c - 1]);
    e.c[2][c] = e.c[2][c] - dtif * jay.c[2][c] + dtdx1 * (b2 - b.c[1][c - 1]) -
                dtdx2 * (b1 - b.c[0][c - sy]);
  } else {
    e.c[0][c] = e.c[0][c] - dtif * jay.c[0][c] + dtdx2 * (b3 - b.c[2][c - sy]) -
                dtdx3 * (b2 - b.c[1][c - sz]);
    e.c[1][c] = e.c[1][c] - dtif * jay.c[1][c] + dtdx3 * (b1 - b.c[0][c - sz]) -
                dtdx1 * (b3 - b.c[2][c - 1]);
    e.c[2][c] = e.c[2][c] - dtif * jay.c[2][c] + dtdx1 * (b2 - b.c[1][c - 1]) -
                dtdx2 * (b1 - b.c[0][c - sy]);
  }
}

template <typename T>
int yee_substep(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt, int sub,
                cudaStream_t st, long long *launches) {
  // advance_yee: B half step, E full step, B half step (os-emf-solver-yee.f03:542-568);
  // dtdx = real(dt/dx, p_k_fld) with the full-precision dt (:223-225, :398-401)
  const double dtb = dt / 2.0;
  const T hb1 = (T)(dtb / dx[0]), hb2 = (T)(dtb / dx[1]), hb3 = (g.dim == 3) ? (T)(dtb / dx[2]) : T(0);
  const T he1 = (T)(dt / dx[0]), he2 = (T)(dt / dx[1]), he3 = (g.dim == 3) ? (T)(dt / dx[2]) : T(0);
  const T dtif = (T)dt;
  Field3<const T> ce{{e.c[0], e.c[1], e.c[2]}}, cb{{b.c[0], b.c[1], b.c[2]}},
      cj{{jay.c[0], jay.c[1], jay.c[2]}};
  const dim3 blk(128);
  const dim3 gb((g.nx[0] + 2 + 127) / 128, g.nx[1] + 2, g.dim == 3 ? g.nx[2] + 2 : 1);
  const dim3 ge((g.nx[0] + 1 + 127) / 128, g.nx[1] + 1, g.dim == 3 ? g.nx[2] + 1 : 1);
  if (g.dim == 2) {
    if (sub != 1) yee_b_kernel<2, T><<<gb, blk, 0, st>>>(b, ce, g, hb1, hb2, hb3);
    else yee_e_kernel<2, T><<<ge, blk, 0, st>>>(e, cb, cj, g, he1, he2, he3, dtif);
  } else {
    if (sub != 1) yee_b_kernel<3, T><<<gb, blk, 0, st>>>(b, ce, g, hb1, hb2, hb3);
    else yee_e_kernel<3, T><<<ge, blk, 0, st>>>(e, cb, cj, g, he1, he2, he3, dtif);
  }
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template <typename T>
int yee_advance(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                cudaStream_t st, long long *launches) {
  for (int sub = 0; sub < 3; sub++) OSGPU_TRY(yee_substep<T>(g, e, b, jay, dx, dt, sub, st, launches));
  return OSGPU_OK;
}
template int yee_substep<float>(const Grid &, Field3<float>, Field3<float>, Field3<float>, const double *, double, int,
                                cudaStream_t, long long *);
template int yee_substep<double>(const Grid &, Field3<double>, Field3<double>, Field3<double>, const double *, double, int,
                                 cudaStream_t, long long *);
template int yee_advance<float>(const Grid &, Field3<float>, Field3<float>, Field3<float>,
                                const double *, double, cudaStream_t, long long *);
template int yee_advance<double>(const Grid &, Field3<double>, Field3<double>, Field3<double>,
                                 const double *, double, cudaStream_t, long long *);

// ---------------------------------------------------------------------------------------------
// Customized-stencil solver (dbdt_2d/3d_fei, dedt_2d/3d_fei, advance_fei: emf/solver/
// os-emf-solver-fei.f03:108-251, 253-404, 406-565, 567-730, 2331-2363), without boundary tapering and
// current filtering.  B over x1 cells 1-n .. nx+n, E over 2-n .. nx+n (the guard cells are advanced
// too: the stencil of the next half step reaches into them), the other axes as Yee.  The stencil sum
// is formed like the reference: tmp = 0, tmp = tmp + coef(j) * (f(i+..) - f(i-..)), j = 1..n.
// ---------------------------------------------------------------------------------------------
template <int DIM, typename T>
__global__ void __launch_bounds__(128)
    fei_b_kernel(Field3<T> b, Field3<const T> e, Grid g, T dtdx1, T dtdx2, T dtdx3, SolverCoef<T> sc) {
  const int n = sc.n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 1 - n;  // cell index 1-n .. nx+n
  const int j = blockIdx.y, k = blockIdx.z;
  if (i > g.nx[0] + n) return;
  const long long sy = g.P0, sz = (long long)g.P0 * g.S[1];
  const long long c = g.at(i - 1 + g.gl[0], j - 1 + g.gl[1], (DIM == 3) ? k - 1 + g.gl[2] : 0);
  T t3 = T(0), t2 = T(0);  // stencil sums of E3 (for B2) and E2 (for B3)
  for (int m = 1; m <= n; m++) {
    t3 = t3 + sc.ce[m - 1] * (e.c[2][c + m] - e.c[2][c - m + 1]);
    t2 = t2 + sc.ce[m - 1] * (e.c[1][c + m] - e.c[1][c - m + 1]);
  }
  const T e1 = e.c[0][c], e2 = e.c[1][c], e3 = e.c[2][c];
  if (DIM == 2) {
    b.c[0][c] = b.c[0][c] - dtdx2 * (e.c[2][c + sy] - e3);
    b.c[1][c] = b.c[1][c] + dtdx1 * t3;
    b.c[2][c] = b.c[2][c] - dtdx1 * t2 + dtdx2 * (e.c[0][c + sy] - e1);
  } else {
    b.c[0][c] = b.c[0][c] - dtdx2 * (e.c[2][c + sy] - e3) + dtdx3 * (e.c[1][c + sz] - e2);
    b.c[1][c] = b.c[1][c] + dtdx1 * t3 - dtdx3 * (e.c[0][c + sz] - e1);
    b.c[2][c] = b.c[2][c] - dtdx1 * t2 + dtdx2 * (e.c[0][c + sy] - e1);
  }
}

template <int DIM, typename T>
__global__ void __launch_bounds__(128)
    fei_e_kernel(Field3<T> e, Field3<const T> b, Field3<const T> jay, Grid g, T dtdx1, T dtdx2, T dtdx3, T dtif,
                 SolverCoef<T> sc) {
  const int n = sc.n;
  const int i = blockIdx.x * blockDim.x + threadIdx.x + 2 - n;  // cell index 2-n .. nx+n
  const int j = blockIdx.y + 1, k = blockIdx.z + 1;
  if (i > g.nx[0] + n) return;
  const long long sy = g.P0, sz = (long long)g.P0 * g.S[1];
  const long long c = g.at(i - 1 + g.gl[0], j - 1 + g.gl[1], (DIM == 3) ? k - 1 + g.gl[2] : 0);
  T t3 = T(0), t2 = T(0);  // stencil sums of B3 (for E2) and B2 (for E3)
  for (int m = 1; m <= n; m++) {
    t3 = t3 + sc.cb[m - 1] * (b.c[2][c + m - 1] - b.c[2][c - m]);
    t2 = t2 + sc.cb[m - 1] * (b.c[1][c + m - 1] - b.c[1][c - m]);
  }
  const T b1 = b.c[0][c], b2 = b.c[1][c], b3 = b.c[2][c];
  if (DIM == 2) {
    e.c[0][c] = e.c[0][c] - dtif * jay.c[0][c] + dtdx2 * (b3 - b.c[2][c - sy]);
    e.c[1][c] = e.c[1][c] - dtif * jay.c[1][c] - dtdx1 * t3;
    e.c[2][c] = e.c[2][c] - dtif * jay.c[2][c] - dtdx2 * (b1 - b.c[0][c - sy]) + dtdx1 * t2;
  } else {
    e.c[0][c] = e.c[0][c] - dtif * jay.c[0][c] + dtdx2 * (b3 - b.c[2][c - sy]) - dtdx3 * (b2 - b.c[1][c - sz]);
    e.c[1][c] = e.c[1][c] - dtif * jay.c[1][c] + dtdx3 * (b1 - b.c[0][c - sz]) - dtdx1 * t3;
    e.c[2][c] = e.c[2][c] - dtif * jay.c[2][c] - dtdx2 * (b1 - b.c[0][c - sy]) + dtdx1 * t2;
  }
}

template <typename T>
int fei_substep(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                const SolverCoef<T> &sc, int sub, cudaStream_t st, long long *launches) {
  const double dtb = dt / 2.0;
  const T hb1 = (T)(dtb / dx[0]), hb2 = (T)(dtb / dx[1]), hb3 = (g.dim == 3) ? (T)(dtb / dx[2]) : T(0);
  const T he1 = (T)(dt / dx[0]), he2 = (T)(dt / dx[1]), he3 = (g.dim == 3) ? (T)(dt / dx[2]) : T(0);
  const T dtif = (T)dt;
  Field3<const T> ce{{e.c[0], e.c[1], e.c[2]}}, cb{{b.c[0], b.c[1], b.c[2]}}, cj{{jay.c[0], jay.c[1], jay.c[2]}};
  const dim3 blk(128);
  const dim3 gb((g.nx[0] + 2 * sc.n + 127) / 128, g.nx[1] + 2, g.dim == 3 ? g.nx[2] + 2 : 1);
  const dim3 ge((g.nx[0] + 2 * sc.n - 1 + 127) / 128, g.nx[1] + 1, g.dim == 3 ? g.nx[2] + 1 : 1);
  if (g.dim == 2) {
    if (sub != 1) fei_b_kernel<2, T><<<gb, blk, 0, st>>>(b, ce, g, hb1, hb2, hb3, sc);
    else fei_e_kernel<2, T><<<ge, blk, 0, st>>>(e, cb, cj, g, he1, he2, he3, dtif, sc);
  } else {
    if (sub != 1) fei_b_kernel<3, T><<<gb, blk, 0, st>>>(b, ce, g, hb1, hb2, hb3, sc);
    else fei_e_kernel<3, T><<<ge, blk, 0, st>>>(e, cb, cj, g, he1, he2, he3, dtif, sc);
  }
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template <typename T>
int fei_advance(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                const SolverCoef<T> &sc, cudaStream_t st, long long *launches) {
  for (int sub = 0; sub < 3; sub++) OSGPU_TRY(fei_substep<T>(g, e, b, jay, dx, dt, sc, sub, st, launches));
  return OSGPU_OK;
}
template int fei_substep<float>(const Grid &, Field3<float>, Field3<float>, Field3<float>, const double *, double,
                                const SolverCoef<float> &, int, cudaStream_t, long long *);
template int fei_substep<double>(const Grid &, Field3<double>, Field3<double>, Field3<double>, const double *, double,
                                 const SolverCoef<double> &, int, cudaStream_t, long long *);
template int fei_advance<float>(const Grid &, Field3<float>, Field3<float>, Field3<float>, const double *, double,
                                const SolverCoef<float> &, cudaStream_t, long long *);
template int fei_advance<double>(const Grid &, Field3<double>, Field3<double>, Field3<double>, const double *, double,
                                 const SolverCoef<double> &, cudaStream_t, long long *);

// ---------------------------------------------------------------------------------------------
// Mirror walls.  One thread per point of the wall plane (all padded cells of the other two axes, as
// the reference loops over the whole extent of the t_vdf there), each walking the few guard cells of
// the three components along the wall axis.
// ---------------------------------------------------------------------------------------------
template <typename T> __global__ void __launch_bounds__(128) wall_kernel(Field3<T> f, Grid g, WallOp op) {
  const int a = op.axis;
  const int u = (a == 0) ? 1 : 0, v = (a == 2) ? 1 : 2;  // the other two axes
  const int iu = blockIdx.x * blockDim.x + threadIdx.x, iv = blockIdx.y;
  if (iu >= g.S[u] || iv >= g.S[v]) return;
  const long long stride = (a == 0) ? 1 : (a == 1 ? (long long)g.P0 : (long long)g.P0 * g.S[1]);
  int q[3] = {0, 0, 0};
  q[u] = iu;
  q[v] = iv;
  const long long base = g.at(q[0], q[1], q[2]) + (long long)(g.gl[a] - 1) * stride;  // cell 0 -> + i * stride
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const WallComp w = op.c[c];
    T *p = f.c[c] + base;
    if (op.fold) {
      for (int i = w.lo; i <= w.hi; i++) {
        const T src = p[i * stride];
        T *dst = p + (long long)(w.m - i) * stride;
        *dst = (w.s > 0) ? *dst + src : *dst - src;
      }
    } else {
      for (int i = w.lo; i <= w.hi; i++) {
        const T src = p[(long long)(w.m - i) * stride];
        p[i * stride] = (w.s > 0) ? src : -src;
      }
    }
    if (w.plane != INT_MIN) {
      T *pp = p + (long long)w.plane * stride;
      *pp = (w.plane_op == 1) ? T(0) : *pp * T(2);
    }
  }
}

// The reference's odd / even and PEC / PMC routines are one rule: a component sits at cell index i + o/2
// along the wall axis (E: o = 1 for the normal component, B: o = 1 for the tangential ones), the mirror
// plane at P/2 (wall on a cell edge, "odd": P = 2 | 2 nx + 2; through the cell centres, "even": P = 1 |
// 2 nx + 1), so cell i maps to m - i with m = P - o; the guard cells beyond the plane take s times their
// image, and a point on the plane (m even) is zero when s = -1.  PEC: normal E and tangential B even
// (s = +1), tangential E and normal B odd; PMC the other way round.  One exception is kept as the
// reference has it: pec_bc_b_*_odd starts the upper tangential loop at nx + 2 (os-emf-pec.f90:590-600).
WallOp field_wall_op(const Grid &g, int axis, int side, int is_b, int pmc, int even, const int gc[2][3]) {
  WallOp op;
  op.axis = axis;
  op.fold = 0;
  const int nx = g.nx[axis];
  const int P = side == 0 ? (even ? 1 : 2) : (even ? 2 * nx + 1 : 2 * nx + 2);
  for (int c = 0; c < 3; c++) {
    const bool normal = (c == axis);
    const int o = is_b ? !normal : normal;
    int s = is_b ? (normal ? -1 : +1) : (normal ? +1 : -1);
    if (pmc) s = -s;
    WallComp w;
    w.m = P - o;
    w.s = s;
    w.plane = INT_MIN;
    w.plane_op = 0;
    const bool on_plane = (w.m % 2 == 0);
    if (side == 0) {
      w.lo = 1 - gc[0][axis];
      w.hi = on_plane ? w.m / 2 - 1 : (w.m - 1) / 2;
    } else {
      w.hi = nx + gc[1][axis];
      w.lo = on_plane ? w.m / 2 + 1 : (w.m + 1) / 2;
      if (!pmc && !even && is_b && !normal) w.lo += 1;
    }
    if (on_plane && s < 0) { w.plane = w.m / 2; w.plane_op = 1; }
    op.c[c] = w;
  }
  return op;
}
// spec_bc_2d / spec_bc_3d: the same mirror (plane on the cell edge for odd interpolation orders, through
// the cell centres for even ones) applied as a fold: what was deposited beyond the wall is added to its
// image with s = -1 for the component along the wall normal, +1 for the others; on the plane the normal
// component is zero and the others doubled.  Guard cells that can hold current: order/2 + 1 below the
// wall, (order+1)/2 + 1 above it (one less for the normal component, which sits half a cell higher).
WallOp current_wall_op(const Grid &g, int axis, int side, int order) {
  WallOp op;
  op.axis = axis;
  op.fold = 1;
  const int nx = g.nx[axis];
  const bool even = (order % 2 == 0);
  const int P = side == 0 ? (even ? 1 : 2) : (even ? 2 * nx + 1 : 2 * nx + 2);
  const int Lg = order / 2 + 1, Lu = (order + 1) / 2 + 1;
  for (int c = 0; c < 3; c++) {
    const bool normal = (c == axis);
    WallComp w;
    w.m = P - (normal ? 1 : 0);
    w.s = normal ? -1 : +1;
    w.plane = INT_MIN;
    w.plane_op = 0;
    const bool on_plane = (w.m % 2 == 0);
    if (side == 0) {
      w.lo = 1 - Lg;
      w.hi = on_plane ? w.m / 2 - 1 : (w.m - 1) / 2;
    } else {
      w.hi = nx + Lu - (normal ? 1 : 0);
      w.lo = on_plane ? w.m / 2 + 1 : (w.m + 1) / 2;
    }
    if (on_plane) { w.plane = w.m / 2; w.plane_op = normal ? 1 : 2; }
    op.c[c] = w;
  }
  return op;
}
template <typename T> int wall_apply(const Grid &g, Field3<T> f, const WallOp &op, cudaStream_t st, long long *launches) {
  const int a = op.axis;
  const int u = (a == 0) ? 1 : 0, v = (a == 2) ? 1 : 2;
  const dim3 grid((g.S[u] + 127) / 128, g.S[v]);
  wall_kernel<T><<<grid, 128, 0, st>>>(f, g, op);
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int wall_apply<float>(const Grid &, Field3<float>, const WallOp &, cudaStream_t, long long *);
template int wall_apply<double>(const Grid &, Field3<double>, const WallOp &, cudaStream_t, long long *);

// ---------------------------------------------------------------------------------------------
// Current smoothing: one pass of the separable digital filter along `axis`, src -> dst
// (filter_2d / filter_3d, source/vdf/os-vdf-smooth.f90:639-828, 834-1105).  The reference filters
// in place with a sliding window of unfiltered values; out of place that is dst = filter(src).
// Filtered range: box[axis] shrunk by n on both sides, whole box extent on the other axes (guard
// cells included); every other point is copied.  Sum order as the reference: f = o(-n)*s(-n),
// then f = f + o(k)*s(k) for k = -n+1..n, separate multiply and add (-fmad=false).
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
    smooth_axis_kernel(Field3<T> dst, Field3<const T> src, Grid g, SmoothBox bx, int axis, int n, SmoothCoef<T> sc) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= g.S[0]) return;
  const int i1 = blockIdx.y, i2 = blockIdx.z;
  const long long p = g.at(i0, i1, i2);
  const int i[3] = {i0, i1, i2};
  bool in = true;
#pragma unroll
  for (int d = 0; d < 3; d++) in &= (i[d] >= bx.lo[d]) && (i[d] <= bx.hi[d]);
  in &= (i[axis] >= bx.lo[axis] + n) && (i[axis] <= bx.hi[axis] - n);
  const long long stride = (axis == 0) ? 1 : (axis == 1 ? (long long)g.P0 : (long long)g.P0 * g.S[1]);
  const T *s = sc.s + SMOOTH_ORDER_MAX;  // s[k], k = -n..n
#pragma unroll
  for (int c = 0; c < 3; c++) {
    const T *q = src.c[c] + p;
    T v;
    if (in) {
      v = q[-n * stride] * s[-n];
      for (int k = -n + 1; k <= n; k++) v = v + q[k * stride] * s[k];
    } else {
      v = q[0];
    }
    dst.c[c][p] = v;
  }
}

// All passes of the filter: src -> dst (src == dst allowed: in place), scratch planes tmp.  The
// passes alternate between dst and tmp and are ordered so that the last one writes dst; with no
// active axis dst is a copy of src.  Every Field3 is three contiguous component planes.
template <typename T>
int smooth_field(const Grid &g, Field3<T> src, Field3<T> dst, Field3<T> tmp, const SmoothBox &bx, const int *order,
                 const SmoothCoef<T> *coef, cudaStream_t st, long long *launches) {
  int npass = 0;
  for (int ax = 0; ax < g.dim; ax++) npass += order[ax] > 0;
  if (npass == 0) {
    if (src.c[0] != dst.c[0])
      OSGPU_CUDA(cudaMemcpyAsync(dst.c[0], src.c[0], sizeof(T) * 3 * g.vol, cudaMemcpyDeviceToDevice, st));
    return OSGPU_OK;
  }
  const bool inplace = src.c[0] == dst.c[0];
  // targets of pass k = 0..npass-1: ..., tmp, dst, tmp, dst (last = dst); in place the first target
  // cannot be dst itself, so an odd pass count ends in tmp and is copied back
  const dim3 blk(128), grid((g.S[0] + 127) / 128, g.S[1], g.S[2]);
  Field3<T> cur = src;
  int k = 0;
  bool to_dst = inplace ? false : ((npass & 1) == 1);
  for (int ax = 0; ax < g.dim; ax++) {
    if (order[ax] <= 0) continue;
    Field3<T> out = to_dst ? dst : tmp;
    Field3<const T> ca{{cur.c[0], cur.c[1], cur.c[2]}};
    smooth_axis_kernel<T><<<grid, blk, 0, st>>>(out, ca, g, bx, ax, order[ax], coef[ax]);
    *launches += 1;
    cur = out;
    to_dst = !to_dst;
    k++;
  }
  OSGPU_CUDA(cudaGetLastError());
  if (cur.c[0] != dst.c[0])
    OSGPU_CUDA(cudaMemcpyAsync(dst.c[0], cur.c[0], sizeof(T) * 3 * g.vol, cudaMemcpyDeviceToDevice, st));
  return OSGPU_OK;
}
template int smooth_field<float>(const Grid &, Field3<float>, Field3<float>, Field3<float>, const SmoothBox &, const int *,
                                 const SmoothCoef<float> *, cudaStream_t, long long *);
template int smooth_field<double>(const Grid &, Field3<double>, Field3<double>, Field3<double>, const SmoothBox &, const int *,
                                  const SmoothCoef<double> *, cudaStream_t, long long *);

// ---------------------------------------------------------------------------------------------
// Moving window: f(i) = f(i + nmove) over the t_vdf's own bounds, the cells uncovered at the upper end
// set to zero (move_window_vdf, source/vdf/os-vdf.f03:56-205).  Out of place (src -> dst); points of
// the storage outside the box are copied.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128)
    shift_kernel(Field3<T> dst, Field3<const T> src, Grid g, SmoothBox bx, int n0, int n1, int n2) {
  const int i0 = blockIdx.x * blockDim.x + threadIdx.x;
  if (i0 >= g.S[0]) return;
  const int i1 = blockIdx.y, i2 = blockIdx.z;
  const long long p = g.at(i0, i1, i2);
  const bool in = i0 >= bx.lo[0] && i0 <= bx.hi[0] && i1 >= bx.lo[1] && i1 <= bx.hi[1] && i2 >= bx.lo[2] && i2 <= bx.hi[2];
  const bool have = (i0 + n0 <= bx.hi[0]) && (i1 + n1 <= bx.hi[1]) && (i2 + n2 <= bx.hi[2]);
  const long long q = g.at(i0 + n0, i1 + n1, i2 + n2);
#pragma unroll
  for (int c = 0; c < 3; c++) dst.c[c][p] = in ? (have ? src.c[c][q] : T(0)) : src.c[c][p];
}
template <typename T>
int shift_field(const Grid &g, Field3<T> f, Field3<T> tmp, const SmoothBox &bx, const int *nmove, cudaStream_t st,
                long long *launches) {
  const dim3 blk(128), grid((g.S[0] + 127) / 128, g.S[1], g.S[2]);
  Field3<const T> cf{{f.c[0], f.c[1], f.c[2]}};
  shift_kernel<T><<<grid, blk, 0, st>>>(tmp, cf, g, bx, nmove[0], g.dim > 1 ? nmove[1] : 0, g.dim > 2 ? nmove[2] : 0);
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  OSGPU_CUDA(cudaMemcpyAsync(f.c[0], tmp.c[0], sizeof(T) * 3 * g.vol, cudaMemcpyDeviceToDevice, st));
  return OSGPU_OK;
}
template int shift_field<float>(const Grid &, Field3<float>, Field3<float>, const SmoothBox &, const int *, cudaStream_t, long long *);
template int shift_field<double>(const Grid &, Field3<double>, Field3<double>, const SmoothBox &, const int *, cudaStream_t, long long *);

// ---------------------------------------------------------------------------------------------
// Slab helpers.  A slab is the padded range [a0, a1] along `axis` times the full padded extent of
// the other axes (that is how update_boundary_vdf propagates edges and corners, os-vdf-comm.f90
// :686-719).  r enumerates the slab with the x index fastest.
// ---------------------------------------------------------------------------------------------
struct Slab {
  int axis, a0, ext[3];
  long long vol;
};
static Slab make_slab(const Grid &g, int axis, int a0, int a1) {
  Slab s;
  s.axis = axis;
  s.a0 = a0;
  for (int d = 0; d < 3; d++) s.ext[d] = g.S[d];
  s.ext[axis] = a1 - a0 + 1;
  s.vol = (long long)s.ext[0] * s.ext[1] * s.ext[2];
  return s;
}
__device__ __forceinline__ long long slab_at(const Slab &s, const Grid &g, long long r) {
  int i[3];
  i[0] = (int)(r % s.ext[0]); r /= s.ext[0];
  i[1] = (int)(r % s.ext[1]);
  i[2] = (int)(r / s.ext[1]);
  i[s.axis] += s.a0;
  return g.at(i[0], i[1], i[2]);
}

template <typename T, int NC> struct Comps {
  T *c[NC];
};

// in-rank periodic update along one axis (update_periodic_vdf): mode 0 copies the opposite
// interior cells into the guard cells, mode 1 stores guard + image in both places.
template <typename T, int NC>
__global__ void __launch_bounds__(256)
    periodic_kernel(Comps<T, NC> f, Grid g, Slab lo, Slab up, long long shift, int mode) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (r < lo.vol) {
    const long long p = slab_at(lo, g, r);
    if (mode == 0) f.c[c][p] = f.c[c][p + shift];
    else { const T v = f.c[c][p] + f.c[c][p + shift]; f.c[c][p] = v; f.c[c][p + shift] = v; }
  } else if (r < lo.vol + up.vol) {
    const long long p = slab_at(up, g, r - lo.vol);
    if (mode == 0) f.c[c][p] = f.c[c][p - shift];
    else { const T v = f.c[c][p] + f.c[c][p - shift]; f.c[c][p - shift] = v; f.c[c][p] = v; }
  }
}

template <typename T>
int periodic_update(const Grid &g, T *const *comps, int ncomp, int axis, int mode, cudaStream_t st,
                    long long *launches) {
  if (g.nx[axis] < g.ul[axis] + g.uu[axis]) {
    set_error("periodic guard update needs nx >= gc_lower + gc_upper on axis %d", axis);
    return OSGPU_ERR_INVALID;
  }
  // lower guards: cells 1-gl..0 -> padded 0..gl-1; upper guards: cells nx+1..nx+gu
  const Slab lo = make_slab(g, axis, g.gl[axis] - g.ul[axis], g.gl[axis] - 1);
  const Slab up = make_slab(g, axis, g.nx[axis] + g.gl[axis], g.nx[axis] + g.gl[axis] + g.uu[axis] - 1);
  long long stride = 1;
  if (axis == 1) stride = g.P0;
  if (axis == 2) stride = (long long)g.P0 * g.S[1];
  const long long shift = stride * g.nx[axis];
  const long long tot = lo.vol + up.vol;
  if (tot == 0) return OSGPU_OK;
  const dim3 grid((unsigned)((tot + 255) / 256), ncomp);
  if (ncomp == 6) {
    Comps<T, 6> f;
    for (int c = 0; c < 6; c++) f.c[c] = comps[c];
    periodic_kernel<T, 6><<<grid, 256, 0, st>>>(f, g, lo, up, shift, mode);
  } else {
    Comps<T, 3> f;
    for (int c = 0; c < 3; c++) f.c[c] = comps[c];
    periodic_kernel<T, 3><<<grid, 256, 0, st>>>(f, g, lo, up, shift, mode);
  }
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int periodic_update<float>(const Grid &, float *const *, int, int, int, cudaStream_t, long long *);
template int periodic_update<double>(const Grid &, double *const *, int, int, int, cudaStream_t, long long *);

// pack / unpack a slab of all components into / from a contiguous message buffer
// (pack_vdf / unpack_vdf / unpack_add_vdf, os-vdf-comm.f90:332-473)
template <typename T, int NC>
__global__ void __launch_bounds__(256) pack_kernel(Comps<T, NC> f, Grid g, Slab s, T *buf, int dir) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const int c = blockIdx.y;
  if (r >= s.vol) return;
  const long long p = slab_at(s, g, r);
  T *m = buf + c * s.vol + r;
  if (dir == 0) *m = f.c[c][p];          // pack
  else if (dir == 1) f.c[c][p] = *m;     // unpack, replace
  else f.c[c][p] = f.c[c][p] + *m;       // unpack, add
}

template <typename T>
int slab_copy(const Grid &g, T *const *comps, int ncomp, int axis, int a0, int a1, T *buf, int dir,
              cudaStream_t st, long long *launches) {
  const Slab s = make_slab(g, axis, a0, a1);
  if (s.vol <= 0) return OSGPU_OK;
  const dim3 grid((unsigned)((s.vol + 255) / 256), ncomp);
  if (ncomp == 6) {
    Comps<T, 6> f;
    for (int c = 0; c < 6; c++) f.c[c] = comps[c];
    pack_kernel<T, 6><<<grid, 256, 0, st>>>(f, g, s, buf, dir);
  } else {
    Comps<T, 3> f;
    for (int c = 0; c < 3; c++) f.c[c] = comps[c];
    pack_kernel<T, 3><<<grid, 256, 0, st>>>(f, g, s, buf, dir);
  }
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int slab_copy<float>(const Grid &, float *const *, int, int, int, int, float *, int, cudaStream_t, long long *);
template int slab_copy<double>(const Grid &, double *const *, int, int, int, int, double *, int, cudaStream_t, long long *);

long long slab_volume(const Grid &g, int axis, int a0, int a1) {
  long long v = a1 - a0 + 1;
  for (int d = 0; d < 3; d++)
    if (d != axis) v *= g.S[d];
  return v < 0 ? 0 : v;
}

// ---------------------------------------------------------------------------------------------
// host t_vdf layout f(3, S1h, S2h, S3h) (component fastest, its own guard cells) <-> device planes
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256)
    layout_kernel(Field3<T> f, Grid g, T *aos, int hs0, int hs1, int hs2, int hg0, int hg1, int hg2,
                  int to_device) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const long long hvol = (long long)hs0 * hs1 * hs2;
  if (r >= hvol) return;
  const int i0 = (int)(r % hs0), i1 = (int)((r / hs0) % hs1), i2 = (int)(r / ((long long)hs0 * hs1));
  // host padded coordinate -> device padded coordinate
  const long long p = g.at(i0 - hg0 + g.gl[0], i1 - hg1 + g.gl[1], (g.dim == 3) ? i2 - hg2 + g.gl[2] : 0);
#pragma unroll
  for (int c = 0; c < 3; c++) {
    if (to_device) f.c[c][p] = aos[3 * r + c];
    else aos[3 * r + c] = f.c[c][p];
  }
}
template <typename T>
int layout_convert(const Grid &g, Field3<T> f, T *aos, const int *hsize, const int *hgc, int to_device,
                   cudaStream_t st, long long *launches) {
  const long long hvol = (long long)hsize[0] * hsize[1] * hsize[2];
  layout_kernel<T><<<(unsigned)((hvol + 255) / 256), 256, 0, st>>>(f, g, aos, hsize[0], hsize[1], hsize[2],
                                                                   hgc[0], hgc[1], hgc[2], to_device);
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int layout_convert<float>(const Grid &, Field3<float>, float *, const int *, const int *, int, cudaStream_t, long long *);
template int layout_convert<double>(const Grid &, Field3<double>, double *, const int *, const int *, int, cudaStream_t, long long *);

// ---------------------------------------------------------------------------------------------
// field energy: sum over interior cells of f_c^2, double accumulation
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) energy_kernel(Comps<T, 6> f, Grid g, double *out6) {
  const int c = blockIdx.y;
  const long long ncell = (long long)g.nx[0] * g.nx[1] * g.nx[2];
  double acc = 0.0;
  for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < ncell;
       r += (long long)gridDim.x * blockDim.x) {
    const int i0 = (int)(r % g.nx[0]), i1 = (int)((r / g.nx[0]) % g.nx[1]);
    const int i2 = (int)(r / ((long long)g.nx[0] * g.nx[1]));
    const double v = (double)f.c[c][g.at(i0 + g.gl[0], i1 + g.gl[1], (g.dim == 3) ? i2 + g.gl[2] : 0)];
    acc += v * v;
  }
  __shared__ double s[8];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_down_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) s[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int w = 0; w < 8; w++) t += s[w];
    atomicAdd(&out6[c], t);
  }
}
template <typename T>
int field_energy(const Grid &g, T *const *comps6, double *dev_out6, cudaStream_t st, long long *launches) {
  Comps<T, 6> f;
  for (int c = 0; c < 6; c++) f.c[c] = comps6[c];
  OSGPU_CUDA(cudaMemsetAsync(dev_out6, 0, 6 * sizeof(double), st));
  energy_kernel<T><<<dim3(296, 6), 256, 0, st>>>(f, g, dev_out6);
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int field_energy<float>(const Grid &, float *const *, double *, cudaStream_t, long long *);
template int field_energy<double>(const Grid &, double *const *, double *, cudaStream_t, long long *);

template <typename T> __global__ void __launch_bounds__(256) plane_add_kernel(T *dst, const T *src, long long n) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i < n) dst[i] = dst[i] + src[i];
}
template <typename T> int plane_add(const Grid &g, T *dst, const T *src, cudaStream_t st, long long *launches) {
  plane_add_kernel<T><<<(unsigned)((g.vol + 255) / 256), 256, 0, st>>>(dst, src, g.vol);
  *launches += 1;
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int plane_add<float>(const Grid &, float *, const float *, cudaStream_t, long long *);
template int plane_add<double>(const Grid &, double *, const double *, cudaStream_t, long long *);

}  // namespace osgpu
