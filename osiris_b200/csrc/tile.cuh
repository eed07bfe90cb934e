// tile.cuh -- cell-tile geometry shared by the kernels and the host-side layout code.
//
// Particles are binned into tiles of TX x TY (x TZ) cells (16 x 16 in 2D, 8 x 8 x 4 in 3D: 256 cells,
// one per thread of the CTA that owns the tile).  One CTA owns one tile: it stages
// the E/B values every particle of the tile can touch into shared memory (tile + stencil halo),
// accumulates the tile's current into a shared-memory J tile (tile + stencil halo + one cell of
// motion), and writes that back once.
#pragma once
#include <utility>

#include "pic_math.cuh"

namespace pic {

constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

#ifndef OSGPU_PAD_EY
#define OSGPU_PAD_EY 1
#endif
// smallest row count >= ey whose plane pitch (row_words * rows, in 4-byte banks) is 8..24 mod 32
constexpr int pad_tile_rows(int row_words, int ey) {
  return ((row_words * ey) % 32 >= 8 && (row_words * ey) % 32 <= 24) ? ey : pad_tile_rows(row_words, ey + 1);
}

template <int DIM, int ORDER, typename T> struct Tile {
  using S = Stencil<ORDER>;
  static constexpr int NP = S::NP, LP = S::LP;
  static constexpr int TX = (DIM == 2) ? 16 : 8;
  static constexpr int TY = (DIM == 2) ? 16 : 8;
  static constexpr int TZ = (DIM == 2) ? 1 : 4;
  static constexpr int NCELL = TX * TY * TZ;  // 256 in both geometries
  // E/B tile: interpolation points l+LP-1 .. l+UP for every local cell l in [0,T)
  static constexpr int EX = TX + NP, EZ = (DIM == 3) ? TZ + NP : 1;
  // inner extent padded so a TMA box row is a multiple of 16 bytes
  static constexpr int EXP = round_up(EX, 16 / (int)sizeof(T));
  // rows of the box: in 3D padded until the plane pitch is 8..24 banks (mod 32) away from a
  // multiple of the bank count.  The particles of a warp sit in a few consecutive cells of one row;
  // those whose half-cell shift selects the plane below would otherwise read the banks of their
  // neighbours (3D quadratic: 12 x 11 words = 4 banks; measured 1.35 wavefronts per load)
  static constexpr int EY = (DIM == 2 || !OSGPU_PAD_EY) ? TY + NP : pad_tile_rows(EXP * (int)sizeof(T) / 4, TY + NP);
  static constexpr int EVOL = EXP * EY * EZ;
  // per-component stride keeps every TMA destination 128-byte aligned
  static constexpr int ESTRIDE = round_up(EVOL * (int)sizeof(T), 128) / (int)sizeof(T);
  // J tile: deposit points of a segment in cell l + d, d in {-1,0,1}
  // (rows padded to 16 bytes, components to 128 bytes: the J tile leaves as TMA reduce boxes)
  static constexpr int JX = round_up(TX + NP + 1, 16 / (int)sizeof(T));
  static constexpr int JY = TY + NP + 1, JZ = (DIM == 3) ? TZ + NP + 1 : 1;
  static constexpr int JVOL = round_up(JX * JY * JZ * (int)sizeof(T), 128) / (int)sizeof(T);
  // first E/B (and J) tile point, in cells relative to the tile's first cell
  static constexpr int HALO_LO = LP - 1;
  static constexpr int SMEM_BYTES = (6 * ESTRIDE + 3 * JVOL) * (int)sizeof(T);
};

// tile edge used by the host-side layout code for a given dimensionality
inline int tile_edge(int dim, int axis) { return (dim == 2) ? (axis < 2 ? 16 : 1) : (axis < 2 ? 8 : 4); }
constexpr int TILE_NCELL = 256;

// packed per-particle cell word: (l_d + 1) in 10-bit fields, bit 31 = particle left the tile
constexpr int CELL_BITS = 10;
constexpr int CELL_MASK = (1 << CELL_BITS) - 1;
PIC_HD int cell_pack(int l0, int l1, int l2) {
  return (l0 + 1) | ((l1 + 1) << CELL_BITS) | ((l2 + 1) << (2 * CELL_BITS));
}
PIC_HD int cell_l(int cell, int d) { return ((cell >> (d * CELL_BITS)) & CELL_MASK) - 1; }
constexpr int CELL_DEAD = (int)0x80000000;

template <typename T> struct PushConst {
  T tem;        // 0.5 * dt / rqm          (simd/os-spec-push-avx2.c:1160)
  T dt_dx[3];   // dt / dx_d               (:1161-1163)
  T jnorm[3];   // dx_d / dt / 2 (2D) | 3 (3D)  (:1184-1186, :517-518)
};


// The NP^DIM tile values of one component around a particle, loaded back to back with constant
// offsets from one address.  On the device the loads are volatile asm statements: the compiler
// keeps them together ahead of the arithmetic instead of interleaving every load with its use,
// so the shared-memory latency is paid once per component and not once per line.
template <int BYTES> struct TileLd;
template <> struct TileLd<4> {
  template <int OFF> static PIC_HD float ld(const float *p) {
#if defined(__CUDA_ARCH__)
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"((unsigned)__cvta_generic_to_shared(p)), "n"(OFF * 4));
    return v;
#else
    return p[OFF];
#endif
  }
};
template <> struct TileLd<8> {
  template <int OFF> static PIC_HD double ld(const double *p) {
#if defined(__CUDA_ARCH__)
    double v;
    asm volatile("ld.shared.f64 %0, [%1+%2];" : "=d"(v) : "r"((unsigned)__cvta_generic_to_shared(p)), "n"(OFF * 8));
    return v;
#else
    return p[OFF];
#endif
  }
};
template <int DIM, int ORDER, typename T, int... I>
PIC_HD void tile_block_impl(const T *p, T *v, std::integer_sequence<int, I...>) {
  using G = Tile<DIM, ORDER, T>;
  constexpr int NP = G::NP;
  ((v[I] = TileLd<(int)sizeof(T)>::template ld<(I % NP) + G::EXP * (((I / NP) % NP) + G::EY * (I / (NP * NP)))>(p)), ...);
}
template <int DIM, int ORDER, typename T> PIC_HD void tile_block(const T *p, T *v) {
  constexpr int NP = Tile<DIM, ORDER, T>::NP;
  tile_block_impl<DIM, ORDER, T>(p, v, std::make_integer_sequence<int, (DIM == 3 ? NP * NP * NP : NP * NP)>{});
}

// ------------------------------------------------------------------------------------------
// fp32 on the device: the six components are interpolated two at a time with the packed
// fma.rn.f32x2 (SASS FFMA2, sm_100+): one instruction carries the same statement of two components
// -- each half rounds exactly like the scalar fma it replaces, so the result is bit-identical to the
// scalar nest below -- and halves the issue slots of the 39 (3D quadratic) fused multiply-adds per
// component.  A weight that both halves share is a scalar operand (FFMA2 broadcasts a .F32 register),
// a (plain, staggered) pair is kept as one register pair per axis and stencil point, and the
// (staggered, plain) order is the same pair read with the halves swapped (operand modifier).
// ------------------------------------------------------------------------------------------
#ifndef OSGPU_FFMA2
#define OSGPU_FFMA2 1
#endif
#if defined(__CUDA_ARCH__) && OSGPU_FFMA2
// SEL: 0 both halves take the plain weight, 1 both the staggered one, 2 (plain, staggered), 3 (staggered, plain)
template <int SEL> __device__ __forceinline__ float2 pair_weight(const float2 w) {
  if constexpr (SEL == 0) return make_float2(w.x, w.x);
  else if constexpr (SEL == 1) return make_float2(w.y, w.y);
  else if constexpr (SEL == 2) return w;
  else return make_float2(w.y, w.x);
}
// components A and B from their first stencil points pa, pb; wp[d][k] = (plain, staggered) weight of
// point k along axis d.  Returns (A, B).
template <int DIM, int ORDER, int SX, int SY, int SZ>
__device__ __forceinline__ float2 interp_pair(const float *pa, const float *pb,
                                              const float2 (*wp)[Stencil<ORDER>::NP]) {
  using G = Tile<DIM, ORDER, float>;
  constexpr int NP = G::NP;
  constexpr int NZ = (DIM == 3) ? NP : 1;
  float va[NP * NP * NZ], vb[NP * NP * NZ];
  tile_block<DIM, ORDER, float>(pa, va);
  tile_block<DIM, ORDER, float>(pb, vb);
  float2 acc = make_float2(0.0f, 0.0f);
#pragma unroll
  for (int k3 = 0; k3 < NZ; k3++) {
    float2 plane = make_float2(0.0f, 0.0f);
#pragma unroll
    for (int k2 = 0; k2 < NP; k2++) {
      float2 line = make_float2(0.0f, 0.0f);
#pragma unroll
      for (int k1 = 0; k1 < NP; k1++)
        line = __ffma2_rn(make_float2(va[k1 + NP * (k2 + NP * k3)], vb[k1 + NP * (k2 + NP * k3)]),
                          pair_weight<SX>(wp[0][k1]), line);
      plane = __ffma2_rn(line, pair_weight<SY>(wp[1][k2]), plane);
    }
    if constexpr (DIM == 3) acc = __ffma2_rn(plane, pair_weight<SZ>(wp[DIM - 1][k3]), acc);
    else acc = plane;
  }
  return acc;
}
#endif

// ------------------------------------------------------------------------------------------
// One particle, first half: interpolate E,B from the tile, Boris, move, trim.
template <int DIM, int ORDER, typename T>
PIC_HD T push_move(T *x, T *u, int *l, const T *eb, const PushConst<T> &k, T *x1, int *di, T *vz) {
  using G = Tile<DIM, ORDER, T>;
  constexpr int NP = G::NP;
  T w[DIM][NP], wh[DIM][NP];
  int b0[DIM], bh[DIM];  // first interpolation point in tile coordinates
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    spline<ORDER, T>(x[d], w[d]);
    const bool neg = x[d] < T(0.0);
    splineh<ORDER, T>(x[d], neg ? T(1.0) : T(0.0), wh[d]);
    b0[d] = l[d] + 1;
    bh[d] = l[d] + 1 - (neg ? 1 : 0);
  }
  T f[6];
#if defined(__CUDA_ARCH__) && OSGPU_FFMA2
  if constexpr (sizeof(T) == 4) {
    // (plain, staggered) weights of every axis from one packed evaluation of the spline: the halves are
    // spline(x) = spline_t(1/2 - x, 1/2 + x) and splineh(x, h) = spline_t((1 - h) - x, h + x)
    float2 wp[DIM][NP];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      const float hh = (x[d] < 0.0f) ? 1.0f : 0.0f;
      F2 wk[NP];
      spline_t<ORDER, F2>(F2(0.5f, 1.0f - hh) - F2(x[d]), F2(0.5f, hh) + F2(x[d]), wk);
#pragma unroll
      for (int k = 0; k < NP; k++) wp[d][k] = wk[k].v;
    }
    // first stencil point of a component with stagger flags (hx, hy, hz)
    auto at = [&](int c, bool hx, bool hy, bool hz) -> const float * {
      int o = (hx ? bh[0] : b0[0]) + G::EXP * (hy ? bh[1] : b0[1]);
      if constexpr (DIM == 3) o += G::EXP * G::EY * (hz ? bh[DIM - 1] : b0[DIM - 1]);
      return eb + c * G::ESTRIDE + o;
    };
    if constexpr (DIM == 2) {
      // E1 (h,.) E2 (.,h) E3 (.,.)  B1 (.,h) B2 (h,.) B3 (h,h)
      const float2 r0 = interp_pair<DIM, ORDER, 0, 2, 0>(at(2, false, false, false), at(1, false, true, false), wp);  // (E3, E2)
      const float2 r1 = interp_pair<DIM, ORDER, 1, 0, 0>(at(0, true, false, false), at(4, true, false, false), wp);   // (E1, B2)
      const float2 r2 = interp_pair<DIM, ORDER, 2, 1, 0>(at(3, false, true, false), at(5, true, true, false), wp);    // (B1, B3)
      f[2] = r0.x; f[1] = r0.y; f[0] = r1.x; f[4] = r1.y; f[3] = r2.x; f[5] = r2.y;
    } else {
      // E1 (h,.,.) E2 (.,h,.) E3 (.,.,h)  B1 (.,h,h) B2 (h,.,h) B3 (h,h,.)
      const float2 r0 = interp_pair<DIM, ORDER, 0, 2, 3>(at(2, false, false, true), at(1, false, true, false), wp);   // (E3, E2)
      const float2 r1 = interp_pair<DIM, ORDER, 1, 2, 3>(at(4, true, false, true), at(5, true, true, false), wp);     // (B2, B3)
      const float2 r2 = interp_pair<DIM, ORDER, 3, 2, 2>(at(0, true, false, false), at(3, false, true, true), wp);    // (E1, B1)
      f[2] = r0.x; f[1] = r0.y; f[4] = r1.x; f[5] = r1.y; f[0] = r2.x; f[3] = r2.y;
    }
  } else
#endif
  if constexpr (DIM == 2) {
    // staggering: E1 (h,.) E2 (.,h) E3 (.,.)  B1 (.,h) B2 (h,.) B3 (h,h)
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const bool hx = (c == 0) || (c == 4) || (c == 5);
      const bool hy = (c == 1) || (c == 3) || (c == 5);
      const T *wx = hx ? wh[0] : w[0], *wy = hy ? wh[1] : w[1];
      const T *p = eb + c * G::ESTRIDE + (hx ? bh[0] : b0[0]) + G::EXP * (hy ? bh[1] : b0[1]);
      T v[NP * NP];
      tile_block<DIM, ORDER, T>(p, v);
      T acc = T(0.0);
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++) {
        T line = T(0.0);
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) line = fmadd(v[k1 + NP * k2], wx[k1], line);
        acc = fmadd(line, wy[k2], acc);
      }
      f[c] = acc;
    }
  } else {
    // E1 (h,.,.) E2 (.,h,.) E3 (.,.,h)  B1 (.,h,h) B2 (h,.,h) B3 (h,h,.)
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const bool hx = (c == 0) || (c == 4) || (c == 5);
      const bool hy = (c == 1) || (c == 3) || (c == 5);
      const bool hz = (c == 2) || (c == 3) || (c == 4);
      const T *wx = hx ? wh[0] : w[0], *wy = hy ? wh[1] : w[1], *wz = hz ? wh[DIM - 1] : w[DIM - 1];
      const T *p = eb + c * G::ESTRIDE + (hx ? bh[0] : b0[0]) +
                   G::EXP * ((hy ? bh[1] : b0[1]) + G::EY * (hz ? bh[DIM - 1] : b0[DIM - 1]));
      T v[NP * NP * NP];
      tile_block<DIM, ORDER, T>(p, v);
      T acc = T(0.0);
#pragma unroll
      for (int k3 = 0; k3 < NP; k3++) {
        T plane = T(0.0);
#pragma unroll
        for (int k2 = 0; k2 < NP; k2++) {
          T line = T(0.0);
#pragma unroll
          for (int k1 = 0; k1 < NP; k1++)
            line = fmadd(v[k1 + NP * (k2 + NP * k3)], wx[k1], line);
          plane = fmadd(line, wy[k2], plane);
        }
        acc = fmadd(plane, wz[k3], acc);
      }
      f[c] = acc;
    }
  }
  const T ene = boris<T>(k.tem, f[0], f[1], f[2], f[3], f[4], f[5], u);
  const T rg = rgamma<T>(u);
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    x1[d] = fmadd(u[d] * rg, k.dt_dx[d], x[d]);
    di[d] = ntrim<T>(x1[d]);
    x[d] = x1[d] - T(di[d]);
    l[d] += di[d];
  }
  if constexpr (DIM == 2) *vz = u[2] * rg;
  return ene;
}

// The same particle statement with the fields read straight from the grids in global memory (runtime row /
// plane pitch): for the few particles that are pushed outside the cell tiles -- the cells a window move left
// below the box (push.cu: push_list_kernel).  c[] = padded grid coordinates of the particle's cell; fused
// multiply-adds exactly where push_move's scalar nest has them.
template <int DIM, int ORDER, typename T>
PIC_HD T push_move_grid(T *x, T *u, const int *c, const T *const *eb6, long long sy, long long sz, const PushConst<T> &k,
                        T *x1, int *di, T *vz) {
  constexpr int NP = Stencil<ORDER>::NP, LP = Stencil<ORDER>::LP;
  T w[DIM][NP], wh[DIM][NP];
  int b0[DIM], bh[DIM];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    spline<ORDER, T>(x[d], w[d]);
    const bool neg = x[d] < T(0.0);
    splineh<ORDER, T>(x[d], neg ? T(1.0) : T(0.0), wh[d]);
    b0[d] = c[d] + LP;
    bh[d] = c[d] + LP - (neg ? 1 : 0);
  }
  T f[6];
#pragma unroll
  for (int comp = 0; comp < 6; comp++) {
    bool hx, hy, hz;
    if (DIM == 2) {  // E1 (h,.) E2 (.,h) E3 (.,.)  B1 (.,h) B2 (h,.) B3 (h,h)
      hx = (comp == 0) || (comp == 4) || (comp == 5);
      hy = (comp == 1) || (comp == 3) || (comp == 5);
      hz = false;
    } else {         // E1 (h,.,.) E2 (.,h,.) E3 (.,.,h)  B1 (.,h,h) B2 (h,.,h) B3 (h,h,.)
      hx = (comp == 0) || (comp == 4) || (comp == 5);
      hy = (comp == 1) || (comp == 3) || (comp == 5);
      hz = (comp == 2) || (comp == 3) || (comp == 4);
    }
    const T *wx = hx ? wh[0] : w[0], *wy = hy ? wh[1] : w[1], *wz = hz ? wh[DIM - 1] : w[DIM - 1];
    const T *p = eb6[comp] + (hx ? bh[0] : b0[0]) + sy * (hy ? bh[1] : b0[1]) +
                 (DIM == 3 ? sz * (hz ? bh[DIM - 1] : b0[DIM - 1]) : 0);
    T acc = T(0.0);
#pragma unroll
    for (int k3 = 0; k3 < (DIM == 3 ? NP : 1); k3++) {
      T plane = T(0.0);
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++) {
        T line = T(0.0);
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) line = fmadd(p[k1 + sy * k2 + sz * k3], wx[k1], line);
        plane = fmadd(line, wy[k2], plane);
      }
      if (DIM == 3) acc = fmadd(plane, wz[k3], acc);
      else acc = plane;
    }
    f[comp] = acc;
  }
  const T ene = boris<T>(k.tem, f[0], f[1], f[2], f[3], f[4], f[5], u);
  const T rg = rgamma<T>(u);
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    x1[d] = fmadd(u[d] * rg, k.dt_dx[d], x[d]);
    di[d] = ntrim<T>(x1[d]);
    x[d] = x1[d] - T(di[d]);
  }
  if constexpr (DIM == 2) *vz = u[2] * rg;
  return ene;
}

// One particle of accelerate_deposit_{2d,3d} (spec/os-spec-accel.f90:230-397): x1 advances by
// vdt_dx1 = u_accel dt/dx1 / sqrt(1 + u_accel^2), the other coordinates and the momentum stay; the
// deposit sees no transverse motion (dxi(2:3) = 0, rg = 0 cancels the 2D out-of-plane current).
// Returns u^2/(gamma+1) of the (unchanged) momentum, the term get_energy recomputes when the push left
// p_ene_recalc (spec/os-spec-define.f03:1030-1086).
template <int DIM, typename T>
PIC_HD T accel_move(T *x, const T *u, int *l, T dt_dx1, T u_accel, T *x1, int *di, T *vz) {
  const T vdt_dx1 = (u_accel * dt_dx1) / Real<T>::sqrt(T(1.0) + u_accel * u_accel);
  x1[0] = x[0] + vdt_dx1;
  di[0] = ntrim<T>(x1[0]);
  x[0] = x1[0] - T(di[0]);
  l[0] += di[0];
#pragma unroll
  for (int d = 1; d < DIM; d++) {
    x1[d] = x[d];
    di[d] = 0;
  }
  if constexpr (DIM == 2) *vz = T(0.0);
  const T u2 = u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
  return u2 / (Real<T>::sqrt(u2 + T(1.0)) + T(1.0));
}

// Split a trajectory (seg[0]: p0, untrimmed p1, cell c, 2D vz) at the cell faces it crosses
// (simd/split-vec.h); di[] = ntrim(p1).  Returns the number of pieces.
template <int DIM, typename T> PIC_HD int split_trajectory(Seg<DIM, T> *seg, const int *di) {
  if constexpr (DIM == 2) return split2d<T>(seg, di[0], di[1]);
  else return split3d<T>(seg, di);
}

// push_move + split: seg[0..nseg) are the trajectory pieces (each inside one cell).
template <int DIM, int ORDER, typename T>
PIC_HD T push_one(T *x, T *u, int *l, const T *eb, const PushConst<T> &k, Seg<DIM, T> *seg,
                  int *nseg_out) {
  int di[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    seg[0].p0[d] = x[d];
    seg[0].c[d] = l[d];
  }
  T vz = T(0);
  const T ene = push_move<DIM, ORDER, T>(x, u, l, eb, k, seg[0].p1, di, &vz);
  if constexpr (DIM == 2) seg[0].vz = vz;
  *nseg_out = split_trajectory<DIM, T>(seg, di);
  return ene;
}

// Esirkepov weights of one segment (DEP_CURRENT_2D/3D, simd/os-spec-current-avx2.c:730-865,
// 1147-1318): calls out(comp, k1, k2, k3, f1, f2) for every non-zero stencil entry, whose value is
// f1 * f2; the entry belongs to grid point (cell + k + LP) of component comp.
// COMP >= 0 restricts the call to one current component (the weights of the others are never formed):
// the per-cell register accumulation of the 3D cubic scheme takes one component per pass.
template <int DIM, int ORDER, typename T, int COMP = -1, class Out>
PIC_HD void segment_current(const Seg<DIM, T> &a, T q, const PushConst<T> &k, Out out) {
  constexpr bool C0 = COMP < 0 || COMP == 0, C1 = COMP < 0 || COMP == 1, C2 = COMP < 0 || COMP == 2;
  constexpr int NP = Stencil<ORDER>::NP;
  {
  T s0[DIM][NP], s1[DIM][NP], wlv[DIM][NP];
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    spline<ORDER, T>(a.p0[d], s0[d]);
    spline<ORDER, T>(a.p1[d], s1[d]);
    wl<ORDER, T>(q * k.jnorm[d], a.p0[d], a.p1[d], wlv[d]);
  }
  if constexpr (DIM == 2) {
    const T qvz = (q * a.vz) * (T(1.0) / T(3.0));
#pragma unroll
    for (int k2 = 0; k2 < NP; k2++) {
      const T wp1 = s0[1][k2] + s1[1][k2];
#pragma unroll
      for (int k1 = 0; k1 < NP; k1++) {
        const T wp2 = s0[0][k1] + s1[0][k1];
        if (C0 && k1 < NP - 1) out(0, k1, k2, 0, wlv[0][k1], wp1);
        if (C1 && k2 < NP - 1) out(1, k1, k2, 0, wlv[1][k2], wp2);
        if (C2) {
          const T s00 = s0[0][k1] * s0[1][k2];
          T tmp1 = fmadd(s1[1][k2], s1[0][k1], s00);
          const T s10 = s0[0][k1] * s1[1][k2];
          const T tmp2 = fmadd(s0[1][k2], s1[0][k1], s10);
          tmp1 = fmadd(tmp2, T(0.5), tmp1);
          out(2, k1, k2, 0, qvz, tmp1);
        }
      }
    }
  } else {
    constexpr int Z = DIM - 1;  // == 2; written this way so the 2D instantiation compiles
    // wp?[kb][ka] as in the reference (wp1[kz][ky], wp2[kz][kx], wp3[ky][kx]):
    //   wp = S0a S0b + S1a S1b + (S0a S1b + S1a S0b) / 2 = S0a (S0b + S1b / 2) + S1a (S1b + S0b / 2)
    // the transverse factors in brackets are formed once per axis (same value up to rounding;
    // J is compared to the reference at 1e-5 max|J|, the summation order differs anyway)
    T hz0[NP], hz1[NP], hy0[NP], hy1[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) {
      hz0[k] = fmadd(T(0.5), s1[Z][k], s0[Z][k]);
      hz1[k] = fmadd(T(0.5), s0[Z][k], s1[Z][k]);
      hy0[k] = fmadd(T(0.5), s1[1][k], s0[1][k]);
      hy1[k] = fmadd(T(0.5), s0[1][k], s1[1][k]);
    }
    T wp1[NP][NP], wp2[NP][NP], wp3[NP][NP];
#pragma unroll
    for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
      for (int k1 = 0; k1 < NP; k1++) {
        wp1[k2][k1] = fmadd(s1[1][k1], hz1[k2], s0[1][k1] * hz0[k2]);
        wp2[k2][k1] = fmadd(s1[0][k1], hz1[k2], s0[0][k1] * hz0[k2]);
        wp3[k2][k1] = fmadd(s1[0][k1], hy1[k2], s0[0][k1] * hy0[k2]);
      }
#pragma unroll
    for (int k3 = 0; k3 < NP; k3++)
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) {
          if (C0 && k1 < NP - 1) out(0, k1, k2, k3, wlv[0][k1], wp1[k3][k2]);
          if (C1 && k2 < NP - 1) out(1, k1, k2, k3, wlv[1][k2], wp2[k3][k1]);
          if (C2 && k3 < NP - 1) out(2, k1, k2, k3, wlv[Z][k3], wp3[k2][k1]);
        }
  }
}
}

// the (comp, k1, k2, k3) of every non-zero stencil entry, in segment_current's order
template <int DIM, int ORDER, class F> PIC_HD void for_each_stencil_entry(F f) {
  constexpr int NP = Stencil<ORDER>::NP;
  constexpr int NZ = (DIM == 3) ? NP : 1;
#pragma unroll
  for (int k3 = 0; k3 < NZ; k3++)
#pragma unroll
    for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
      for (int k1 = 0; k1 < NP; k1++) {
        if (k1 < NP - 1) f(0, k1, k2, k3);
        if (k2 < NP - 1) f(1, k1, k2, k3);
        if (DIM == 2 || k3 < NP - 1) f(2, k1, k2, k3);
      }
}

// first J-tile element of the stencil of a segment in tile-local cell c[]
template <int DIM, int ORDER, typename T> PIC_HD int jtile_base(const int *c) {
  using G = Tile<DIM, ORDER, T>;
  if constexpr (DIM == 2) return (c[0] + 1) + G::JX * (c[1] + 1);
  else return (c[0] + 1) + G::JX * ((c[1] + 1) + G::JY * (c[DIM - 1] + 1));
}

// deposit one segment straight into the J tile; `add(ptr, v)` accumulates (shared-memory atomic
// on the device)
template <int DIM, int ORDER, typename T, class Add>
PIC_HD void deposit_segment(const Seg<DIM, T> &a, T q, T *jt, const PushConst<T> &k, Add add) {
  using G = Tile<DIM, ORDER, T>;
  T *pj = jt + jtile_base<DIM, ORDER, T>(a.c);
  segment_current<DIM, ORDER, T>(a, q, k, [&](int comp, int k1, int k2, int k3, T f1, T f2) {
    add(pj + comp * G::JVOL + k1 + G::JX * (k2 + G::JY * k3), f1 * f2);
  });
}

// deposit the segments selected by `mask`; the segment is picked with selects so that seg[] can
// stay in registers (no dynamically indexed local array)
template <int DIM, int ORDER, typename T, class Add>
PIC_HD void deposit_segments(const Seg<DIM, T> *seg, int nseg, unsigned mask, T q, T *jt,
                             const PushConst<T> &k, Add add) {
#pragma unroll 1
  for (int s = 0; s < nseg; s++) {
    if (!((mask >> s) & 1u)) continue;
    Seg<DIM, T> sg = seg[0];
#pragma unroll
    for (int m = 1; m <= DIM; m++)
      if (s == m) sg = seg[m];
    deposit_segment<DIM, ORDER, T>(sg, q, jt, k, add);
  }
}

// One particle, both halves (push_one + deposit of every segment).
template <int DIM, int ORDER, typename T, class Add>
PIC_HD T push_deposit_one(T *x, T *u, T q, int *l, const T *eb, T *jt, const PushConst<T> &k,
                          Add add) {
  Seg<DIM, T> seg[DIM + 1];
  int nseg;
  const T ene = push_one<DIM, ORDER, T>(x, u, l, eb, k, seg, &nseg);
  deposit_segments<DIM, ORDER, T>(seg, nseg, 0xf, q, jt, k, add);
  return ene;
}

}  // namespace pic
