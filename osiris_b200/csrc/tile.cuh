// tile.cuh -- cell-tile geometry shared by the kernels and the host-side layout code.
//
// Particles are binned into tiles of TX x TY (x TZ) cells.  One CTA owns one tile: it stages
// the E/B values every particle of the tile can touch into shared memory (tile + stencil halo),
// accumulates the tile's current into a shared-memory J tile (tile + stencil halo + one cell of
// motion), and writes that back once.
#pragma once
#include "pic_math.cuh"

namespace pic {

constexpr int round_up(int a, int b) { return (a + b - 1) / b * b; }

template <int DIM, int ORDER, typename T> struct Tile {
  using S = Stencil<ORDER>;
  static constexpr int NP = S::NP, LP = S::LP;
  static constexpr int TX = (DIM == 2) ? 16 : 8;
  static constexpr int TY = (DIM == 2) ? 16 : 8;
  static constexpr int TZ = (DIM == 2) ? 1 : 8;
  // E/B tile: interpolation points l+LP-1 .. l+UP for every local cell l in [0,T)
  static constexpr int EX = TX + NP, EY = TY + NP, EZ = (DIM == 3) ? TZ + NP : 1;
  // inner extent padded so a TMA box row is a multiple of 16 bytes
  static constexpr int EXP = round_up(EX, 16 / (int)sizeof(T));
  static constexpr int EVOL = EXP * EY * EZ;
  // per-component stride keeps every TMA destination 128-byte aligned
  static constexpr int ESTRIDE = round_up(EVOL * (int)sizeof(T), 128) / (int)sizeof(T);
  // J tile: deposit points of a segment in cell l + d, d in {-1,0,1}
  static constexpr int JX = TX + NP + 1, JY = TY + NP + 1, JZ = (DIM == 3) ? TZ + NP + 1 : 1;
  static constexpr int JVOL = JX * JY * JZ;
  // first E/B (and J) tile point, in cells relative to the tile's first cell
  static constexpr int HALO_LO = LP - 1;
  static constexpr int SMEM_BYTES = (6 * ESTRIDE + 3 * JVOL) * (int)sizeof(T);
};

// tile edge used by the host-side layout code for a given dimensionality
inline int tile_edge(int dim, int axis) { return (dim == 2) ? (axis < 2 ? 16 : 1) : 8; }

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

// ------------------------------------------------------------------------------------------
// One particle: interpolate E,B from the tile, Boris, move, trim, split, deposit into the J
// tile.  x[], u[] and l[] are advanced in place; l[] may leave [0,T) by one cell.
// Statement order follows ADVANCE_DEPOSIT_2D/3D (simd/os-spec-push-avx2.c:471-760, 1141-1472).
// `add(ptr, v)` accumulates into the J tile (shared-memory atomic on the device).
// ------------------------------------------------------------------------------------------
template <int DIM, int ORDER, typename T, class Add>
PIC_HD T push_deposit_one(T *x, T *u, T q, int *l, const T *eb, T *jt, const PushConst<T> &k,
                          Add add) {
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
  if constexpr (DIM == 2) {
    // staggering: E1 (h,.) E2 (.,h) E3 (.,.)  B1 (.,h) B2 (h,.) B3 (h,h)
#pragma unroll
    for (int c = 0; c < 6; c++) {
      const bool hx = (c == 0) || (c == 4) || (c == 5);
      const bool hy = (c == 1) || (c == 3) || (c == 5);
      const T *wx = hx ? wh[0] : w[0], *wy = hy ? wh[1] : w[1];
      const T *p = eb + c * G::ESTRIDE + (hx ? bh[0] : b0[0]) + G::EXP * (hy ? bh[1] : b0[1]);
      T acc = T(0.0);
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++) {
        T line = T(0.0);
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) line = fmadd(p[k1 + G::EXP * k2], wx[k1], line);
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
      T acc = T(0.0);
#pragma unroll
      for (int k3 = 0; k3 < NP; k3++) {
        T plane = T(0.0);
#pragma unroll
        for (int k2 = 0; k2 < NP; k2++) {
          T line = T(0.0);
#pragma unroll
          for (int k1 = 0; k1 < NP; k1++)
            line = fmadd(p[k1 + G::EXP * (k2 + G::EY * k3)], wx[k1], line);
          plane = fmadd(line, wy[k2], plane);
        }
        acc = fmadd(plane, wz[k3], acc);
      }
      f[c] = acc;
    }
  }
  const T ene = boris<T>(k.tem, f[0], f[1], f[2], f[3], f[4], f[5], u);
  const T rg = rgamma<T>(u);

  Seg<DIM, T> seg[DIM + 1];
  int di[3] = {0, 0, 0};
#pragma unroll
  for (int d = 0; d < DIM; d++) {
    const T x1 = fmadd(u[d] * rg, k.dt_dx[d], x[d]);
    di[d] = ntrim<T>(x1);
    seg[0].p0[d] = x[d];
    seg[0].p1[d] = x1;
    seg[0].c[d] = l[d];
    x[d] = x1 - T(di[d]);
    l[d] += di[d];
  }
  int nseg;
  if constexpr (DIM == 2) {
    seg[0].vz = u[2] * rg;
    nseg = split2d<T>(seg, di[0], di[1]);
  } else {
    nseg = split3d<T>(seg, di);
  }

  // deposit (DEP_CURRENT_2D/3D, simd/os-spec-current-avx2.c:730-865, 1147-1318)
  for (int s = 0; s < nseg; s++) {
    const Seg<DIM, T> &a = seg[s];
    T s0[DIM][NP], s1[DIM][NP], wlv[DIM][NP];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      spline<ORDER, T>(a.p0[d], s0[d]);
      spline<ORDER, T>(a.p1[d], s1[d]);
      wl<ORDER, T>(q * k.jnorm[d], a.p0[d], a.p1[d], wlv[d]);
    }
    if constexpr (DIM == 2) {
      const T qvz = (q * a.vz) * (T(1.0) / T(3.0));
      T *pj = jt + (a.c[0] + 1) + G::JX * (a.c[1] + 1);
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++) {
        const T wp1 = s0[1][k2] + s1[1][k2];
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) {
          const T wp2 = s0[0][k1] + s1[0][k1];
          T *c = pj + k1 + G::JX * k2;
          if (k1 < NP - 1) add(c, wlv[0][k1] * wp1);
          if (k2 < NP - 1) add(c + G::JVOL, wlv[1][k2] * wp2);
          const T s00 = s0[0][k1] * s0[1][k2];
          T tmp1 = fmadd(s1[1][k2], s1[0][k1], s00);
          const T s10 = s0[0][k1] * s1[1][k2];
          const T tmp2 = fmadd(s0[1][k2], s1[0][k1], s10);
          tmp1 = fmadd(tmp2, T(0.5), tmp1);
          add(c + 2 * G::JVOL, qvz * tmp1);
        }
      }
    } else {
      constexpr int Z = DIM - 1;  // == 2; written this way so the 2D instantiation compiles
      T *pj = jt + (a.c[0] + 1) + G::JX * ((a.c[1] + 1) + G::JY * (a.c[Z] + 1));
      // wp?[kb][ka] as in the reference: wp1[kz][ky], wp2[kz][kx], wp3[ky][kx]
      T wp1[NP][NP], wp2[NP][NP], wp3[NP][NP];
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) {
          T t1, t2;
          t1 = s0[1][k1] * s0[Z][k2]; t2 = s0[1][k1] * s1[Z][k2];
          t1 = fmadd(s1[1][k1], s1[Z][k2], t1); t2 = fmadd(s1[1][k1], s0[Z][k2], t2);
          wp1[k2][k1] = fmadd(T(0.5), t2, t1);
          t1 = s0[0][k1] * s0[Z][k2]; t2 = s0[0][k1] * s1[Z][k2];
          t1 = fmadd(s1[0][k1], s1[Z][k2], t1); t2 = fmadd(s1[0][k1], s0[Z][k2], t2);
          wp2[k2][k1] = fmadd(T(0.5), t2, t1);
          t1 = s0[0][k1] * s0[1][k2]; t2 = s0[0][k1] * s1[1][k2];
          t1 = fmadd(s1[0][k1], s1[1][k2], t1); t2 = fmadd(s1[0][k1], s0[1][k2], t2);
          wp3[k2][k1] = fmadd(T(0.5), t2, t1);
        }
#pragma unroll
      for (int k3 = 0; k3 < NP; k3++)
#pragma unroll
        for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
          for (int k1 = 0; k1 < NP; k1++) {
            T *c = pj + k1 + G::JX * (k2 + G::JY * k3);
            if (k1 < NP - 1) add(c, wlv[0][k1] * wp1[k3][k2]);
            if (k2 < NP - 1) add(c + G::JVOL, wlv[1][k2] * wp2[k3][k1]);
            if (k3 < NP - 1) add(c + 2 * G::JVOL, wlv[Z][k3] * wp3[k2][k1]);
          }
    }
  }
  return ene;
}

}  // namespace pic
