// pic_math.cuh -- per-particle arithmetic of the PIC hot path (splines, Boris, trajectory
// split, Esirkepov deposit weights).  Every fused multiply-add is explicit (the translation
// units are compiled with -fmad=false) and sits where the reference's FMA statement puts it
// (reference: source/simd/os-spec-push-avx2.c, os-spec-current-avx2.c, splines-avx2.h), so x, p
// and ix come out bit-identical to the reference's AVX2 kernels.
//
// The functions are __host__ __device__ so the same statements can be exercised on the host by
// the build-time self check (csrc/host_emul.cpp) where no GPU is available.
#pragma once
#include <math.h>

#if defined(__CUDACC__)
#define PIC_HD __host__ __device__ __forceinline__
#else
#define PIC_HD inline
#endif

namespace pic {

template <typename T> struct Real;
template <> struct Real<float> {
  static PIC_HD float fma(float a, float b, float c) { return ::fmaf(a, b, c); }
  static PIC_HD float sqrt(float a) { return ::sqrtf(a); }
};
template <> struct Real<double> {
  static PIC_HD double fma(double a, double b, double c) { return ::fma(a, b, c); }
  static PIC_HD double sqrt(double a) { return ::sqrt(a); }
};

#if defined(__CUDA_ARCH__)
// Two fp32 values that go through the same statements side by side (fma.rn.f32x2 / FFMA2 and its
// add / multiply forms, sm_100+).  Each half rounds exactly like the scalar operation, so a template
// written for T = float gives bit-identical halves when instantiated with T = F2.
struct F2 {
  float2 v;
  __device__ __forceinline__ F2() {}
  __device__ __forceinline__ F2(float a) : v(make_float2(a, a)) {}
  __device__ __forceinline__ F2(double a) : v(make_float2((float)a, (float)a)) {}
  __device__ __forceinline__ F2(float a, float b) : v(make_float2(a, b)) {}
  __device__ __forceinline__ F2(float2 a) : v(a) {}
};
__device__ __forceinline__ F2 operator-(F2 a) { return F2(-a.v.x, -a.v.y); }
__device__ __forceinline__ F2 operator+(F2 a, F2 b) { return F2(__fadd2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator-(F2 a, F2 b) { return F2(__fadd2_rn(a.v, (-b).v)); }
__device__ __forceinline__ F2 operator*(F2 a, F2 b) { return F2(__fmul2_rn(a.v, b.v)); }
__device__ __forceinline__ F2 operator/(F2 a, F2 b) { return F2(a.v.x / b.v.x, a.v.y / b.v.y); }  // (spline constants only)
template <> struct Real<F2> {
  static __device__ __forceinline__ F2 fma(F2 a, F2 b, F2 c) { return F2(__ffma2_rn(a.v, b.v, c.v)); }
};
#endif

template <typename T> PIC_HD T fmadd(T a, T b, T c) { return Real<T>::fma(a, b, c); }
template <typename T> PIC_HD T fnmadd(T a, T b, T c) { return Real<T>::fma(-a, b, c); }
template <typename T> PIC_HD T fmsub(T a, T b, T c) { return Real<T>::fma(a, b, -c); }

// interpolation stencil of an order: points LP..UP relative to the particle's cell
// (emf/os-emf-interpolate.f03:396-459; C OFFSET = 1 - LP, simd/os-spec-push-avx2.c:186-229)
template <int ORDER> struct Stencil {
  static constexpr int NP = ORDER + 1;
  static constexpr int LP = (ORDER == 1) ? 0 : (ORDER == 4 ? -2 : -1);
  static constexpr int UP = LP + ORDER;
};

// B-spline weights from t0 = 1/2 - x, t1 = 1/2 + x (simd/splines-avx2.h:27-193)
template <int ORDER, typename T> PIC_HD void spline_t(T t0, T t1, T *w) {
  if (ORDER == 1) {
    w[0] = t0;
    w[1] = t1;
  } else if (ORDER == 2) {
    w[0] = (t0 * t0) * T(0.5);
    w[1] = fmadd(t0, t1, T(0.5));
    w[2] = (t1 * t1) * T(0.5);
  } else if (ORDER == 3) {
    const T c1_2 = T(1.0) / T(2.0), c1_6 = T(1.0) / T(6.0), c2_3 = T(2.0) / T(3.0);
    T t2 = t0 * t0, t3 = t1 * t1;
    T a0 = t0 * t2, a1 = t1 * t3;
    w[0] = c1_6 * a0;
    w[1] = fmadd(c1_2, a1, c2_3 - t3);
    w[2] = fmadd(c1_2, a0, c2_3 - t2);
    w[3] = c1_6 * a1;
  } else {
    const T c1_2 = T(1.0) / T(2.0), c1_24 = T(1.0) / T(24.0), c1_4 = T(1.0) / T(4.0);
    const T c1 = T(1.0), c3_2 = T(3.0) / T(2.0), c1_6 = T(1.0) / T(6.0);
    const T c11_24 = T(11.0) / T(24.0);
    T t2 = t0 * t1, t02 = t0 * t0, t12 = t1 * t1;
    w[0] = c1_24 * (t02 * t02);
    w[1] = c1_6 * fmadd(t0, fmadd(t0, c3_2 + t2, c1), c1_4);
    w[2] = fmadd(t2, fmadd(c1_4, t2, c1_2), c11_24);
    w[3] = c1_6 * fmadd(t1, fmadd(t1, c3_2 + t2, c1), c1_4);
    w[4] = c1_24 * (t12 * t12);
  }
}
template <int ORDER, typename T> PIC_HD void spline(T x, T *w) {
  spline_t<ORDER, T>(T(0.5) - x, T(0.5) + x, w);
}
// staggered (half-cell) weights, h = (x < 0) (simd/os-spec-push-avx2.c:76-81)
template <int ORDER, typename T> PIC_HD void splineh(T x, T h, T *w) {
  spline_t<ORDER, T>((T(1.0) - h) - x, h + x, w);
}

// position trim to [-1/2, 1/2) (vmntrim, simd/os-spec-push-avx2.c:44-58)
template <typename T> PIC_HD int ntrim(T x) { return (x >= T(0.5)) - (x < T(-0.5)); }

// Boris rotation (VDUDT_BORIS, simd/os-spec-push-avx2.c:90-146); returns u^2/(gamma+1) of the
// time-centred momentum
template <typename T> PIC_HD T boris(T tem, T e1, T e2, T e3, T b1, T b2, T b3, T *u) {
  const T c1 = T(1.0);
  e1 *= tem; e2 *= tem; e3 *= tem;
  T ut1 = u[0] + e1, ut2 = u[1] + e2, ut3 = u[2] + e3;
  T utsq = fmadd(ut3, ut3, fmadd(ut2, ut2, ut1 * ut1));
  T gamma = Real<T>::sqrt(utsq + c1);
  T ene = utsq / (gamma + c1);
  T tem_gamma = tem / gamma;
  b1 *= tem_gamma; b2 *= tem_gamma; b3 *= tem_gamma;
  T u1 = fmadd(b3, ut2, ut1), u2 = fmadd(b1, ut3, ut2), u3 = fmadd(b2, ut1, ut3);
  u1 = fnmadd(b2, ut3, u1);
  u2 = fnmadd(b3, ut1, u2);
  u3 = fnmadd(b1, ut2, u3);
  T otsq = fmadd(b1, b1, c1);
  otsq = fmadd(b2, b2, otsq);
  otsq = fmadd(b3, b3, otsq);
  otsq = T(2.0) / otsq;
  b1 *= otsq; b2 *= otsq; b3 *= otsq;
  ut1 = fmadd(b3, u2, ut1);
  ut2 = fmadd(b1, u3, ut2);
  ut3 = fmadd(b2, u1, ut3);
  ut1 = fnmadd(b2, u3, ut1);
  ut2 = fnmadd(b3, u1, ut2);
  ut3 = fnmadd(b1, u2, ut3);
  u[0] = ut1 + e1;
  u[1] = ut2 + e2;
  u[2] = ut3 + e3;
  return ene;
}
// 1/gamma (VRGAMMA, simd/os-spec-push-avx2.c:154-159)
template <typename T> PIC_HD T rgamma(const T *u) {
  return T(1.0) / Real<T>::sqrt(fmadd(u[2], u[2], fmadd(u[1], u[1], fmadd(u[0], u[0], T(1.0)))));
}

// longitudinal deposit weights (vwl_s1..4, simd/os-spec-current-avx2.c:20-116); the last of the
// NP entries is zero (wl[NP-1] = 0, :801-803)
template <int ORDER, typename T> PIC_HD void wl(T qn, T x0, T x1, T *w) {
  if (ORDER == 1) {
    w[0] = qn * (x1 - x0);
  } else if (ORDER == 2) {
    T d = x1 - x0, s1_2 = T(0.5) - x0, p1_2 = T(0.5) + x0, n = qn * d;
    w[0] = n * fnmadd(T(0.5), d, s1_2);
    w[1] = n * fmadd(T(0.5), d, p1_2);
  } else if (ORDER == 3) {
    const T c1_2 = T(0.5), c3_4 = T(0.75), c1_3 = T(1.0) / T(3.0);
    T d = x1 - x0, s = c1_2 - x0, p = c1_2 + x0, d_3 = c1_3 * d;
    T w0 = c1_2 * fmsub(s, s, d * (s - d_3));
    T w1 = fnmadd(d, x0 + d_3, fnmadd(x0, x0, c3_4));
    T w2 = c1_2 * fmadd(p, p, d * (p + d_3));
    T n = qn * d;
    w[0] = n * w0;
    w[1] = n * w1;
    w[2] = n * w2;
  } else {
    const T c1_2 = T(0.5), c1_4 = T(0.25), c1_6 = T(1.0) / T(6.0), c3_2 = T(1.5);
    const T c3_4 = T(0.75), c1_3 = T(1.0) / T(3.0), c2_3 = T(2.0) / T(3.0);
    T d = x1 - x0, s = c1_2 - x0, p = c1_2 + x0;
    T t = x0 - c1_6, u = x0 + c1_6;
    T s2 = s * s, s3 = s2 * s, p2 = p * p, p3 = p2 * p;
    T d_2 = d * c1_2, d_4 = d * c1_4;
    T w0 = c1_6 * fmadd(d, fmsub(d, s - d_4, c3_2 * s2), s3);
    T w1 = fmadd(d, fmadd(d_2, t + d_4, fmsub(c3_4, t * t, c1_3)), fmadd(c1_2, p3, c2_3 - p2));
    T w2 = fnmadd(d, fmadd(d_2, u + d_4, fmsub(c3_4, u * u, c1_3)), fmadd(c1_2, s3, c2_3 - s2));
    T w3 = c1_6 * fmadd(d, fmadd(d, p + d_4, c3_2 * p2), p3);
    T n = qn * d;
    w[0] = n * w0;
    w[1] = n * w1;
    w[2] = n * w2;
    w[3] = n * w3;
  }
  w[ORDER] = T(0.0);
}

// A trajectory segment that stays inside one cell ("virtual particle",
// simd/os-spec-push-avx2.h:59-126).  c[] is the cell relative to the particle's tile.
template <int DIM, typename T> struct Seg {
  T p0[DIM], p1[DIM];
  T vz;  // 2D only: share of the out-of-plane velocity
  int c[DIM];
};

template <typename T> PIC_HD T lerp(T a, T b, T delta) { return a + (b - a) * delta; }

// 2D splitter (vsplit2D, simd/split-vec.h:105-254).  s[0] holds the whole trajectory with p1
// still in the start cell's frame; returns the number of segments (1..3).
template <typename T> PIC_HD int split2d(Seg<2, T> *s, int dix, int diy) {
  const int cross = (dix != 0) + 2 * (diy != 0);
  Seg<2, T> &a = s[0];
  if (cross == 0) return 1;
  if (cross == 1 || cross == 2) {
    const int c = cross - 1, o = 1 - c, di = c ? diy : dix;
    T cint = T(0.5) * di;
    T delta = (cint - a.p0[c]) / (a.p1[c] - a.p0[c]);
    T oint = lerp(a.p0[o], a.p1[o], delta);
    Seg<2, T> &n = s[1];
    n.p0[c] = -cint; n.p1[c] = a.p1[c] - di;
    n.p0[o] = oint;  n.p1[o] = a.p1[o];
    n.c[c] = a.c[c] + di; n.c[o] = a.c[o];
    n.vz = a.vz * (1 - delta);
    a.p1[c] = cint; a.p1[o] = oint;
    a.vz *= delta;
    return 2;
  }
  // x-y crossing: cut in x first
  T xint = T(0.5) * dix;
  T delta = (xint - a.p0[0]) / (a.p1[0] - a.p0[0]);
  T yint = lerp(a.p0[1], a.p1[1], delta);
  Seg<2, T> &n1 = s[1], &n2 = s[2];
  if ((yint >= T(-0.5)) && (yint < T(0.5))) {
    T vzint = a.vz * delta;
    T vzint2 = a.vz * (1 - delta);
    T yint2 = T(0.5) * diy;
    delta = (yint2 - yint) / (a.p1[1] - yint);
    T xint2 = -xint + (a.p1[0] - xint) * delta;
    n1.p0[0] = -xint; n1.p1[0] = xint2;
    n1.p0[1] = yint;  n1.p1[1] = yint2;
    n1.c[0] = a.c[0] + dix; n1.c[1] = a.c[1];
    n1.vz = vzint2 * delta;
    n2.p0[0] = xint2;  n2.p1[0] = a.p1[0] - dix;
    n2.p0[1] = -yint2; n2.p1[1] = a.p1[1] - diy;
    n2.c[0] = a.c[0] + dix; n2.c[1] = a.c[1] + diy;
    n2.vz = vzint2 * (1 - delta);
    a.p1[0] = xint; a.p1[1] = yint;
    a.vz = vzint;
  } else {
    T vzint = a.vz * delta;
    T yint2 = T(0.5) * diy;
    delta = (yint2 - a.p0[1]) / (yint - a.p0[1]);
    T xint2 = lerp(a.p0[0], xint, delta);
    n1.p0[0] = xint2;  n1.p1[0] = xint;
    n1.p0[1] = -yint2; n1.p1[1] = yint - diy;
    n1.c[0] = a.c[0]; n1.c[1] = a.c[1] + diy;
    n1.vz = vzint * (1 - delta);
    n2.p0[0] = -xint;       n2.p1[0] = a.p1[0] - dix;
    n2.p0[1] = yint - diy;  n2.p1[1] = a.p1[1] - diy;
    n2.c[0] = a.c[0] + dix; n2.c[1] = a.c[1] + diy;
    n2.vz = a.vz - vzint;
    a.p1[0] = xint2; a.p1[1] = yint2;
    a.vz = vzint * delta;
  }
  return 3;
}

// one crossing on axis c (split-vec.h cases 1, 2, 4)
template <typename T> PIC_HD void split3d_one(Seg<3, T> &a, Seg<3, T> &n, int c, int di) {
  const int o1 = (c + 1) % 3, o2 = (c + 2) % 3;
  T cint = T(0.5) * di;
  T delta = (cint - a.p0[c]) / (a.p1[c] - a.p0[c]);
  T i1 = lerp(a.p0[o1], a.p1[o1], delta);
  T i2 = lerp(a.p0[o2], a.p1[o2], delta);
  n.p0[c] = -cint; n.p1[c] = a.p1[c] - di;
  n.p0[o1] = i1;   n.p1[o1] = a.p1[o1];
  n.p0[o2] = i2;   n.p1[o2] = a.p1[o2];
  n.c[c] = a.c[c] + di; n.c[o1] = a.c[o1]; n.c[o2] = a.c[o2];
  a.p1[c] = cint; a.p1[o1] = i1; a.p1[o2] = i2;
}

// two crossings: cut on c1 first, then c2 in whichever piece crosses it (cases 3, 5, 6)
template <typename T>
PIC_HD void split3d_two(Seg<3, T> &a, Seg<3, T> &n1, Seg<3, T> &n2, int c1, int c2, int d1, int d2) {
  const int o = 3 - c1 - c2;
  T aint = T(0.5) * d1;
  T delta = (aint - a.p0[c1]) / (a.p1[c1] - a.p0[c1]);
  T bint = lerp(a.p0[c2], a.p1[c2], delta);
  T oint = lerp(a.p0[o], a.p1[o], delta);
  if ((bint >= T(-0.5)) && (bint < T(0.5))) {
    T bint2 = T(0.5) * d2;
    delta = (bint2 - bint) / (a.p1[c2] - bint);
    T aint2 = -aint + (a.p1[c1] - aint) * delta;
    T oint2 = lerp(oint, a.p1[o], delta);
    n1.p0[c1] = -aint; n1.p1[c1] = aint2;
    n1.p0[c2] = bint;  n1.p1[c2] = bint2;
    n1.p0[o] = oint;   n1.p1[o] = oint2;
    n1.c[c1] = a.c[c1] + d1; n1.c[c2] = a.c[c2]; n1.c[o] = a.c[o];
    n2.p0[c1] = aint2;  n2.p1[c1] = a.p1[c1] - d1;
    n2.p0[c2] = -bint2; n2.p1[c2] = a.p1[c2] - d2;
    n2.p0[o] = oint2;   n2.p1[o] = a.p1[o];
    n2.c[c1] = a.c[c1] + d1; n2.c[c2] = a.c[c2] + d2; n2.c[o] = a.c[o];
    a.p1[c1] = aint; a.p1[c2] = bint; a.p1[o] = oint;
  } else {
    T bint2 = T(0.5) * d2;
    delta = (bint2 - a.p0[c2]) / (bint - a.p0[c2]);
    T aint2 = lerp(a.p0[c1], aint, delta);
    T oint2 = lerp(a.p0[o], oint, delta);
    n1.p0[c1] = aint2;  n1.p1[c1] = aint;
    n1.p0[c2] = -bint2; n1.p1[c2] = bint - d2;
    n1.p0[o] = oint2;   n1.p1[o] = oint;
    n1.c[c1] = a.c[c1]; n1.c[c2] = a.c[c2] + d2; n1.c[o] = a.c[o];
    n2.p0[c1] = -aint;      n2.p1[c1] = a.p1[c1] - d1;
    n2.p0[c2] = bint - d2;  n2.p1[c2] = a.p1[c2] - d2;
    n2.p0[o] = oint;        n2.p1[o] = a.p1[o];
    n2.c[c1] = a.c[c1] + d1; n2.c[c2] = a.c[c2] + d2; n2.c[o] = a.c[o];
    a.p1[c1] = aint2; a.p1[c2] = bint2; a.p1[o] = oint2;
  }
}

// 3D splitter (vsplit3D, simd/split-vec.h:266-716); returns the number of segments (1..4)
template <typename T> PIC_HD int split3d(Seg<3, T> *s, const int *di) {
  const int cross = (di[0] != 0) + 2 * (di[1] != 0) + 4 * (di[2] != 0);
  switch (cross) {
  case 0: return 1;
  case 1: split3d_one(s[0], s[1], 0, di[0]); return 2;
  case 2: split3d_one(s[0], s[1], 1, di[1]); return 2;
  case 4: split3d_one(s[0], s[1], 2, di[2]); return 2;
  case 3: split3d_two(s[0], s[1], s[2], 0, 1, di[0], di[1]); return 3;
  case 5: split3d_two(s[0], s[1], s[2], 0, 2, di[0], di[2]); return 3;
  case 6: split3d_two(s[0], s[1], s[2], 1, 2, di[1], di[2]); return 3;
  default: break;
  }
  // x-y-z: x-y logic first, then cut the piece whose z1 left the cell (:591-708)
  split3d_two(s[0], s[1], s[2], 0, 1, di[0], di[1]);
  const T zint = T(0.5) * di[2];
  bool cut = false;  // unrolled with a flag instead of a break: s[] is only indexed statically
#pragma unroll
  for (int i = 0; i < 3; i++) {
    Seg<3, T> &a = s[i];
    if (!cut && ((a.p1[2] < T(-0.5)) || (a.p1[2] >= T(0.5)))) {
      cut = true;
      T delta = (zint - a.p0[2]) / (a.p1[2] - a.p0[2]);
      T xint = lerp(a.p0[0], a.p1[0], delta);
      T yint = lerp(a.p0[1], a.p1[1], delta);
      Seg<3, T> &n = s[3];
      n.p0[0] = xint;  n.p1[0] = a.p1[0];
      n.p0[1] = yint;  n.p1[1] = a.p1[1];
      n.p0[2] = -zint; n.p1[2] = a.p1[2] - di[2];
      n.c[0] = a.c[0]; n.c[1] = a.c[1]; n.c[2] = a.c[2] + di[2];
      a.p1[0] = xint; a.p1[1] = yint; a.p1[2] = zint;
#pragma unroll
      for (int j = i + 1; j < 3; j++) {
        s[j].p0[2] -= di[2];
        s[j].p1[2] -= di[2];
        s[j].c[2] += di[2];
      }
    }
  }
  return 4;
}

}  // namespace pic
