// host_emul.cpp -- TEST ONLY.  Runs the product's per-particle statements (csrc/pic_math.cuh,
// csrc/tile.cuh: the exact functions the CUDA kernel calls) on the host, tile by tile, so that the
// arithmetic and the tile indexing can be checked against the oracle where there is no GPU.
// It is not a fallback: nothing in osiris_b200/ links or loads it.
#include <cstdint>
#include <cstring>
#include <vector>
#include "../../osiris_b200/csrc/tile.cuh"

using namespace pic;

template <int DIM, int ORDER, typename T>
static double run(int *ix, T *x, T *u, const T *q, long np, T rqm, const T *e, const T *b,
                  const int *esz, const int *eoff, T *j, const int *jsz, const int *joff,
                  const int *nx, const double *dx, double dt) {
  using G = Tile<DIM, ORDER, T>;
  PushConst<T> k;
  k.tem = (T)(0.5 * dt / rqm);
  for (int d = 0; d < DIM; d++) {
    k.dt_dx[d] = (T)(dt / dx[d]);
    k.jnorm[d] = (T)(dx[d] / dt / (DIM == 2 ? 2 : 3));
  }
  const int TT[3] = {G::TX, G::TY, G::TZ};
  int nt[3] = {1, 1, 1};
  for (int d = 0; d < DIM; d++) nt[d] = (nx[d] + TT[d] - 1) / TT[d];
  const long ntile = (long)nt[0] * nt[1] * nt[2];
  std::vector<std::vector<long>> bins(ntile);
  for (long p = 0; p < np; p++) {
    long t = 0, m = 1;
    for (int d = 0; d < DIM; d++) { t += m * ((ix[DIM * p + d] - 1) / TT[d]); m *= nt[d]; }
    bins[t].push_back(p);
  }
  std::vector<T> eb(6 * G::ESTRIDE), jt(3 * G::JVOL);
  double ene = 0;
  const long es[3] = {1, esz[0], (long)esz[0] * esz[1]};
  const long js[3] = {1, jsz[0], (long)jsz[0] * jsz[1]};
  for (long t = 0; t < ntile; t++) {
    int tc[3] = {(int)(t % nt[0]), (int)((t / nt[0]) % nt[1]), (int)(t / ((long)nt[0] * nt[1]))};
    int o[3] = {1, 1, 1};
    for (int d = 0; d < DIM; d++) o[d] = 1 + tc[d] * TT[d];
    // stage E/B (what the TMA box load does; out-of-range points read as zero)
    std::fill(eb.begin(), eb.end(), T(0));
    for (int c = 0; c < 6; c++)
      for (int kz = 0; kz < G::EZ; kz++)
        for (int ky = 0; ky < G::EY; ky++)
          for (int kx = 0; kx < G::EX; kx++) {
            int kk[3] = {kx, ky, kz};
            long pos = 0; bool ok = true;
            for (int d = 0; d < DIM; d++) {
              long pc = o[d] - 1 + eoff[d] + G::HALO_LO + kk[d];
              if (pc < 0 || pc >= esz[d]) ok = false;
              pos += pc * es[d];
            }
            const T *f = (c < 3) ? e : b;
            eb[c * G::ESTRIDE + kx + G::EXP * (ky + G::EY * kz)] = ok ? f[3 * pos + (c % 3)] : T(0);
          }
    std::fill(jt.begin(), jt.end(), T(0));
    for (long p : bins[t]) {
      int l[3] = {0, 0, 0};
      for (int d = 0; d < DIM; d++) l[d] = ix[DIM * p + d] - o[d];
      T en = push_deposit_one<DIM, ORDER, T>(x + DIM * p, u + 3 * p, q[p], l, eb.data(), jt.data(), k,
                                             [](T *a, T v) { *a += v; });
      ene += (double)q[p] * (double)en;
      for (int d = 0; d < DIM; d++) ix[DIM * p + d] = o[d] + l[d];
    }
    for (int c = 0; c < 3; c++)
      for (int kz = 0; kz < G::JZ; kz++)
        for (int ky = 0; ky < G::JY; ky++)
          for (int kx = 0; kx < G::JX; kx++) {
            T v = jt[c * G::JVOL + kx + G::JX * (ky + G::JY * kz)];
            if (v == T(0)) continue;
            int kk[3] = {kx, ky, kz};
            long pos = 0;
            for (int d = 0; d < DIM; d++) pos += (long)(o[d] - 1 + joff[d] + G::HALO_LO + kk[d]) * js[d];
            j[3 * pos + c] += v;
          }
  }
  return ene;
}

template <typename T>
static double dispatch(int dim, int order, int *ix, void *x, void *u, const void *q, long np, double rqm,
                       const void *e, const void *b, const int *esz, const int *eoff, void *j,
                       const int *jsz, const int *joff, const int *nx, const double *dx, double dt) {
#define CASE(D, O)                                                                              \
  if (dim == D && order == O)                                                                   \
    return run<D, O, T>(ix, (T *)x, (T *)u, (const T *)q, np, (T)rqm, (const T *)e, (const T *)b, \
                        esz, eoff, (T *)j, jsz, joff, nx, dx, dt);
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
  return -1;
}

extern "C" double emul_advance_deposit(int dim, int order, int real_bytes, int *ix, void *x, void *u,
                                       const void *q, long np, double rqm, const void *e, const void *b,
                                       const int *esz, const int *eoff, void *j, const int *jsz,
                                       const int *joff, const int *nx, const double *dx, double dt) {
  if (real_bytes == 4)
    return dispatch<float>(dim, order, ix, x, u, q, np, rqm, e, b, esz, eoff, j, jsz, joff, nx, dx, dt);
  return dispatch<double>(dim, order, ix, x, u, q, np, rqm, e, b, esz, eoff, j, jsz, joff, nx, dx, dt);
}
