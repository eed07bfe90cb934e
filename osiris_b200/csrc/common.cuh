// common.cuh -- device views, grid geometry and error plumbing shared by all translation units.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>

#include "../../include/osiris_b200.h"
#include "tile.cuh"

namespace osgpu {

void set_error(const char *fmt, ...);
extern thread_local int g_last_status;

#define OSGPU_CUDA(call)                                                                     \
  do {                                                                                       \
    cudaError_t e__ = (call);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      osgpu::set_error("CUDA error %s at %s:%d: %s", cudaGetErrorName(e__), __FILE__,        \
                       __LINE__, cudaGetErrorString(e__));                                   \
      return OSGPU_ERR_CUDA;                                                                 \
    }                                                                                        \
  } while (0)

#define OSGPU_TRY(call)          \
  do {                           \
    int s__ = (call);            \
    if (s__ != OSGPU_OK) return s__; \
  } while (0)

// Planar, padded device grid.  A t_vdf f(3, S1, S2, S3) becomes three component planes
// c[0..2], each [S3][S2][P0] with P0 = S1 rounded up to 4 elements (16-byte rows for TMA and
// vector loads).  Cell i (1-based, as in the reference) sits at padded position i - 1 + gl.
struct Grid {
  int dim;
  int nx[3];
  int gl[3], gu[3];  // storage guard cells (gl[0] is padded so tile rows start 16-byte aligned)
  int ul[3], uu[3];  // guard cells that take part in boundary updates (the host arrays' widths)
  int S[3];       // padded extents (S[2] = 1 in 2D)
  int P0;         // row pitch in elements
  long long vol;  // elements per component = P0 * S[1] * S[2]
  __host__ __device__ long long at(int i0, int i1, int i2) const {  // padded coordinates
    return i0 + (long long)P0 * (i1 + (long long)S[1] * i2);
  }
};

// Tile decomposition of the local domain
struct Tiling {
  int T[3];   // tile edge in cells
  int nt[3];  // tiles per axis
  int ntile;
};

template <typename T> struct Field3 {
  T *c[3];
};

// Particle buffer of one species, tile-binned: tile t owns slots [tile_off[t], tile_off[t+1]) of
// which the first tile_cnt[t] are live.  The real columns x(D) p(3) q are stored struct-of-arrays
// inside blocks of 32 slots: column c of slot s sits at base[((s / 32) * nc + c) * 32 + s % 32], so
// the 32 particles a warp handles in one round are nc consecutive 128-byte (fp32) rows -- one
// contiguous range for a TMA bulk copy, and constant row offsets from one per-lane address.
template <typename T> struct SpeciesView {
  T *base;
  int nc;           // real columns per slot: dim + 4
  int *cell;        // packed tile-local cell (tile.cuh), one int per slot
  long long *tag;   // null when the species carries no tags
  __host__ __device__ __forceinline__ long long blk(long long s) const { return (((s >> 5) * nc) << 5) + (s & 31); }
  __host__ __device__ __forceinline__ T &r(int c, long long s) const { return base[blk(s) + 32 * c]; }
};

// Particles that left their tile (or the rank) during a push: global ix, SoA.
template <typename T> struct MoverView {
  T *x[3];
  T *p[3];
  T *q;
  int *ix[3];       // 1-based local cell, may be 0 or nx+1 (guard cell) until the boundary step
  long long *tag;
  int *a, *b;       // scratch lists used by the in-tile compaction
  int *count;       // device counter
  int cap;
};

enum : int { ERRF_MOVER_OVERFLOW = 1, ERRF_TILE_OVERFLOW = 2, ERRF_BAD_CELL = 4, ERRF_SEND_OVERFLOW = 8 };

}  // namespace osgpu
