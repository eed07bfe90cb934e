// push.cuh -- launch interface of the fused push + deposit kernel (push.cu)
#pragma once
#include <cuda.h>

#include "common.cuh"

namespace osgpu {

// tensor maps of E1..3, B1..3 (box = one tile + interpolation halo, loaded) and J1..3 (box = one
// tile + deposit halo, reduce-added), passed as a __grid_constant__
struct alignas(64) TmaMaps {
  CUtensorMap m[9];
};

// SoA columns at a constant stride (the mover buffer as the kernel addresses it): real column c of
// slot s is base[c * stride + s] -- x(D), p(3), q in this order; `cell` is the first int column
// behind them (ix(D)); `tag` is null when the species carries no tags
template <typename T> struct Cols {
  T *base;
  long long stride;
  int *cell;
  long long *tag;
};

// accelerate_deposit (spec/os-spec-accel.f90:41-397), the pusher of the first n_accelerate / n_q_incr
// steps of a beam species: no fields, the particle free-streams along x1 with the velocity of
// u_accel = u_frac * ufl(1) (or u_frac * its own p1), momenta untouched, current deposited with q * q_frac
template <typename T> struct Accel {
  int on;             // 0: the standard pusher
  int per_particle;   // udist%use_particle_uacc
  T u_frac, q_frac;   // grad_linear / grad_regressive of the step (host side)
  T ufl1;             // udist%ufl(1)
};

template <typename T> struct PushArgs {
  SpeciesView<T> spc;   // species buffer
  SpeciesView<T> scc;   // per-CTA scratch ranges: new state in the old order (cell = new packed cell)
  int *scr_rank;        // scratch: rank of the particle inside its new cell
  Cols<T> mvc;          // mover buffer
  MoverView<T> mv;      // (count, cap)
  const int *tile_off;
  int *tile_cnt;
  int *tile_nsorted;    // the first tile_nsorted[t] particles of tile t are sorted by cell
  int *cell_start;      // [ntile][TILE_NCELL + 1] offsets of the cells inside the sorted range
  int scr_stride;       // slots per CTA scratch range (multiple of 32)
  int *tile_counter;    // persistent CTAs take tiles from here (zeroed before the launch)
  T *J[3];
  Grid g;
  Tiling tl;
  pic::PushConst<T> k;
  Accel<T> accel;
  double *ene;   // += sum(q * u^2/(gamma+1)) over the species
  int *errflag;
};

// push + deposit of the particles of a list (MoverView `in`, any cell the grids cover, guard cells included)
// straight from / into global memory; the pushed particles are appended to `out` with their new cell index.
// For the cells a window move left below the box: the reference pushes them once more before its
// update_boundary removes them (spec/os-species.f03:171-229).  E, B: the grids the particles interpolate.
template <typename T>
int launch_push_list(int dim, int order, const MoverView<T> &in, const MoverView<T> &out, const T *const *eb6, T *const *j3,
                     const Grid &g, const pic::PushConst<T> &k, double *ene, int *errflag, cudaStream_t st);

template <typename T>
int launch_push_deposit(int dim, int order, const TmaMaps &maps, const PushArgs<T> &a, int nblocks,
                        cudaStream_t st);
// resident CTAs per SM of the kernel for (dim, order)
template <typename T> int push_blocks_per_sm(int dim, int order);
// TMA box (elements, fastest first) and dynamic shared memory of the kernel for (dim, order)
template <typename T> void push_tile_shape(int dim, int order, int *box, int *jbox, int *smem_bytes);

}  // namespace osgpu
