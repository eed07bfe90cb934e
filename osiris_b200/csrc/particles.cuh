// particles.cuh -- launch interface of the particle bookkeeping kernels (particles.cu)
#pragma once
#include <limits.h>

#include "common.cuh"

namespace osgpu {

// host-layout (Fortran) particle arrays staged on the device: x(D,np), ix(D,np), p(3,np), q(np)
template <typename T> struct AosParticles {
  const int *ix;
  const T *x;
  const T *p;
  const T *q;
  const long long *tag;  // may be null
};

// cells subtracted from ix before binning (moving window)
struct Shift3 {
  int n[3];
};
// what a mover meets on each side [lower/upper][axis] of this rank (route_kernel)
enum : int { PBC_NONE = 0, PBC_WRAP = 1, PBC_ABSORB = 2, PBC_SPECULAR = 3 };
struct PartBc {
  int t[2][3];
  // axes whose lower absorbing wall does not remove a particle now but hands it to the species' guest list:
  // the cells a window move pushed out of the box are pushed once more before the wall takes them
  // (move_window_species only lowers ix, spec/os-species.f03:171-229; update_boundary comes after the push)
  unsigned defer_lower;
};

// upload = counting sort by (tile, cell).  cell_cnt: [ntile][TILE_NCELL] zeroed scratch (counts, then
// scatter cursors); cell_start: [ntile][TILE_NCELL + 1] offsets; tile_cnt / tile_nsorted: per tile
template <typename T>
int bin_count(const AosParticles<T> &in, long long np, const Grid &g, const Tiling &tl, const Shift3 &shift,
              int to_movers, int *cell_cnt, int *cell_start, int *tile_cnt, int *tile_nsorted, int *errflag,
              cudaStream_t st, long long *launches);
template <typename T>
int bin_scatter(const AosParticles<T> &in, long long np, long long tag_base, const Grid &g, const Tiling &tl,
                const Shift3 &shift, int to_movers, const MoverView<T> &mv, int *errflag, const int *tile_off,
                const int *cell_start, int *cursor, const SpeciesView<T> &sp, cudaStream_t st, long long *launches);
// np host-layout particles appended to the mover buffer at `base` (count becomes base + np)
template <typename T>
int movers_inject(const AosParticles<T> &in, int np, long long tag_base, const Grid &g, const MoverView<T> &mv,
                  int base, cudaStream_t st, long long *launches);
template <typename T>
int gather_download(const SpeciesView<T> &sp, const int *tile_off, const int *tile_cnt, const long long *out_off,
                    const Grid &g, const Tiling &tl, int *oix, T *ox, T *op, T *oq, long long *otag, int by_tag,
                    cudaStream_t st, long long *launches);
template <typename T>
int movers_download(const MoverView<T> &mv, long long out_base, const Grid &g, int *oix, T *ox, T *op, T *oq,
                    long long *otag, int by_tag, cudaStream_t st, long long *launches);
template <typename T>
int movers_route(const MoverView<T> &mv, const Grid &g, const Tiling &tl, const PartBc &bc, const int *tile_off,
                 const int *tile_cnt, int *tile_arr, int *errflag, const MoverView<T> &guests, cudaStream_t st,
                 long long *launches);
template <typename T>
int movers_insert(const MoverView<T> &mv, const Grid &g, const Tiling &tl, const int *tile_off, int *tile_cnt,
                  const SpeciesView<T> &sp, cudaStream_t st, long long *launches);
template <typename T>
int relayout(const SpeciesView<T> &src, const int *old_off, const int *tile_cnt, const SpeciesView<T> &dst,
             const int *new_off, const Grid &g, const Tiling &tl, cudaStream_t st, long long *launches);
// charge density of a species on the grid nodes (deposit_rho, spec/os-spec-charge.f90:139-197, 463-589):
// rho(i + k) += q prod_d S_d(k_d) with the species' interpolation order, into one device plane (ADDED)
template <typename T>
int deposit_charge(const SpeciesView<T> &sp, const int *tile_off, const int *tile_cnt, const Grid &g, const Tiling &tl,
                   int order, T *rho, cudaStream_t st, long long *launches);
// phasespace deposit of a species (deposit_1D / deposit_2D of spec/diagnostics/os-spec-phasespace.f03) into a
// device buffer of n[0] * n[1] floats (ADDED); rqm and dx[] of the species' grid
template <typename T>
int deposit_phasespace(const SpeciesView<T> &sp, const int *tile_off, const int *tile_cnt, const Grid &g, const Tiling &tl,
                       const osgpu_phasespace &ps, double rqm, const double *dx, float *buf, cudaStream_t st,
                       long long *launches);
// one padded device plane -> host t_vdf layout f(1, S1h, S2h, S3h) staged on the device
template <typename T>
int plane_to_host_layout(const Grid &g, const T *plane, T *out, const int *hsize, const int *hgc, cudaStream_t st,
                         long long *launches);
template <typename T> int record_words(int dim);
template <typename T>
int movers_classify(const MoverView<T> &mv, const Grid &g, int axis, int send_lo, int send_up, uint32_t *buf_lo,
                    uint32_t *buf_up, int *cnt2, int cap_rec, int *errflag, cudaStream_t st, long long *launches);
template <typename T>
int movers_unpack(const uint32_t *buf, int nrec, const Grid &g, int axis, int add_axis, const MoverView<T> &mv,
                  int *errflag, cudaStream_t st, long long *launches);

}  // namespace osgpu
