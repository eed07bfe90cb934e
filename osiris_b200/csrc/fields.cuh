// fields.cuh -- launch interface of the grid kernels (fields.cu)
#pragma once
#include "common.cuh"

namespace osgpu {

template <typename T>
int yee_advance(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                cudaStream_t st, long long *launches);
// mode 0 = replace, 1 = add; comps = 3 (J) or 6 (E,B) component planes
template <typename T>
int periodic_update(const Grid &g, T *const *comps, int ncomp, int axis, int mode, cudaStream_t st,
                    long long *launches);
// dir 0 = pack slab -> buf, 1 = unpack (replace), 2 = unpack (add); a0..a1 padded coords on axis
template <typename T>
int slab_copy(const Grid &g, T *const *comps, int ncomp, int axis, int a0, int a1, T *buf, int dir,
              cudaStream_t st, long long *launches);
long long slab_volume(const Grid &g, int axis, int a0, int a1);
template <typename T>
int layout_convert(const Grid &g, Field3<T> f, T *aos, const int *hsize, const int *hgc, int to_device,
                   cudaStream_t st, long long *launches);
template <typename T>
int field_energy(const Grid &g, T *const *comps6, double *dev_out6, cudaStream_t st, long long *launches);

}  // namespace osgpu
