// fields.cuh -- launch interface of the grid kernels (fields.cu)
#pragma once
#include "common.cuh"

namespace osgpu {

template <typename T>
int yee_advance(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                cudaStream_t st, long long *launches);
// one substep of advance_yee: 0 = first half step of B, 1 = E, 2 = second half step of B (the callers
// that have wall hooks between the substeps use these; yee_advance == the three in a row)
template <typename T>
int yee_substep(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt, int sub,
                cudaStream_t st, long long *launches);
// customized-stencil ("Fei") solver, emf/solver/os-emf-solver-fei.f03: the x1 derivative of the Yee
// curl becomes an antisymmetric n_coef-point stencil with coef_e (in dB/dt) and coef_b (in dE/dt)
constexpr int SOLVER_MAX_NCOEF = 64;
template <typename T> struct SolverCoef {
  int n;
  T ce[SOLVER_MAX_NCOEF], cb[SOLVER_MAX_NCOEF];
};
template <typename T>
int fei_advance(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                const SolverCoef<T> &sc, cudaStream_t st, long long *launches);
// Mirror walls (perfect electric / magnetic conductor for E and B: emf/boundary/os-emf-pec.f90,
// os-emf-pmc.f90; specular fold of the current: current/os-current-boundary.f03).  Every statement of
// those files is "f(i) = s f(m - i)" over a run of guard cells (fields) or "f(m - i) += s f(i)" (current)
// plus one point on the mirror plane set to zero or doubled; one component's share of a wall:
struct WallComp {
  int m;          // mirror: cell i <-> cell m - i along the wall axis
  int s;          // +1 | -1
  int lo, hi;     // guard cells i = lo..hi (cell indices, inclusive; empty when lo > hi)
  int plane;      // cell on the mirror plane that gets plane_op, or INT_MIN
  int plane_op;   // 1: = 0, 2: *= 2
};
struct WallOp {
  int axis, fold;    // fold 0: assign guard cells from their images; 1: add guard cells onto their images
  WallComp c[3];
};
// the three components of one field (or J) at one wall
WallOp field_wall_op(const Grid &g, int axis, int side, int is_b, int pmc, int even, const int gc[2][3]);
WallOp current_wall_op(const Grid &g, int axis, int side, int order);
template <typename T> int wall_apply(const Grid &g, Field3<T> f, const WallOp &op, cudaStream_t st, long long *launches);
template <typename T>
int fei_substep(const Grid &g, Field3<T> e, Field3<T> b, Field3<T> jay, const double *dx, double dt,
                const SolverCoef<T> &sc, int sub, cudaStream_t st, long long *launches);
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
// current smoothing (vdf/os-vdf-smooth.f90): filter kernel of one axis, centre at SMOOTH_ORDER_MAX
// (p_order_max = 15, os-vdf-smooth.f90:34), and the box of the t_vdf's own extent in padded coordinates
constexpr int SMOOTH_ORDER_MAX = 15;
template <typename T> struct SmoothCoef {
  T s[2 * SMOOTH_ORDER_MAX + 1];
};
struct SmoothBox {
  int lo[3], hi[3];
};
// every pass of the filter, src -> dst (src == dst: in place) through the scratch planes tmp; each
// Field3 is three contiguous component planes (c[1] = c[0] + vol, c[2] = c[0] + 2 vol)
template <typename T>
int smooth_field(const Grid &g, Field3<T> src, Field3<T> dst, Field3<T> tmp, const SmoothBox &bx, const int *order,
                 const SmoothCoef<T> *coef, cudaStream_t st, long long *launches);
// moving window: f(i) = f(i + nmove) inside the box, zero where uncovered; tmp = three scratch planes
template <typename T>
int shift_field(const Grid &g, Field3<T> f, Field3<T> tmp, const SmoothBox &bx, const int *nmove, cudaStream_t st,
                long long *launches);
// dst += src over one padded plane
template <typename T> int plane_add(const Grid &g, T *dst, const T *src, cudaStream_t st, long long *launches);
template <typename T>
int field_energy(const Grid &g, T *const *comps6, double *dev_out6, cudaStream_t st, long long *launches);

}  // namespace osgpu
