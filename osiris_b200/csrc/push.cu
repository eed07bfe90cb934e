// push.cu -- fused field interpolation + Boris push + trajectory split + charge-conserving current
// deposit, one CTA per cell tile.  Replaces ADVANCE_DEPOSIT_2D/3D + vsplit?D + DEP_CURRENT_?D
// (reference: source/simd/os-spec-push-avx2.c:471-760, 1141-1472; split-vec.h;
// os-spec-current-avx2.c) and the per-thread J grids + reduce of advance_deposit_particles
// (source/particles/os-particles.f03:1362-1483).
//
// Per CTA:
//   1. one thread issues six TMA box loads (E1..3, B1..3 tile + stencil halo) into shared memory
//      and everybody zeroes the shared J tile meanwhile;
//   2. the tile's particles are streamed from the SoA buffers (coalesced), pushed, and their
//      current is accumulated into the shared J tile with shared-memory atomics;
//   3. the J tile goes to HBM once (red.global.add, tiles overlap in their halos);
//   4. particles that left the tile are exported to the species' mover buffer and the tile's
//      slot range is compacted in place (holes filled from the tail).
#include <cuda.h>

#include "common.cuh"
#include "push.cuh"

namespace osgpu {

using namespace pic;

constexpr int NTHREADS = 256;

__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int c0,
                                            int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, "
      "{%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}

// exclusive block scan of a small non-negative value; returns the exclusive prefix, *total = sum
__device__ __forceinline__ int block_scan(int v, int *total, int *s_warp) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  __syncthreads();  // protect s_warp from the previous call's readers
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
  constexpr int NW = NTHREADS / 32;
#pragma unroll
  for (int w = 0; w < NW; w++) {
    int t = s_warp[w];
    if (w < wid) base += t;
    tot += t;
  }
  *total = tot;
  return base + inc - v;
}

template <int DIM, int ORDER, typename T>
__global__ void __launch_bounds__(NTHREADS)
    push_deposit_kernel(const __grid_constant__ TmaMaps maps, const PushArgs<T> a) {
  using G = Tile<DIM, ORDER, T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *eb = reinterpret_cast<T *>(smem_raw);
  T *jt = eb + 6 * G::ESTRIDE;
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int s_nmov, s_base;
  __shared__ int s_warp[NTHREADS / 32];
  __shared__ double s_ene[NTHREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int tile = blockIdx.x;
  const int n = a.tile_cnt[tile];
  if (n == 0) return;  // uniform for the CTA
  const long long off = a.tile_off[tile];

  int tc[3], o0[3], tlim[3];
  {
    int t = tile;
    tc[0] = t % a.tl.nt[0]; t /= a.tl.nt[0];
    tc[1] = t % a.tl.nt[1]; t /= a.tl.nt[1];
    tc[2] = t;
#pragma unroll
    for (int d = 0; d < 3; d++) {
      o0[d] = tc[d] * a.tl.T[d];
      tlim[d] = min(a.tl.T[d], a.g.nx[d] - o0[d]);
    }
  }

  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_nmov = 0;
  }
  __syncthreads();
  if (tid == 0) {
    mbar_expect_tx(&mbar, 6u * (uint32_t)(G::EVOL * sizeof(T)));
    const int c0 = o0[0] + a.g.gl[0] + G::HALO_LO;
    const int c1 = o0[1] + a.g.gl[1] + G::HALO_LO;
    const int c2 = o0[2] + a.g.gl[2] + G::HALO_LO;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      if (DIM == 2) tma_load_2d(eb + c * G::ESTRIDE, &maps.m[c], &mbar, c0, c1);
      else tma_load_3d(eb + c * G::ESTRIDE, &maps.m[c], &mbar, c0, c1, c2);
    }
  }
  for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) jt[i] = T(0);
  __syncthreads();
  mbar_wait(&mbar, 0);

  // ---- push + deposit ---------------------------------------------------------------------
  double ene = 0.0;
  const int nround = (n + 31) & ~31;
  for (int i = tid; i < nround; i += NTHREADS) {
    bool moved = false;
    if (i < n) {
      const long long s = off + i;
      T x[DIM], u[3];
      int l[DIM];
#pragma unroll
      for (int d = 0; d < DIM; d++) x[d] = a.sp.x[d][s];
#pragma unroll
      for (int d = 0; d < 3; d++) u[d] = a.sp.p[d][s];
      const T q = a.sp.q[s];
      const int cell = a.sp.cell[s];
#pragma unroll
      for (int d = 0; d < DIM; d++) l[d] = cell_l(cell, d);
      const T en = push_deposit_one<DIM, ORDER, T>(
          x, u, q, l, eb, jt, a.k, [](T *p, T v) { atomicAdd(p, v); });
      ene += (double)q * (double)en;
      int ncell = 0;
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        moved |= (l[d] < 0) || (l[d] >= tlim[d]);
        ncell |= (l[d] + 1) << (d * CELL_BITS);
      }
      if (DIM == 2) ncell |= 1 << (2 * CELL_BITS);
      if (moved) ncell |= CELL_DEAD;
#pragma unroll
      for (int d = 0; d < DIM; d++) a.sp.x[d][s] = x[d];
#pragma unroll
      for (int d = 0; d < 3; d++) a.sp.p[d][s] = u[d];
      a.sp.cell[s] = ncell;
    }
    const unsigned m = __ballot_sync(0xffffffffu, moved);
    if (lane == 0 && m) atomicAdd(&s_nmov, __popc(m));
  }

  // ---- energy: warp shuffle, then one atomic per CTA --------------------------------------
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ene += __shfl_down_sync(0xffffffffu, ene, o);
  if (lane == 0) s_ene[wid] = ene;
  __syncthreads();  // also orders the J tile and the particle stores
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; w++) t += s_ene[w];
    atomicAdd(a.ene, t);
  }

  // ---- J tile -> HBM ----------------------------------------------------------------------
  for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) {
    const T v = jt[i];
    if (v != T(0)) {
      const int c = i / G::JVOL;
      int r = i - c * G::JVOL;
      const int kx = r % G::JX; r /= G::JX;
      const int ky = r % G::JY;
      const int kz = r / G::JY;
      const int p0 = o0[0] + a.g.gl[0] + G::HALO_LO + kx;
      const int p1 = o0[1] + a.g.gl[1] + G::HALO_LO + ky;
      const int p2 = (DIM == 3) ? o0[2] + a.g.gl[2] + G::HALO_LO + kz : 0;
      if (p0 >= 0 && p0 < a.g.S[0] && p1 >= 0 && p1 < a.g.S[1] && p2 >= 0 && p2 < a.g.S[2])
        atomicAdd(&a.J[c][a.g.at(p0, p1, p2)], v);
      else
        atomicOr(a.errflag, ERRF_BAD_CELL);
    }
  }

  // ---- export leavers and compact the tile ------------------------------------------------
  const int nmov = s_nmov;
  if (nmov == 0) return;
  const int nlive = n - nmov;
  if (tid == 0) {
    int b = atomicAdd(a.mv.count, nmov);
    if (b + nmov > a.mv.cap) {
      atomicOr(a.errflag, ERRF_MOVER_OVERFLOW);
      b = -1;
    }
    s_base = b;
  }
  __syncthreads();
  const int base = s_base;
  if (base < 0) return;  // reported through errflag; the step is invalid

  int run_dead = 0, run_low = 0;
  for (int start = 0; start < n; start += NTHREADS) {
    const int i = start + tid;
    int cell = 0;
    bool dead = false;
    if (i < n) {
      cell = a.sp.cell[off + i];
      dead = cell < 0;
    }
    const bool low = dead && (i < nlive);
    int total;
    const int ex = block_scan((dead ? 1 : 0) | (low ? (1 << 16) : 0), &total, s_warp);
    if (dead) {
      const long long s = off + i;
      const int r = base + run_dead + (ex & 0xffff);
#pragma unroll
      for (int d = 0; d < DIM; d++) {
        a.mv.x[d][r] = a.sp.x[d][s];
        a.mv.ix[d][r] = o0[d] + cell_l(cell & ~CELL_DEAD, d) + 1;
      }
#pragma unroll
      for (int d = 0; d < 3; d++) a.mv.p[d][r] = a.sp.p[d][s];
      a.mv.q[r] = a.sp.q[s];
      a.mv.tag[r] = a.sp.tag[s];
      if (low) a.mv.a[base + run_low + (ex >> 16)] = i;
    }
    run_dead += total & 0xffff;
    run_low += total >> 16;
  }
  int run = 0;
  for (int start = nlive; start < n; start += NTHREADS) {
    const int i = start + tid;
    const bool live = (i < n) && (a.sp.cell[off + i] >= 0);
    int total;
    const int ex = block_scan(live ? 1 : 0, &total, s_warp);
    if (live) a.mv.b[base + run + ex] = i;
    run += total;
  }
  __syncthreads();
  for (int k = tid; k < run_low; k += NTHREADS) {
    const long long src = off + a.mv.b[base + k], dst = off + a.mv.a[base + k];
#pragma unroll
    for (int d = 0; d < DIM; d++) a.sp.x[d][dst] = a.sp.x[d][src];
#pragma unroll
    for (int d = 0; d < 3; d++) a.sp.p[d][dst] = a.sp.p[d][src];
    a.sp.q[dst] = a.sp.q[src];
    a.sp.cell[dst] = a.sp.cell[src];
    a.sp.tag[dst] = a.sp.tag[src];
  }
  if (tid == 0) a.tile_cnt[tile] = nlive;
}

template <int DIM, int ORDER, typename T>
static int launch(const TmaMaps &maps, const PushArgs<T> &a, int ntile, cudaStream_t st) {
  using G = Tile<DIM, ORDER, T>;
  auto kern = push_deposit_kernel<DIM, ORDER, T>;
  static bool configured = false;
  const int smem = G::SMEM_BYTES;
  if (!configured) {
    OSGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  kern<<<ntile, NTHREADS, smem, st>>>(maps, a);
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}

template <typename T>
int launch_push_deposit(int dim, int order, const TmaMaps &maps, const PushArgs<T> &a, int ntile,
                        cudaStream_t st) {
#define CASE(D, O) \
  if (dim == D && order == O) return launch<D, O, T>(maps, a, ntile, st);
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
  set_error("unsupported dim/order %d/%d", dim, order);
  return OSGPU_ERR_INVALID;
}
template int launch_push_deposit<float>(int, int, const TmaMaps &, const PushArgs<float> &, int,
                                        cudaStream_t);
template int launch_push_deposit<double>(int, int, const TmaMaps &, const PushArgs<double> &, int,
                                         cudaStream_t);

template <typename T> void push_tile_shape(int dim, int order, int *box, int *smem_bytes) {
#define CASE(D, O)                                         \
  if (dim == D && order == O) {                            \
    using G = Tile<D, O, T>;                               \
    box[0] = G::EXP; box[1] = G::EY; box[2] = G::EZ;       \
    *smem_bytes = G::SMEM_BYTES;                           \
    return;                                                \
  }
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
}
template void push_tile_shape<float>(int, int, int *, int *);
template void push_tile_shape<double>(int, int, int *, int *);

}  // namespace osgpu
