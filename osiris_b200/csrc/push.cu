// push.cu -- fused field interpolation + Boris push + trajectory split + charge-conserving current
// deposit + per-tile cell sort.  Replaces ADVANCE_DEPOSIT_2D/3D + vsplit?D + DEP_CURRENT_?D
// (reference: source/simd/os-spec-push-avx2.c:471-760, 1141-1472; split-vec.h;
// os-spec-current-avx2.c), the per-thread J grids + reduce of advance_deposit_particles
// (source/particles/os-particles.f03:1362-1483) and the cell sort of sort_species
// (source/spec/os-species.f03:269-702), which here happens every step inside the tile.
//
// Persistent CTAs take cell tiles (256 cells) from a device counter.  Per tile:
//   0. one thread issues six TMA box loads (E1..3, B1..3 tile + stencil halo) into shared memory;
//      everybody zeroes the shared J tile and the new-cell histogram meanwhile;
//   1. "push" phase, one thread per particle, coalesced SoA loads: interpolate from shared memory,
//      Boris, move, split the trajectory at cell faces.  The new state goes to an L2-resident
//      scratch range in the old order together with the particle's rank inside its new cell
//      (one native shared-memory integer atomic).  The part of the trajectory inside the old cell
//      (x0, x1, q) is parked in shared memory; the pieces in neighbouring cells (~15 % of the
//      particles) are appended to a small shared list;
//   2. "deposit" phase, one thread per cell: the tile's particles arrive sorted by cell (step 5 of
//      the previous launch), so a thread walks its cell's parked trajectories, accumulates the
//      whole stencil in registers and touches the shared J tile once per cell instead of once per
//      particle (fp32 shared atomics are compare-and-swap loops on sm_100a: ATOMS.CAST.SPIN);
//   3. the listed neighbour-cell pieces and the tile's unsorted tail (particles that arrived from
//      other tiles last step) are deposited one thread per piece with shared atomics;
//   4. the J tile goes to HBM once (red.global.add, tiles overlap in their halos);
//   5. exclusive scan of the histogram -> cell offsets; scratch -> particle buffer at
//      offset[new cell] + rank (the tile is sorted by cell again); particles that left the tile go
//      to the species' mover buffer.
#include <cuda.h>

#include "common.cuh"
#include "push.cuh"

namespace osgpu {

using namespace pic;

constexpr int NTHREADS = 256;
static_assert(NTHREADS == TILE_NCELL, "one thread per tile cell");

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

// exclusive block scan of one non-negative value per thread; *total = sum
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

// shared-memory plan of one instantiation
template <int DIM, int ORDER, typename T> struct Plan {
  using G = Tile<DIM, ORDER, T>;
  // orders 1-2: the in-cell piece of every sorted particle is accumulated in registers per cell
  // (12 / 54 values in 3D); orders 3-4 would need 144 / 300 registers and deposit directly
  static constexpr bool FAST = ORDER <= 2;
  static constexpr int DW = (DIM == 2) ? 6 : 7;  // parked trajectory: x0[D], x1[D], q, (vz)
  static constexpr int XW = DW + 1;              // listed piece: + packed tile-local cell
  static constexpr int BUF_BYTES = FAST ? 67584 : 0;
  static constexpr int XMIN_BYTES = 10240;  // room always left for the list
  static constexpr int BUF_WORDS = BUF_BYTES / (int)sizeof(T);
  static constexpr int DCAP = FAST ? (((BUF_BYTES - XMIN_BYTES) / (DW * (int)sizeof(T))) & ~31) : 0;
  static constexpr int NACC = 3 * G::NP * G::NP * (DIM == 3 ? G::NP : 1);
  static constexpr int OFF_JT = 6 * G::ESTRIDE * (int)sizeof(T);
  static constexpr int OFF_BUF = OFF_JT + round_up(3 * G::JVOL * (int)sizeof(T), 16);
  static constexpr int OFF_CS = OFF_BUF + BUF_BYTES;
  static constexpr int OFF_HIST = OFF_CS + round_up((G::NCELL + 1) * 4, 16);
  static constexpr int SMEM_BYTES = OFF_HIST + G::NCELL * 4;
  PIC_HD static int acc_index(int comp, int k1, int k2, int k3) {
    return ((comp * (DIM == 3 ? G::NP : 1) + k3) * G::NP + k2) * G::NP + k1;
  }
};

template <int DIM> __device__ __forceinline__ int linear_cell(const int *l) {
  using G = Tile<DIM, 1, float>;
  if constexpr (DIM == 2) return l[0] + G::TX * l[1];
  else return l[0] + G::TX * (l[1] + G::TY * l[DIM - 1]);
}

template <int DIM, int ORDER, typename T>
__global__ void __launch_bounds__(NTHREADS, (sizeof(T) == 8 && DIM == 3) ? 1 : 2)
    push_deposit_kernel(const __grid_constant__ TmaMaps maps, const PushArgs<T> a) {
  using G = Tile<DIM, ORDER, T>;
  using P = Plan<DIM, ORDER, T>;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *eb = reinterpret_cast<T *>(smem_raw);
  T *jt = reinterpret_cast<T *>(smem_raw + P::OFF_JT);
  T *buf = reinterpret_cast<T *>(smem_raw + P::OFF_BUF);
  int *cs = reinterpret_cast<int *>(smem_raw + P::OFF_CS);
  int *hist = reinterpret_cast<int *>(smem_raw + P::OFF_HIST);
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int s_tile, s_nx[2], s_mcnt, s_base;
  __shared__ int s_warp[NTHREADS / 32];
  __shared__ double s_ene[NTHREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const auto atomic_add = [](T *p, T v) { atomicAdd(p, v); };

  // this CTA's scratch range (new particle states in the old order, until the sort)
  const long long so = (long long)blockIdx.x * a.scr_stride;

  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  uint32_t phase = 0;
  double ene = 0.0;

  for (;;) {
    __syncthreads();  // previous tile done with shared memory; mbarrier initialised
    if (tid == 0) s_tile = atomicAdd(a.tile_counter, 1);
    __syncthreads();
    const int tile = s_tile;
    if (tile >= a.tl.ntile) break;
    const int n = a.tile_cnt[tile];
    if (n == 0) {
      if (tid == 0) a.tile_nsorted[tile] = 0;
      continue;
    }
    const int ns = P::FAST ? min(a.tile_nsorted[tile], n) : 0;
    const long long off = a.tile_off[tile];

    int o0[3], tlim[3];
    {
      int t = tile;
      const int t0 = t % a.tl.nt[0]; t /= a.tl.nt[0];
      const int t1 = t % a.tl.nt[1]; t /= a.tl.nt[1];
      const int tc[3] = {t0, t1, t};
#pragma unroll
      for (int d = 0; d < 3; d++) {
        o0[d] = tc[d] * a.tl.T[d];
        tlim[d] = min(a.tl.T[d], a.g.nx[d] - o0[d]);
      }
    }

    // ---- 0. stage E/B, clear J tile and histogram ------------------------------------------
    if (tid == 0) {
      s_nx[0] = 0;
      s_nx[1] = 0;
      s_mcnt = 0;
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
    hist[tid] = 0;
    if (P::FAST && ns > 0) {
      const int *g_cs = a.cell_start + (size_t)tile * (G::NCELL + 1);
      cs[tid] = g_cs[tid];
      if (tid == 0) cs[G::NCELL] = g_cs[G::NCELL];
    }
    __syncthreads();
    mbar_wait(&mbar, phase);
    phase ^= 1;

    // ---- 1..3: pieces of the sorted range, then the unsorted tail --------------------------
    int p0 = 0, piece = 0;
    while (p0 < n) {
      const bool sorted = P::FAST && (p0 < ns);
      const int p1 = sorted ? min(ns, p0 + P::DCAP) : n;
      const int np = p1 - p0;
      // shared list of neighbour-cell pieces: whatever the parked trajectories leave free
      T *xl = buf + P::DW * np;
      const int xcap = sorted ? (P::BUF_WORDS - P::DW * np) / P::XW : 0;

      // -- 1. push ---------------------------------------------------------------------------
      for (int i = p0 + tid; i < p1; i += NTHREADS) {
        const long long s = off + i;
        T x[DIM], u[3];
        int l[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) x[d] = a.sp.x[d][s];
#pragma unroll
        for (int d = 0; d < 3; d++) u[d] = a.sp.p[d][s];
        const T q = a.sp.q[s];
        const int cell = a.sp.cell[s];
        const long long tag = a.sp.tag[s];
#pragma unroll
        for (int d = 0; d < DIM; d++) l[d] = cell_l(cell, d);
        Seg<DIM, T> seg[DIM + 1];
        int nseg;
        const T en = push_one<DIM, ORDER, T>(x, u, l, eb, a.k, seg, &nseg);
        ene += (double)q * (double)en;
        bool moved = false;
        int ncell = 0;
#pragma unroll
        for (int d = 0; d < DIM; d++) {
          moved |= (l[d] < 0) || (l[d] >= tlim[d]);
          ncell |= (l[d] + 1) << (d * CELL_BITS);
        }
        if (DIM == 2) ncell |= 1 << (2 * CELL_BITS);
        int rank = 0;
        if (moved) ncell |= CELL_DEAD;
        else rank = atomicAdd(&hist[linear_cell<DIM>(l)], 1);
        const long long r = so + i;
#pragma unroll
        for (int d = 0; d < DIM; d++) a.scr.x[d][r] = x[d];
#pragma unroll
        for (int d = 0; d < 3; d++) a.scr.p[d][r] = u[d];
        a.scr.q[r] = q;
        a.scr.cell[r] = ncell;
        a.scr.tag[r] = tag;
        a.scr_rank[r] = rank;

        if (sorted) {
          const int j = i - p0;
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            buf[d * np + j] = seg[0].p0[d];
            buf[(DIM + d) * np + j] = seg[0].p1[d];
          }
          buf[2 * DIM * np + j] = q;
          if (DIM == 2) buf[(2 * DIM + 1) * np + j] = seg[0].vz;
          unsigned direct = 0;  // pieces the list had no room for
#pragma unroll
          for (int k = 1; k <= DIM; k++) {
            if (k < nseg) {
              const int e = atomicAdd(&s_nx[piece & 1], 1);
              if (e < xcap) {
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                  xl[d * xcap + e] = seg[k].p0[d];
                  xl[(DIM + d) * xcap + e] = seg[k].p1[d];
                }
                xl[2 * DIM * xcap + e] = q;
                if (DIM == 2) xl[(2 * DIM + 1) * xcap + e] = seg[k].vz;
                int pc = 0;  // cell coordinates -1..T in 5-bit fields: exact as a float
#pragma unroll
                for (int d = 0; d < DIM; d++) pc |= (seg[k].c[d] + 1) << (5 * d);
                xl[P::DW * xcap + e] = (T)pc;
              } else {
                direct |= 1u << k;
              }
            }
          }
          if (direct) deposit_segments<DIM, ORDER, T>(seg, nseg, direct, q, jt, a.k, atomic_add);
        } else {
          deposit_segments<DIM, ORDER, T>(seg, nseg, 0xfu, q, jt, a.k, atomic_add);
        }
      }

      if constexpr (P::FAST) {
        if (sorted) {
          __syncthreads();
          // -- 2. deposit, one thread per (cell, part): stencil accumulated in registers ---------
          // cells that overlap [p0, p1): ca = cell of particle p0, cb = cell of particle p1 - 1
          int ca, cb;
          {
            int lo = 0, hi = G::NCELL - 1;  // first c with cs[c + 1] > p0
            while (lo < hi) { const int m = (lo + hi) >> 1; if (cs[m + 1] > p0) hi = m; else lo = m + 1; }
            ca = lo;
            lo = ca; hi = G::NCELL - 1;     // first c with cs[c + 1] >= p1
            while (lo < hi) { const int m = (lo + hi) >> 1; if (cs[m + 1] >= p1) hi = m; else lo = m + 1; }
            cb = lo;
          }
          const int ncp = cb - ca + 1;
          int parts = 1;
          while (parts * 2 * ncp <= NTHREADS && parts < 32) parts *= 2;
          if (tid < ncp * parts) {
            const int c = ca + tid % ncp, part = tid / ncp;
            const int ia = max(cs[c], p0) + part, ib = min(cs[c + 1], p1);
            if (ia < ib) {
              T acc[P::NACC];
#pragma unroll
              for (int k = 0; k < P::NACC; k++) acc[k] = T(0);
              for (int i = ia; i < ib; i += parts) {
                const int j = i - p0;
                Seg<DIM, T> sg;
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                  sg.p0[d] = buf[d * np + j];
                  sg.p1[d] = buf[(DIM + d) * np + j];
                }
                const T q = buf[2 * DIM * np + j];
                if (DIM == 2) sg.vz = buf[(2 * DIM + 1) * np + j];
                segment_current<DIM, ORDER, T>(sg, q, a.k, [&](int comp, int k1, int k2, int k3, T f1, T f2) {
                  T &r = acc[P::acc_index(comp, k1, k2, k3)];
                  r = fmadd(f1, f2, r);
                });
              }
              int lc[3];
              lc[0] = c % G::TX;
              lc[1] = (c / G::TX) % G::TY;
              lc[2] = c / (G::TX * G::TY);
              T *pj = jt + jtile_base<DIM, ORDER, T>(lc);
              for_each_stencil_entry<DIM, ORDER>([&](int comp, int k1, int k2, int k3) {
                atomicAdd(pj + comp * G::JVOL + k1 + G::JX * (k2 + G::JY * k3),
                          acc[P::acc_index(comp, k1, k2, k3)]);
              });
            }
          }
          // -- 3. listed neighbour-cell pieces, one thread per piece -----------------------------
          const int nx = min(s_nx[piece & 1], xcap);
          if (tid == 0) s_nx[(piece + 1) & 1] = 0;
          for (int e = tid; e < nx; e += NTHREADS) {
            Seg<DIM, T> sg;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
              sg.p0[d] = xl[d * xcap + e];
              sg.p1[d] = xl[(DIM + d) * xcap + e];
            }
            const T q = xl[2 * DIM * xcap + e];
            if (DIM == 2) sg.vz = xl[(2 * DIM + 1) * xcap + e];
            const int pc = (int)xl[P::DW * xcap + e];
#pragma unroll
            for (int d = 0; d < DIM; d++) sg.c[d] = ((pc >> (5 * d)) & 31) - 1;
            deposit_segment<DIM, ORDER, T>(sg, q, jt, a.k, atomic_add);
          }
          __syncthreads();  // the next piece reuses buf
        }
      }
      p0 = p1;
      piece++;
    }
    __syncthreads();  // J tile, histogram and scratch complete

    // ---- 4. J tile -> HBM -------------------------------------------------------------------
    for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) {
      const T v = jt[i];
      if (v != T(0)) {
        const int c = i / G::JVOL;
        int r = i - c * G::JVOL;
        const int kx = r % G::JX; r /= G::JX;
        const int ky = r % G::JY;
        const int kz = r / G::JY;
        const int q0 = o0[0] + a.g.gl[0] + G::HALO_LO + kx;
        const int q1 = o0[1] + a.g.gl[1] + G::HALO_LO + ky;
        const int q2 = (DIM == 3) ? o0[2] + a.g.gl[2] + G::HALO_LO + kz : 0;
        if (q0 >= 0 && q0 < a.g.S[0] && q1 >= 0 && q1 < a.g.S[1] && q2 >= 0 && q2 < a.g.S[2])
          atomicAdd(&a.J[c][a.g.at(q0, q1, q2)], v);
        else
          atomicOr(a.errflag, ERRF_BAD_CELL);
      }
    }

    // ---- 5. sort by new cell: scratch -> particle buffer, leavers -> movers -----------------
    int nlive;
    const int start = block_scan(hist[tid], &nlive, s_warp);
    cs[tid] = start;
    const int nmov = n - nlive;
    if (tid == 0) {
      int b = 0;
      if (nmov > 0) {
        b = atomicAdd(a.mv.count, nmov);
        if (b + nmov > a.mv.cap) {
          atomicOr(a.errflag, ERRF_MOVER_OVERFLOW);
          b = -1;
        }
      }
      s_base = b;
      a.tile_cnt[tile] = nlive;
      a.tile_nsorted[tile] = nlive;
    }
    {
      int *g_cs = a.cell_start + (size_t)tile * (G::NCELL + 1);
      g_cs[tid] = start;
      if (tid == 0) g_cs[G::NCELL] = nlive;
    }
    __syncthreads();
    const int base = s_base;
    for (int i = tid; i < n; i += NTHREADS) {
      const long long r = so + i;
      const int cell = a.scr.cell[r];
      if (cell < 0) {
        if (base >= 0) {
          const int m = base + atomicAdd(&s_mcnt, 1);
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            a.mv.x[d][m] = a.scr.x[d][r];
            a.mv.ix[d][m] = o0[d] + cell_l(cell & ~CELL_DEAD, d) + 1;
          }
#pragma unroll
          for (int d = 0; d < 3; d++) a.mv.p[d][m] = a.scr.p[d][r];
          a.mv.q[m] = a.scr.q[r];
          a.mv.tag[m] = a.scr.tag[r];
        }
      } else {
        int l[DIM];
#pragma unroll
        for (int d = 0; d < DIM; d++) l[d] = cell_l(cell, d);
        const long long s = off + cs[linear_cell<DIM>(l)] + a.scr_rank[r];
#pragma unroll
        for (int d = 0; d < DIM; d++) a.sp.x[d][s] = a.scr.x[d][r];
#pragma unroll
        for (int d = 0; d < 3; d++) a.sp.p[d][s] = a.scr.p[d][r];
        a.sp.q[s] = a.scr.q[r];
        a.sp.cell[s] = cell;
        a.sp.tag[s] = a.scr.tag[r];
      }
    }
  }

  // ---- energy: warp shuffle, then one atomic per CTA ----------------------------------------
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ene += __shfl_down_sync(0xffffffffu, ene, o);
  if (lane == 0) s_ene[wid] = ene;
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; w++) t += s_ene[w];
    if (t != 0.0) atomicAdd(a.ene, t);
  }
}

template <int DIM, int ORDER, typename T>
static int launch(const TmaMaps &maps, const PushArgs<T> &a, int nblocks, cudaStream_t st) {
  using P = Plan<DIM, ORDER, T>;
  auto kern = push_deposit_kernel<DIM, ORDER, T>;
  static bool configured = false;
  const int smem = P::SMEM_BYTES;
  if (!configured) {
    OSGPU_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    configured = true;
  }
  kern<<<nblocks, NTHREADS, smem, st>>>(maps, a);
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}

template <typename T>
int launch_push_deposit(int dim, int order, const TmaMaps &maps, const PushArgs<T> &a, int nblocks,
                        cudaStream_t st) {
#define CASE(D, O) \
  if (dim == D && order == O) return launch<D, O, T>(maps, a, nblocks, st);
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
    *smem_bytes = Plan<D, O, T>::SMEM_BYTES;               \
    return;                                                \
  }
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
}
template void push_tile_shape<float>(int, int, int *, int *);
template void push_tile_shape<double>(int, int, int *, int *);

// resident CTAs per SM of the instantiation (the persistent grid is nSM x this)
template <typename T> int push_blocks_per_sm(int dim, int order) {
  int nb = 1;
#define CASE(D, O)                                                                                   \
  if (dim == D && order == O) {                                                                      \
    auto kern = push_deposit_kernel<D, O, T>;                                                        \
    cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Plan<D, O, T>::SMEM_BYTES); \
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, NTHREADS, Plan<D, O, T>::SMEM_BYTES);    \
  }
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
  return nb < 1 ? 1 : nb;
}
template int push_blocks_per_sm<float>(int, int);
template int push_blocks_per_sm<double>(int, int);

}  // namespace osgpu
