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
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
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

__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap *map, const void *src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_3d(const CUtensorMap *map, const void *src, int c0, int c1,
                                                  int c2) {
  asm volatile("cp.reduce.async.bulk.tensor.3d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1), "r"(c2)
               : "memory");
}
__device__ __forceinline__ void l2_prefetch(const void *p, uint32_t bytes) {
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
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
  // one buffer per CTA: the parked trajectories of a piece from the bottom, above them one list
  // region per warp
  static constexpr int BUF_BYTES = FAST ? 74752 : 0;
  static constexpr int BUF_WORDS = BUF_BYTES / (int)sizeof(T);
  static constexpr int XMIN = 40 * (NTHREADS / 32);  // list entries always available
  static constexpr int TCAP = FAST ? (((BUF_WORDS - XMIN * XW) / DW) & ~255) : (1 << 30);
  static constexpr int NACC = 3 * G::NP * G::NP * (DIM == 3 ? G::NP : 1);
  static constexpr int OFF_JT = 6 * G::ESTRIDE * (int)sizeof(T);
  static constexpr int OFF_BUF = OFF_JT + round_up(3 * G::JVOL * (int)sizeof(T), 16);
  static constexpr int OFF_CS = OFF_BUF + BUF_BYTES;
  static constexpr int CS_STRIDE = round_up(G::NCELL + 1, 4);
  static constexpr int OFF_HIST = OFF_CS + 2 * CS_STRIDE * 4;
  // J tile -> HBM with one TMA tensor reduce-add per component (fp32)
  static constexpr bool TMA_J = sizeof(T) == 4;
  static constexpr int OFF_HEADS = OFF_HIST + G::NCELL * 4;
  static constexpr int SMEM_BYTES = OFF_HEADS + G::NCELL * 4;
  PIC_HD static int acc_index(int comp, int k1, int k2, int k3) {
    return ((comp * (DIM == 3 ? G::NP : 1) + k3) * G::NP + k2) * G::NP + k1;
  }
};

template <int DIM> __device__ __forceinline__ int linear_cell(const int *l) {
  using G = Tile<DIM, 1, float>;
  if constexpr (DIM == 2) return l[0] + G::TX * l[1];
  else return l[0] + G::TX * (l[1] + G::TY * l[DIM - 1]);
}

// N independent shared-memory adds with two round trips instead of N serialised compare-and-swap
// loops (fp32/fp64 atomicAdd on shared memory is an ATOMS.CAST.SPIN loop on sm_100a): read all the
// old values, then try every update with one CAS each; an update that lost a race falls back to
// the ordinary atomicAdd.  The addresses of one batch must not alias inside the warp.
__device__ __forceinline__ bool cas_add(float *p, float old, float v) {
  const int o = __float_as_int(old);
  return atomicCAS(reinterpret_cast<int *>(p), o, __float_as_int(old + v)) == o;
}
__device__ __forceinline__ bool cas_add(double *p, double old, double v) {
  const unsigned long long o = (unsigned long long)__double_as_longlong(old);
  return atomicCAS(reinterpret_cast<unsigned long long *>(p), o,
                   (unsigned long long)__double_as_longlong(old + v)) == o;
}
template <typename T, int N> __device__ __forceinline__ void add_batch(T *const (&p)[N], const T (&v)[N]) {
  T old[N];
#pragma unroll
  for (int k = 0; k < N; k++) old[k] = *reinterpret_cast<volatile T *>(p[k]);
#pragma unroll
  for (int k = 0; k < N; k++)
    if (!cas_add(p[k], old[k], v[k])) atomicAdd(p[k], v[k]);
}

// Flush of the per-cell stencil accumulators when the 32 lanes of a warp hold 32 consecutive cells
// (rows of TX cells along x).  The stencils of x-neighbours overlap: the lane that owns a grid
// column (the one whose entry k1 = KO lands there) collects its neighbours' entries for that
// column with shuffles and issues the only shared-memory update; entries whose owner would lie
// outside the row are added by the lane itself.  Updates are issued in batches of one y-row of
// the stencil (k2 fixed): inside a batch no two lanes of the warp touch the same word.
template <int DIM, int ORDER, typename T, class P>
__device__ __forceinline__ void flush_merged(const T *acc, T *pj, int x) {
  using G = Tile<DIM, ORDER, T>;
  constexpr int NP = G::NP;
  constexpr unsigned FULL = 0xffffffffu;
  constexpr int NZ = (DIM == 3) ? NP : 1;
  // entries along x and the owner entry, per component
  constexpr int NK1[3] = {NP - 1, NP, NP};
  constexpr int KO[3] = {(NP - 2 < 1) ? NP - 2 : 1, (NP - 1 < 1) ? NP - 1 : 1, (NP - 1 < 1) ? NP - 1 : 1};
#pragma unroll
  for (int k2 = 0; k2 < NP; k2++) {
    // -- columns owned by this lane --
    constexpr int NZ2 = (DIM == 3) ? NZ - 1 : NZ;  // component-3 entries per row
    constexpr int NB = 2 * NZ + NZ2;               // entries of a row (the last row has no component 2)
    T *bp[NB];
    T bv[NB];
    int nb = 0;
#pragma unroll
    for (int comp = 0; comp < 3; comp++) {
      if (comp == 1 && k2 >= NP - 1) continue;
#pragma unroll
      for (int k3 = 0; k3 < NZ; k3++) {
        if (DIM == 3 && comp == 2 && k3 >= NP - 1) continue;
        const int nk1 = NK1[comp], ko = KO[comp];
        T m = acc[P::acc_index(comp, ko, k2, k3)];
#pragma unroll
        for (int d = 1; d <= ko; d++) {  // entry ko - d of the lane d cells to the right
          const T t = __shfl_down_sync(FULL, acc[P::acc_index(comp, ko - d, k2, k3)], d);
          if (x + d < G::TX) m += t;
        }
#pragma unroll
        for (int d = 1; d <= nk1 - 1 - ko; d++) {  // entry ko + d of the lane d cells to the left
          const T t = __shfl_up_sync(FULL, acc[P::acc_index(comp, ko + d, k2, k3)], d);
          if (x - d >= 0) m += t;
        }
        bp[nb] = pj + comp * G::JVOL + G::JX * (k2 + G::JY * k3) + ko;
        bv[nb] = m;
        nb++;
      }
    }
    if (k2 < NP - 1) add_batch<T, NB>(bp, bv);
    else {  // the last row has no component-1 entries
      T *cp[NB - NZ];
      T cv[NB - NZ];
#pragma unroll
      for (int k = 0; k < NB - NZ; k++) { cp[k] = bp[k]; cv[k] = bv[k]; }
      add_batch<T, NB - NZ>(cp, cv);
      static_assert(NB - NZ == NZ + NZ2, "row without component-2 entries");
    }
    // -- columns whose owner cell lies outside the row: first and last lane of the row --
    if (x == 0 || x == G::TX - 1) {
#pragma unroll
      for (int comp = 0; comp < 3; comp++) {
        if (comp == 1 && k2 >= NP - 1) continue;
#pragma unroll
        for (int k3 = 0; k3 < NZ; k3++) {
          if (DIM == 3 && comp == 2 && k3 >= NP - 1) continue;
          const int nk1 = NK1[comp], ko = KO[comp];
#pragma unroll
          for (int k1 = 0; k1 < nk1; k1++) {
            if (k1 == ko) continue;
            const int xo = x + k1 - ko;  // owner cell of this column
            if (xo < 0 || xo >= G::TX)
              atomicAdd(pj + comp * G::JVOL + G::JX * (k2 + G::JY * k3) + k1, acc[P::acc_index(comp, k1, k2, k3)]);
          }
        }
      }
    }
  }
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
  int *cs2 = reinterpret_cast<int *>(smem_raw + P::OFF_CS);  // two arrays: current / next tile
  int *hist = reinterpret_cast<int *>(smem_raw + P::OFF_HIST);
  int *heads = reinterpret_cast<int *>(smem_raw + P::OFF_HEADS);  // per cell: list of foreign pieces
  __shared__ __align__(8) uint64_t mbar;
  __shared__ int s_xl[NTHREADS / 32], s_xcap[NTHREADS / 32];
  __shared__ int s_tile, s_n, s_ns, s_mcnt, s_base;
  __shared__ long long s_off, s_next_off;
  __shared__ int s_next[3];
  __shared__ int s_warp[NTHREADS / 32];
  __shared__ double s_ene[NTHREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const auto atomic_add = [](T *p, T v) { atomicAdd(p, v); };

  // this CTA's scratch range (new particle states in the old order, until the sort)
  const long long so = (long long)blockIdx.x * a.scr_stride;

  // Tile bookkeeping is pipelined: while a tile is pushed, one thread has already taken the CTA's
  // next tile from the device counter and asked L2 for its particle columns; as soon as the push
  // is over (E/B tile free) the same thread starts the next tile's TMA loads, which then overlap
  // the J write-back and the sort of the current tile.
  auto take_tile = [&](int *n_out, int *ns_out, long long *off_out) {  // skips empty tiles
    int t;
    for (;;) {
      t = atomicAdd(a.tile_counter, 1);
      if (t >= a.tl.ntile) { *n_out = 0; return t; }
      const int n = a.tile_cnt[t];
      if (n > 0) {
        *n_out = n;
        *ns_out = P::FAST ? min(a.tile_nsorted[t], n) : 0;
        *off_out = a.tile_off[t];
        return t;
      }
      a.tile_nsorted[t] = 0;
    }
  };
  auto tile_origin = [&](int tile, int *o0) {
    int t = tile;
    const int t0 = t % a.tl.nt[0]; t /= a.tl.nt[0];
    const int t1 = t % a.tl.nt[1]; t /= a.tl.nt[1];
    o0[0] = t0 * a.tl.T[0];
    o0[1] = t1 * a.tl.T[1];
    o0[2] = t * a.tl.T[2];
  };
  auto stage_fields = [&](int tile) {  // one thread: six TMA box loads of the tile's E/B + halo
    int o0[3];
    tile_origin(tile, o0);
    mbar_expect_tx(&mbar, 6u * (uint32_t)(G::EVOL * sizeof(T)));
    const int c0 = o0[0] + a.g.gl[0] + G::HALO_LO;
    const int c1 = o0[1] + a.g.gl[1] + G::HALO_LO;
    const int c2 = o0[2] + a.g.gl[2] + G::HALO_LO;
#pragma unroll
    for (int c = 0; c < 6; c++) {
      if (DIM == 2) tma_load_2d(eb + c * G::ESTRIDE, &maps.m[c], &mbar, c0, c1);
      else tma_load_3d(eb + c * G::ESTRIDE, &maps.m[c], &mbar, c0, c1, c2);
    }
  };
  auto prefetch_particles = [&](int n, long long off) {  // one thread: ask L2 for a tile's columns
    const uint32_t bytes = (uint32_t)(((size_t)n * sizeof(T) + 15) & ~(size_t)15);
#pragma unroll
    for (int c = 0; c < DIM + 4; c++) l2_prefetch(a.spc.base + c * a.spc.stride + off, bytes);
    l2_prefetch(a.spc.cell + off, (uint32_t)(((size_t)n * 4 + 15) & ~(size_t)15));
    if (a.spc.tag) l2_prefetch(a.spc.tag + off, (uint32_t)(((size_t)n * 8 + 15) & ~(size_t)15));
  };

  if (tid == 0) {
    mbar_init(&mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mcnt = 0;
  }
  for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) jt[i] = T(0);
  hist[tid] = 0;
  heads[tid] = 0;
  __syncthreads();
  if (tid == 32) {
    int n = 0, ns = 0;
    long long off = 0;
    const int t = take_tile(&n, &ns, &off);
    s_tile = t; s_n = n; s_ns = ns; s_off = off;
    if (t < a.tl.ntile) stage_fields(t);
  }
  uint32_t phase = 0;
  double ene = 0.0;
  T ene_t = T(0);  // per-tile partial in particle precision (the reference sums a chunk in REAL)
  int cur = 0;  // which of the two cell-offset arrays holds the current tile's
  __syncthreads();
  if (P::FAST && s_tile < a.tl.ntile && s_ns > 0) {
    const int *g_cs = a.cell_start + (size_t)s_tile * (G::NCELL + 1);
    cs2[tid] = g_cs[tid];
    if (tid == 0) cs2[G::NCELL] = g_cs[G::NCELL];
  }

  for (;;) {
    __syncthreads();  // s_tile .. s_off of this tile published
    const int tile = s_tile;
    if (tile >= a.tl.ntile) break;
    const int n = s_n, ns = s_ns;
    const long long off = s_off;
    int *cs = cs2 + cur * P::CS_STRIDE;

    int o0[3], tlim[3];
    tile_origin(tile, o0);
#pragma unroll
    for (int d = 0; d < 3; d++) tlim[d] = min(a.tl.T[d], a.g.nx[d] - o0[d]);

    // the CTA's next tile: taken now, its particle columns requested from L2, used at the end
    if (tid == 32) {
      int nx_n = 0, nx_ns = 0;
      long long nx_off = 0;
      const int nx_tile = take_tile(&nx_n, &nx_ns, &nx_off);
      if (nx_tile < a.tl.ntile) prefetch_particles(nx_n, nx_off);
      s_next[0] = nx_tile; s_next[1] = nx_n; s_next[2] = nx_ns; s_next_off = nx_off;
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;

    // ---- 1..3: push, then deposit; every trajectory piece inside the tile is routed to the lane
    // that owns its cell, so its current is accumulated in registers ----------------------------
    {
      constexpr unsigned FULL = 0xffffffffu;
      const unsigned lt_mask = (1u << lane) - 1u;
      const int c_lo = 32 * wid;
      // The tile's particles -- [0, ns) sorted by cell, [ns, n) the unsorted tail -- are pushed in
      // pieces of <= TCAP; inside a piece the rounds of 32 are dealt round-robin to the warps, so
      // every warp has the same number of rounds whatever its cells hold.
      T *wbuf = buf;  // parked trajectories of the piece, SoA with stride `tot`
      int qa = 0;
      bool more = false;
      do {
        const int qb = min(n, qa + P::TCAP);
        const int tot = qb - qa;                          // particles of this piece
        const int np = max(0, min(qb, ns) - qa);          // of which sorted: their trajectories are parked
        const int pa = qa, pb = qa + np;                  // sorted particles [pa, pb)
        more = qb < n;
        // list of pieces outside their particle's cell, above the parked trajectories: pieces in
        // tile cells grow from the bottom (linked to their cell), halo pieces from the top
        const int xcap = P::FAST ? (P::BUF_WORDS - P::DW * tot) / P::XW / (NTHREADS / 32) : 0;  // per warp
        T *xl = wbuf + P::DW * tot + wid * (P::XW * xcap);
        const int xcap_lo = xcap - xcap / 4;
        int nlo = 0, nhi = 0;                         // warp-uniform list lengths
        if (P::FAST && lane == 0) {
          s_xl[wid] = (int)(xl - buf);
          s_xcap[wid] = xcap;
        }

        // -- 1. push, 32 particles per round; the next round's loads are issued first ----------
        T nx_[DIM], nu_[3], nq_ = T(0);
        int ncell_ = 0;
        long long ntag_ = 0;
        auto fetch = [&](int m) {
          if (m < tot) {
            const long long s = off + qa + m;
#pragma unroll
            for (int d = 0; d < DIM; d++) nx_[d] = a.spc.base[d * a.spc.stride + s];
#pragma unroll
            for (int d = 0; d < 3; d++) nu_[d] = a.spc.base[(DIM + d) * a.spc.stride + s];
            nq_ = a.spc.base[(DIM + 3) * a.spc.stride + s];
            ncell_ = a.spc.cell[s];
            if (a.spc.tag) ntag_ = a.spc.tag[s];
          }
        };
        fetch(32 * wid + lane);
        for (int m0 = 32 * wid; m0 < tot; m0 += NTHREADS) {
          const int m = m0 + lane;
          const bool active = m < tot, parked = m < np;
          T x[DIM], u[3];
#pragma unroll
          for (int d = 0; d < DIM; d++) x[d] = nx_[d];
#pragma unroll
          for (int d = 0; d < 3; d++) u[d] = nu_[d];
          const T q = nq_;
          const int cell = ncell_;
          const long long tag = ntag_;
          fetch(m + NTHREADS);
          Seg<DIM, T> seg[DIM + 1];
          int nseg = 0;
          if (active) {
            int l[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) l[d] = cell_l(cell, d);
            const T en = push_one<DIM, ORDER, T>(x, u, l, eb, a.k, seg, &nseg);
            ene_t = fmadd(q, en, ene_t);
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
            const long long r = so + qa + m;
#pragma unroll
            for (int d = 0; d < DIM; d++) a.scc.base[d * a.scc.stride + r] = x[d];
#pragma unroll
            for (int d = 0; d < 3; d++) a.scc.base[(DIM + d) * a.scc.stride + r] = u[d];
            a.scc.base[(DIM + 3) * a.scc.stride + r] = q;
            a.scc.cell[r] = ncell;
            if (a.spc.tag) a.scc.tag[r] = tag;
            a.scc.cell[a.scc.stride + r] = rank;
            if (P::FAST && parked) {
#pragma unroll
              for (int d = 0; d < DIM; d++) {
                wbuf[d * tot + m] = seg[0].p0[d];
                wbuf[(DIM + d) * tot + m] = seg[0].p1[d];
              }
              wbuf[2 * DIM * tot + m] = q;
              if (DIM == 2) wbuf[(2 * DIM + 1) * tot + m] = seg[0].vz;
            }
          }
          // pieces outside the particle's own cell (and every piece of a tail particle)
          unsigned direct = 0;
#pragma unroll
          for (int k = 0; k <= DIM; k++) {
            const bool has = active && (k < nseg) && (k > 0 || !(P::FAST && parked));
            bool inside = true;
            int lcell = 0;
            if (has) {
#pragma unroll
              for (int d = 0; d < DIM; d++) inside &= (seg[k].c[d] >= 0) && (seg[k].c[d] < (d == 0 ? G::TX : (d == 1 ? G::TY : G::TZ)));
              lcell = linear_cell<DIM>(seg[k].c);
            }
            const unsigned hm = __ballot_sync(FULL, has);
            const unsigned im = __ballot_sync(FULL, has && inside);
            if (has) {
              const int e = inside ? nlo + __popc(im & lt_mask) : xcap - 1 - (nhi + __popc((hm & ~im) & lt_mask));
              const bool fits = P::FAST && (inside ? e < xcap_lo : e >= xcap_lo);
              if (fits) {
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                  xl[d * xcap + e] = seg[k].p0[d];
                  xl[(DIM + d) * xcap + e] = seg[k].p1[d];
                }
                xl[2 * DIM * xcap + e] = q;
                if (DIM == 2) xl[(2 * DIM + 1) * xcap + e] = seg[k].vz;
                int link;
                if (inside) {
                  link = atomicExch(&heads[lcell], ((wid << 12) | e) + 1);  // push on the cell's list
                } else {
                  link = 0;  // cell coordinates -1..T in 5-bit fields (exact as a float)
#pragma unroll
                  for (int d = 0; d < DIM; d++) link |= (seg[k].c[d] + 1) << (5 * d);
                }
                xl[P::DW * xcap + e] = (T)link;
              } else {
                direct |= 1u << k;
              }
            }
            nlo += __popc(im);
            nhi += __popc(hm & ~im);
          }
          if (direct) deposit_segments<DIM, ORDER, T>(seg, nseg, direct, q, jt, a.k, atomic_add);
        }

        __syncthreads();  // every warp's parked trajectories and lists are complete

        if constexpr (P::FAST) {
          // -- 2. deposit: one lane per (cell, part) walks the cell's parked trajectories and its
          // list of foreign pieces; the whole stencil is accumulated in registers -------------
          const int mya = max(cs[c_lo + lane], pa), myb = min(cs[c_lo + lane + 1], pb);
          const unsigned cm = __ballot_sync(FULL, (np > 0 && mya < myb) || heads[c_lo + lane] != 0);
          if (cm) {
            const int first = __ffs(cm) - 1;
            const int ncp = 32 - __clz(cm) - first;
            int parts = 1;
            while (parts * 2 * ncp <= 32) parts *= 2;
            const bool on = lane < ncp * parts;
            const int cl = first + lane % ncp, part = lane / ncp;
            const int c = c_lo + cl;
            const int ia = (on && np > 0) ? max(cs[c], pa) + part : 0, ib = (on && np > 0) ? min(cs[c + 1], pb) : 0;
            T acc[P::NACC];
#pragma unroll
            for (int k = 0; k < P::NACC; k++) acc[k] = T(0);
            const auto accumulate = [&](const Seg<DIM, T> &sg, T q) {
              segment_current<DIM, ORDER, T>(sg, q, a.k, [&](int comp, int k1, int k2, int k3, T f1, T f2) {
                T &r = acc[P::acc_index(comp, k1, k2, k3)];
                r = fmadd(f1, f2, r);
              });
            };
            for (int i = ia; i < ib; i += parts) {
              const int j = i - pa;
              Seg<DIM, T> sg;
#pragma unroll
              for (int d = 0; d < DIM; d++) {
                sg.p0[d] = wbuf[d * tot + j];
                sg.p1[d] = wbuf[(DIM + d) * tot + j];
              }
              const T q = wbuf[2 * DIM * tot + j];
              if (DIM == 2) sg.vz = wbuf[(2 * DIM + 1) * tot + j];
              accumulate(sg, q);
            }
            int id = (on && part == 0) ? heads[c] : 0;
            if (on && part == 0 && id) heads[c] = 0;
            bool any = ia < ib;
            while (id) {
              any = true;
              const int w = (id - 1) >> 12, e = (id - 1) & 4095;
              const T *xe = buf + s_xl[w] + e;
              const int st = s_xcap[w];
              Seg<DIM, T> sg;
#pragma unroll
              for (int d = 0; d < DIM; d++) {
                sg.p0[d] = xe[d * st];
                sg.p1[d] = xe[(DIM + d) * st];
              }
              const T q = xe[2 * DIM * st];
              if (DIM == 2) sg.vz = xe[(2 * DIM + 1) * st];
              id = (int)xe[P::DW * st];
              accumulate(sg, q);
            }
            int lc[3];
            lc[0] = c % G::TX;
            lc[1] = (c / G::TX) % G::TY;
            lc[2] = c / (G::TX * G::TY);
            T *pj = jt + jtile_base<DIM, ORDER, T>(lc);
            if (ncp == 32) {
              // lanes are consecutive cells along x: neighbours add up the stencil columns they
              // share with warp shuffles; one shared atomic per grid point of the row instead of
              // one per cell and point
              flush_merged<DIM, ORDER, T, P>(acc, pj, lc[0]);
            } else if (any) {
              for_each_stencil_entry<DIM, ORDER>([&](int comp, int k1, int k2, int k3) {
                atomicAdd(pj + comp * G::JVOL + k1 + G::JX * (k2 + G::JY * k3),
                          acc[P::acc_index(comp, k1, k2, k3)]);
              });
            }
          }
          // -- 3. pieces in halo cells (owned by a neighbouring tile): straight into the J tile --
          for (int e = xcap - 1 - lane; e >= max(xcap - nhi, xcap_lo); e -= 32) {
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
        }
        if (more) __syncthreads();  // the next piece reuses the buffers other warps may still read
        qa = qb;
      } while (more);
    }
    ene += (double)ene_t;
    ene_t = T(0);
    __syncthreads();  // J tile, histogram and scratch complete; E/B tile and cs free

    // ---- 4. J tile -> HBM; the next tile's E/B loads start ----------------------------------
    tile_origin(tile, o0);  // (recomputed: not kept in registers through the push)
    if constexpr (P::TMA_J) {
      if (tid == 0) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // atomics -> visible to TMA
        const int c0 = o0[0] + a.g.gl[0] + G::HALO_LO;
        const int c1 = o0[1] + a.g.gl[1] + G::HALO_LO;
        const int c2 = o0[2] + a.g.gl[2] + G::HALO_LO;
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (DIM == 2) tma_reduce_add_2d(&maps.m[6 + c], jt + c * G::JVOL, c0, c1);
          else tma_reduce_add_3d(&maps.m[6 + c], jt + c * G::JVOL, c0, c1, c2);
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {
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
        }
      }
    }
    if (tid == 32) {
      s_tile = s_next[0]; s_n = s_next[1]; s_ns = s_next[2]; s_off = s_next_off;
      if (s_next[0] < a.tl.ntile) stage_fields(s_next[0]);
    }

    // ---- 5. sort by new cell: scratch -> particle buffer, leavers -> movers -----------------
    int nlive;
    const int start = block_scan(hist[tid], &nlive, s_warp);
    hist[tid] = start;  // counts -> offsets
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
    // cell offsets of the CTA's next tile (published by thread 32 above): loaded now, stored below
    int pre = 0, pre_last = 0;
    const bool pre_on = P::FAST && s_tile < a.tl.ntile && s_ns > 0;
    if (pre_on) {
      const int *g_cs = a.cell_start + (size_t)s_tile * (G::NCELL + 1);
      pre = g_cs[tid];
      if (tid == 0) pre_last = g_cs[G::NCELL];
    }
    // PB particles per thread in flight: the scratch range sits in L2, latency is what costs
    tile_origin(tile, o0);
    constexpr int PB = 2;
#pragma unroll 1
    for (int i0 = 0; i0 < n; i0 += PB * NTHREADS) {
      int cellv[PB], dst[PB];  // dst: slot in the particle buffer, or ~slot in the mover buffer
      T vx[PB][DIM], vp[PB][3], vq[PB];
#pragma unroll
      for (int k = 0; k < PB; k++) {
        const int i = i0 + tid + k * NTHREADS;
        cellv[k] = 0;
        dst[k] = 0;
        if (i < n) {
          cellv[k] = a.scc.cell[so + i];
          dst[k] = a.scc.cell[a.scc.stride + so + i];
        }
      }
#pragma unroll
      for (int k = 0; k < PB; k++) {
        const int i = i0 + tid + k * NTHREADS;
        if (i < n) {
          const long long r = so + i;
#pragma unroll
          for (int d = 0; d < DIM; d++) vx[k][d] = a.scc.base[d * a.scc.stride + r];
#pragma unroll
          for (int d = 0; d < 3; d++) vp[k][d] = a.scc.base[(DIM + d) * a.scc.stride + r];
          vq[k] = a.scc.base[(DIM + 3) * a.scc.stride + r];
        }
      }
#pragma unroll
      for (int k = 0; k < PB; k++) {
        const int i = i0 + tid + k * NTHREADS;
        if (i >= n) continue;
        if (cellv[k] < 0) {
          if (base < 0) continue;  // mover buffer overflow: reported through errflag
          const int m = base + atomicAdd(&s_mcnt, 1);
          if (a.spc.tag) a.scc.cell[a.scc.stride + so + i] = m;  // the tag pass below follows
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            a.mvc.base[d * a.mvc.stride + m] = vx[k][d];
            a.mvc.cell[d * a.mvc.stride + m] = o0[d] + cell_l(cellv[k] & ~CELL_DEAD, d) + 1;
          }
#pragma unroll
          for (int d = 0; d < 3; d++) a.mvc.base[(DIM + d) * a.mvc.stride + m] = vp[k][d];
          a.mvc.base[(DIM + 3) * a.mvc.stride + m] = vq[k];
        } else {
          int l[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) l[d] = cell_l(cellv[k], d);
          const long long s = off + hist[linear_cell<DIM>(l)] + dst[k];
#pragma unroll
          for (int d = 0; d < DIM; d++) a.spc.base[d * a.spc.stride + s] = vx[k][d];
#pragma unroll
          for (int d = 0; d < 3; d++) a.spc.base[(DIM + d) * a.spc.stride + s] = vp[k][d];
          a.spc.base[(DIM + 3) * a.spc.stride + s] = vq[k];
          a.spc.cell[s] = cellv[k];
        }
      }
    }
    if (a.spc.tag) {  // species{add_tag}: the tags follow their particles
#pragma unroll 1
      for (int i0 = 0; i0 < n; i0 += PB * NTHREADS) {
        int cellv[PB], rk[PB];
        long long vt[PB];
#pragma unroll
        for (int k = 0; k < PB; k++) {
          const int i = i0 + tid + k * NTHREADS;
          if (i < n) {
            cellv[k] = a.scc.cell[so + i];
            rk[k] = a.scc.cell[a.scc.stride + so + i];
            vt[k] = a.scc.tag[so + i];
          }
        }
#pragma unroll
        for (int k = 0; k < PB; k++) {
          const int i = i0 + tid + k * NTHREADS;
          if (i >= n) continue;
          if (cellv[k] < 0) {
            if (base >= 0) a.mvc.tag[rk[k]] = vt[k];
          } else {
            int l[DIM];
#pragma unroll
            for (int d = 0; d < DIM; d++) l[d] = cell_l(cellv[k], d);
            a.spc.tag[off + hist[linear_cell<DIM>(l)] + rk[k]] = vt[k];
          }
        }
      }
    }
    if (pre_on) {
      int *cs_next = cs2 + (cur ^ 1) * P::CS_STRIDE;
      cs_next[tid] = pre;
      if (tid == 0) cs_next[G::NCELL] = pre_last;
    }
    if (P::TMA_J && tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();  // sort done (offsets in hist free), J tile read by the TMA
    for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) jt[i] = T(0);
    hist[tid] = 0;
    if (tid == 0) s_mcnt = 0;
    cur ^= 1;
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

template <typename T> void push_tile_shape(int dim, int order, int *box, int *jbox, int *smem_bytes) {
#define CASE(D, O)                                         \
  if (dim == D && order == O) {                            \
    using G = Tile<D, O, T>;                               \
    box[0] = G::EXP; box[1] = G::EY; box[2] = G::EZ;       \
    jbox[0] = G::JX; jbox[1] = G::JY; jbox[2] = G::JZ;     \
    *smem_bytes = Plan<D, O, T>::SMEM_BYTES;               \
    return;                                                \
  }
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
}
template void push_tile_shape<float>(int, int, int *, int *, int *);
template void push_tile_shape<double>(int, int, int *, int *, int *);

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
