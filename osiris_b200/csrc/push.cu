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
// Spline orders 1-3 take this path (struct Plan).  The 3D cubic stencil (3 x 48 sums) is accumulated
// one current component per walk over the cell's trajectories, and its J tile reuses the E/B tile's
// shared memory once the push is over (Plan::ALIAS_J), so that a tile of 2048 particles is staged at
// once; a tile with more particles than the staging buffer holds is processed in pieces.  Order 4
// loads its particles from global memory and adds every trajectory piece to the J tile directly.
#include <cuda.h>

#include "common.cuh"
#include "push.cuh"

namespace osgpu {

using namespace pic;

// the per-cell walk of the deposit as a lambda (1) or expanded in place (0), per dimensionality: the
// same statements, different register allocation -- measured per case (profiles/r2_notes.md)
// (same box: 2D linear 2.32 -> 2.19 ms and 3D cubic 7.44 -> 7.16 ms as a lambda, 3D quadratic 12.19 -> 12.22)
#ifndef OSGPU_WALK_LAMBDA_FOR
#define OSGPU_WALK_LAMBDA_FOR(DIM, ORDER) ((DIM) == 2 || (ORDER) == 3)
#endif
#ifndef OSGPU_BUF3D64
#define OSGPU_BUF3D64 151680
#endif
#ifndef OSGPU_BUF2D
#define OSGPU_BUF2D 94208
#endif
#ifndef OSGPU_HALO_JOBS
#define OSGPU_HALO_JOBS 1
#endif
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
__device__ __forceinline__ void bulk_load(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
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
  // orders 1-3: the particles of a tile are staged in shared memory by TMA bulk copies and the
  // in-cell piece of every sorted particle is accumulated in registers per cell (12 / 54 values in
  // 3D; the 144 of the 3D cubic scheme one current component -- 48 values -- per pass over the cell's
  // trajectories); order 4 would need 100 registers per component: it loads from global memory and
  // deposits directly
  static constexpr bool FAST = ORDER <= 3;
  static constexpr int NPASS = (ORDER == 3 && (DIM == 3 || sizeof(T) == 8)) ? 3 : 1;  // (2D cubic fp64: 40 sums = 80 registers)
  // 3D cubic fp32: the J tile (22 KB) lives where the E/B tile (27 KB) was -- E/B is dead once a piece is
  // pushed, J is born when its deposit starts -- which leaves room for 2624 staged particles: a tile of 2048
  // plus arrivals is pushed and deposited in ONE piece (with both tiles resident it took two, and the second,
  // ~260 particles in ~30 cells, kept one warp busy while seven waited: 43 % of the samples at CTA barriers)
  static constexpr bool ALIAS_J = FAST && DIM == 3 && ORDER == 3 && sizeof(T) == 4;
  static constexpr int NWARP = NTHREADS / 32;
  static constexpr int NC = DIM + 4;            // staged columns: x(D) u(3) q, blocked by 32 slots
  static constexpr int CHUNK_WORDS = NC * 32;   // one block of 32 slots
  static constexpr int XW = NC + 1;             // listed piece: p0[D], p1[D], q, (vz), link
  static constexpr int QN = 64;                 // crosser queue entries per warp (ints)
  static constexpr int QWORDS = QN * 4 / (int)sizeof(T);
  // one buffer per CTA: the staged particles of a piece from the bottom (after the push the
  // momentum slots hold the end point of the trajectory), above them one region per warp with the
  // list of pieces that belong to other cells of the tile and the queue of particles to split, and
  // at the top the CTA's list of pieces that lie in halo cells
  // (2 CTAs per SM: the padded 3D quadratic E/B tile takes 2304 B more; 3D cubic fp32: see ALIAS_J)
  static constexpr int BUF_BYTES = !FAST ? 0
      : ALIAS_J ? 60672 + 22656
      // 3D fp64 runs one CTA per SM (255 registers per thread): the whole 227 KB are its own, enough to stage a
      // tile of 2048 + arrivals in one piece (with the two-CTA buffer a tile took two to three pieces)
      : (DIM == 3 && sizeof(T) == 8) ? (ORDER == 3 ? 127616 : OSGPU_BUF3D64)
      // 2D fp32: the tiles are small (E/B + J < 17 KB), the buffer takes what two CTAs per SM leave: 3600 slots
      // instead of 2784 (64 particles per cell fill a 256-cell tile with 16384: five pieces instead of six)
      : (DIM == 2 && sizeof(T) == 4) ? OSGPU_BUF2D
      : 74752 - (OSGPU_PAD_EY && DIM == 3 && ORDER == 2 && sizeof(T) == 4 ? 2304 : 0);
  static constexpr int BUF_WORDS = BUF_BYTES / (int)sizeof(T);
  static constexpr int XMIN = 20;  // list entries per warp always available
  static constexpr int HCAP = OSGPU_PAD_EY ? 64 : 96;  // CTA-wide list of pieces in halo cells (top of the buffer)
  // (ALIAS_J: at most 2400 of the 2624 slots that fit, so that a full piece still leaves its crossers' pieces
  // ~400 list entries instead of 180 and they do not end up on the global-memory path)
  static constexpr int TCAP_FIT = ((BUF_WORDS - NWARP * (XMIN * XW + QWORDS) - HCAP * XW) / NC) & ~31;
  static constexpr int TCAP = FAST ? (ALIAS_J && TCAP_FIT > 2400 ? 2400 : TCAP_FIT) : (1 << 30);
  static constexpr int NROUND = FAST ? (TCAP + NTHREADS - 1) / NTHREADS : 1;
  static_assert(NROUND <= 16, "one mbarrier (and one phase bit) per staged round");
  static_assert(!FAST || TCAP < 4096, "queue entries keep the slot in 12 bits");
  static constexpr int NACC = (NPASS == 1 ? 3 : 1) * G::NP * G::NP * (DIM == 3 ? G::NP : 1);
  static constexpr int OFF_JT = ALIAS_J ? 0 : 6 * G::ESTRIDE * (int)sizeof(T);
  static_assert(!ALIAS_J || 3 * G::JVOL <= 6 * G::ESTRIDE, "the J tile fits in the E/B tile's memory");
  static constexpr int OFF_BUF = round_up((ALIAS_J ? 6 * G::ESTRIDE : 6 * G::ESTRIDE + 3 * G::JVOL) * (int)sizeof(T), 128);
  static constexpr int OFF_CS = OFF_BUF + BUF_BYTES;
  static constexpr int CS_STRIDE = round_up(G::NCELL + 1, 4);
  static constexpr int OFF_HIST = OFF_CS + 2 * CS_STRIDE * 4;
  // J tile -> HBM with one TMA tensor reduce-add per component (fp32)
  static constexpr bool TMA_J = sizeof(T) == 4;
  static constexpr int OFF_HEADS = OFF_HIST + G::NCELL * 4;
  static constexpr int SMEM_BYTES = OFF_HEADS + G::NCELL * 4;
  // all instantiations but 3D fp64 are meant to run two CTAs per SM: 233472 bytes per SM, 1024 reserved
  // per CTA, < 512 bytes of static shared variables in the kernel
  static_assert((sizeof(T) == 8 && DIM == 3) || SMEM_BYTES + 512 + 1024 <= 233472 / 2, "two CTAs per SM");
  static_assert(SMEM_BYTES + 512 <= 232448, "dynamic shared memory of one CTA");
  PIC_HD static int acc_index(int comp, int k1, int k2, int k3) {  // (one component per pass: comp does not count)
    return (((NPASS == 1 ? comp : 0) * (DIM == 3 ? G::NP : 1) + k3) * G::NP + k2) * G::NP + k1;
  }
  // word offset of staged slot m (column 0) inside the buffer
  PIC_HD static int slot(int m) { return (m >> 5) * CHUNK_WORDS + (m & 31); }
};

template <int DIM> __device__ __forceinline__ int linear_cell(const int *l) {
  using G = Tile<DIM, 1, float>;
  if constexpr (DIM == 2) return l[0] + G::TX * l[1];
  else return l[0] + G::TX * (l[1] + G::TY * l[DIM - 1]);
}

// Optimistic shared-memory add: fp32/fp64 atomicAdd on shared memory is an ATOMS.CAST.SPIN loop
// on sm_100a, so independent updates are issued as "read the old values, then one compare-and-swap
// each"; an update that lost a race falls back to the ordinary atomicAdd.
__device__ __forceinline__ bool cas_add(float *p, float old, float v) {
  const int o = __float_as_int(old);
  return atomicCAS(reinterpret_cast<int *>(p), o, __float_as_int(old + v)) == o;
}
__device__ __forceinline__ bool cas_add(double *p, double old, double v) {
  const unsigned long long o = (unsigned long long)__double_as_longlong(old);
  return atomicCAS(reinterpret_cast<unsigned long long *>(p), o,
                   (unsigned long long)__double_as_longlong(old + v)) == o;
}
// Flush of a cell's stencil accumulators into the shared J tile.  One batch = the same stencil
// point (k1, k2, k3) of the three components: they live in three different arrays, and the lanes of
// a warp hold different cells, so no two updates of a batch can alias inside the warp and the
// optimistic compare-and-swap only ever loses against another warp.  (Batches of a whole x-y plane
// of one component made lanes that own neighbouring cells invalidate each other's old values:
// about half of the lanes fell back to the spin loop on every entry -- profiles/r1_notes.md.)
template <int DIM, int ORDER, typename T, class P, int COMP = -1>
__device__ __forceinline__ void flush_stencil(const T *acc, T *pj) {
  using G = Tile<DIM, ORDER, T>;
  constexpr int NP = G::NP;
  constexpr int NZ = (DIM == 3) ? NP : 1;
#pragma unroll
  for (int k3 = 0; k3 < NZ; k3++)
#pragma unroll
    for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
      for (int k1 = 0; k1 < NP; k1++) {
        const bool on[3] = {(COMP < 0 || COMP == 0) && k1 < NP - 1, (COMP < 0 || COMP == 1) && k2 < NP - 1,
                            (COMP < 0 || COMP == 2) && (DIM == 2 || k3 < NP - 1)};
        T *p[3];
        T old[3];
#pragma unroll
        for (int c = 0; c < 3; c++) {
          p[c] = pj + c * G::JVOL + k1 + G::JX * (k2 + G::JY * k3);
          if (on[c]) old[c] = *reinterpret_cast<volatile T *>(p[c]);
        }
#pragma unroll
        for (int c = 0; c < 3; c++) {
          if (!on[c]) continue;
          const T v = acc[P::acc_index(c, k1, k2, k3)];
          if (!cas_add(p[c], old[c], v)) atomicAdd(p[c], v);
        }
      }
}

// ACCEL: the beam-initialisation pusher (accelerate_deposit, spec/os-spec-accel.f90) instead of
// interpolation + Boris; its own instantiation, because as a run-time branch its live values cost the
// standard path 5 % (profiles/r2_notes.md)
template <int DIM, int ORDER, typename T, bool ACCEL>
__global__ void __launch_bounds__(NTHREADS, (sizeof(T) == 8 && DIM == 3) ? 1 : 2)
    push_deposit_kernel(const __grid_constant__ TmaMaps maps, const PushArgs<T> a) {
  using G = Tile<DIM, ORDER, T>;
  using P = Plan<DIM, ORDER, T>;
  constexpr int NC = P::NC;
  extern __shared__ __align__(128) unsigned char smem_raw[];
  T *eb = reinterpret_cast<T *>(smem_raw);
  T *jt = reinterpret_cast<T *>(smem_raw + P::OFF_JT);
  T *buf = reinterpret_cast<T *>(smem_raw + P::OFF_BUF);
  int *cs2 = reinterpret_cast<int *>(smem_raw + P::OFF_CS);  // two arrays: current / next tile
  int *hist = reinterpret_cast<int *>(smem_raw + P::OFF_HIST);
  int *heads = reinterpret_cast<int *>(smem_raw + P::OFF_HEADS);  // per cell: list of foreign pieces
  __shared__ __align__(8) uint64_t mbar;                 // E/B tile
  __shared__ __align__(8) uint64_t mbar_p[P::NROUND];    // staged particles, one per round of 256
  __shared__ int s_tile, s_n, s_ns, s_mcnt, s_base, s_nh, s_nx, s_arrive, s_sortc, s_pushc;
  __shared__ long long s_off, s_next_off;
  __shared__ int s_next[3];
  __shared__ int s_warp[NTHREADS / 32];
  __shared__ double s_ene[NTHREADS / 32];

  const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const auto atomic_add = [](T *p, T v) { atomicAdd(p, v); };

  // this CTA's scratch range (new particle states in the old order, until the sort)
  const long long so = (long long)blockIdx.x * a.scr_stride;

  // Tile bookkeeping is pipelined: while a tile is pushed, one thread has already taken the CTA's
  // next tile from the device counter; as soon as the deposit is over (E/B tile and particle
  // buffer free) the same thread starts the next tile's TMA loads -- fields and particles -- which
  // then overlap the J write-back and the sort of the current tile.
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
  // one thread: bulk copies of `cnt` particles from slot `slot0` (a multiple of 32) into the
  // buffer, one copy and one mbarrier per round of 256 particles (whole blocks of 32 slots)
  auto stage_particles = [&](long long slot0, int cnt) {
    const int nch = (cnt + 31) >> 5;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic writes to the buffer
    const T *src = a.spc.base + (slot0 >> 5) * P::CHUNK_WORDS;
    for (int r = 0; 8 * r < nch; r++) {
      const uint32_t bytes = (uint32_t)min(8, nch - 8 * r) * (uint32_t)(P::CHUNK_WORDS * sizeof(T));
      mbar_expect_tx(&mbar_p[r], bytes);
      bulk_load(buf + 8 * r * P::CHUNK_WORDS, src + 8 * r * P::CHUNK_WORDS, bytes, &mbar_p[r]);
    }
  };
  auto prefetch_particles = [&](int n, long long off) {  // one thread: ask L2 for a tile's particles
    l2_prefetch(a.spc.base + (off >> 5) * P::CHUNK_WORDS, (uint32_t)(((n + 31) >> 5) * P::CHUNK_WORDS * sizeof(T)));
    l2_prefetch(a.spc.cell + off, (uint32_t)(((size_t)n * 4 + 15) & ~(size_t)15));
  };

  if (tid == 0) {
    mbar_init(&mbar, 1);
#pragma unroll
    for (int r = 0; r < P::NROUND; r++) mbar_init(&mbar_p[r], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    s_mcnt = 0;
    s_nh = 0;
    s_nx = 0;
    s_pushc = 0;
    s_arrive = 0;
    s_sortc = 0;
  }
  if constexpr (!P::ALIAS_J)
    for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) jt[i] = T(0);
  hist[tid] = 0;
  heads[tid] = 0;
  __syncthreads();
  if (tid == 32) {
    int n = 0, ns = 0;
    long long off = 0;
    const int t = take_tile(&n, &ns, &off);
    s_tile = t; s_n = n; s_ns = ns; s_off = off;
    if (t < a.tl.ntile) {
      stage_fields(t);
      if (P::FAST) stage_particles(off, min(n, P::TCAP));
    }
  }
  uint32_t phase = 0, pph = 0;  // mbarrier parities: E/B tile, staged rounds (one bit each)
  T ene_t = T(0);  // per-tile partial in particle precision (the reference sums a chunk in REAL)
  if (lane == 0) s_ene[wid] = 0.0;
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

    int o0[3], tlim[3], nlive = 0;
    tile_origin(tile, o0);
#pragma unroll
    for (int d = 0; d < 3; d++) tlim[d] = min(a.tl.T[d], a.g.nx[d] - o0[d]);

    // the CTA's next tile: taken now, used at the end
    if (tid == 32) {
      int nx_n = 0, nx_ns = 0;
      long long nx_off = 0;
      const int nx_tile = take_tile(&nx_n, &nx_ns, &nx_off);
      if (!P::FAST && nx_tile < a.tl.ntile) prefetch_particles(nx_n, nx_off);
      s_next[0] = nx_tile; s_next[1] = nx_n; s_next[2] = nx_ns; s_next_off = nx_off;
    }
    mbar_wait(&mbar, phase);
    phase ^= 1;

    // ---- 1..3: push, then deposit; every trajectory piece inside the tile is routed to the lane
    // that owns its cell, so its current is accumulated in registers ----------------------------
    {
      constexpr unsigned FULL = 0xffffffffu;
      const unsigned lt_mask = (1u << lane) - 1u;
      // The tile's particles -- [0, ns) sorted by cell, [ns, n) the unsorted tail -- are pushed in
      // pieces of <= TCAP; inside a piece the rounds of 32 are dealt round-robin to the warps, so
      // every warp has the same number of rounds whatever its cells hold.
      int qa = 0;
      bool more = false;
      do {
        const int qb = min(n, qa + P::TCAP);
        const int tot = qb - qa;                          // particles of this piece
        const int np = max(0, min(qb, ns) - qa);          // of which sorted: their trajectories are walked per cell
        const int pa = qa, pb = qa + np;                  // sorted particles [pa, pb)
        const int nrounds = (tot + NTHREADS - 1) / NTHREADS;
        more = qb < n;
        // above the staged particles: the CTA's pool of trajectory pieces that lie in other cells
        // of the tile (each linked to its cell), then per warp the queue of particles whose
        // trajectory is split, and at the top the pieces that lie in halo cells
        const int nchunks = (tot + 31) >> 5;
        const int xbase = NC * 32 * nchunks;
        const int xcap = P::FAST ? (P::BUF_WORDS - P::HCAP * P::XW - P::NWARP * P::QWORDS - xbase) / P::XW : 0;
        T *xl = buf + xbase;  // SoA with stride xcap
        int *queue = reinterpret_cast<int *>(buf + P::BUF_WORDS - P::HCAP * P::XW - (P::NWARP - wid) * P::QWORDS);
        T *hl = buf + P::BUF_WORDS - P::HCAP * P::XW;  // halo pieces, SoA with stride HCAP
        int ncq = 0;  // (warp-uniform) queue length

        // split the queued trajectories [qbase, qbase + cnt) of this warp, one per lane: the piece
        // in the particle's own cell goes back to its staged slot (sorted particles), the others to
        // the list of their cell, or straight into the J tile when they lie in the halo
        auto split_queued = [&](int qbase, int cnt) {
          __syncwarp();
          const bool have = lane < cnt;
          Seg<DIM, T> seg[DIM + 1];
          int nseg = 0;
          bool tail = false;
          T q = T(0);
          if (have) {
            const int ent = queue[qbase + lane];
            const int m = ent & 4095, lc = (ent >> 12) & 255;
            tail = (ent >> 20) & 1;
            T *pb_ = buf + P::slot(m);
            int di[3] = {0, 0, 0};
#pragma unroll
            for (int d = 0; d < DIM; d++) {
              seg[0].p0[d] = pb_[32 * d];
              seg[0].p1[d] = pb_[32 * (DIM + d)];
              di[d] = ntrim<T>(seg[0].p1[d]);
            }
            q = pb_[32 * (DIM + 3)];
            if (DIM == 2) seg[0].vz = pb_[32 * (DIM + 2)];
            seg[0].c[0] = lc % G::TX;
            seg[0].c[1] = (lc / G::TX) % G::TY;
            if (DIM == 3) seg[0].c[DIM - 1] = lc / (G::TX * G::TY);
            nseg = split_trajectory<DIM, T>(seg, di);
            if (!tail) {
#pragma unroll
              for (int d = 0; d < DIM; d++) pb_[32 * (DIM + d)] = seg[0].p1[d];
              if (DIM == 2) pb_[32 * (DIM + 2)] = seg[0].vz;
            }
          }
          unsigned direct = 0;
#pragma unroll
          for (int k = 0; k <= DIM; k++) {
            const bool has = have && (k < nseg) && (k > 0 || tail);
            bool inside = true;
            int lcell = 0;
            if (has) {
#pragma unroll
              for (int d = 0; d < DIM; d++) inside &= (seg[k].c[d] >= 0) && (seg[k].c[d] < (d == 0 ? G::TX : (d == 1 ? G::TY : G::TZ)));
              lcell = linear_cell<DIM>(seg[k].c);
            }
            const unsigned im = __ballot_sync(FULL, has && inside);
            const unsigned om = __ballot_sync(FULL, has && !inside);
            int xb = 0, hb = 0;
            if (im) {  // (warp-uniform) reserve room in the CTA's lists
              if (lane == 0) xb = atomicAdd(&s_nx, __popc(im));
              xb = __shfl_sync(FULL, xb, 0);
            }
            if (om) {
              if (lane == 0) hb = atomicAdd(&s_nh, __popc(om));
              hb = __shfl_sync(FULL, hb, 0);
            }
            if (has) {
              const int e = inside ? xb + __popc(im & lt_mask) : hb + __popc(om & lt_mask);
              if (inside ? e < xcap : e < P::HCAP) {
                T *dl = inside ? xl : hl;
                const int st = inside ? xcap : P::HCAP;
#pragma unroll
                for (int d = 0; d < DIM; d++) {
                  dl[d * st + e] = seg[k].p0[d];
                  dl[(DIM + d) * st + e] = seg[k].p1[d];
                }
                dl[(DIM + 3) * st + e] = q;
                if (DIM == 2) dl[(DIM + 2) * st + e] = seg[k].vz;
                int link;
                if (inside) {
                  link = atomicExch(&heads[lcell], e + 1);  // push on the cell's list
                } else {
                  link = 0;  // cell coordinates -1..T in 5-bit fields (exact as a float)
#pragma unroll
                  for (int d = 0; d < DIM; d++) link |= (seg[k].c[d] + 1) << (5 * d);
                }
                dl[NC * st + e] = (T)link;
              } else {
                direct |= 1u << k;
              }
            }
          }
          if (direct) {
            if constexpr (P::ALIAS_J) {
              // (the shared J tile does not exist during the push: pieces the lists have no room for go to the
              // grid in global memory, addressed through the J-tile position the statement computes)
              deposit_segments<DIM, ORDER, T>(seg, nseg, direct, q, jt, a.k, [&](T *p_, T v) {
                int org[3];
                tile_origin(tile, org);
                const int o_ = (int)(p_ - jt), cc = o_ / G::JVOL;
                int r_ = o_ - cc * G::JVOL;
                const int kx = r_ % G::JX; r_ /= G::JX;
                const int q0 = org[0] + a.g.gl[0] + G::HALO_LO + kx;
                const int q1 = org[1] + a.g.gl[1] + G::HALO_LO + r_ % G::JY;
                const int q2 = org[2] + a.g.gl[2] + G::HALO_LO + r_ / G::JY;
                if (q0 >= 0 && q0 < a.g.S[0] && q1 >= 0 && q1 < a.g.S[1] && q2 >= 0 && q2 < a.g.S[2])
                  atomicAdd(&a.J[cc][a.g.at(q0, q1, q2)], v);
              });
            } else {
              deposit_segments<DIM, ORDER, T>(seg, nseg, direct, q, jt, a.k, atomic_add);
            }
          }
          __syncwarp();
        };

        // -- 1. push, 32 particles per warp and round; the warps take their rounds from a counter
        // (the two or three warps that share a scheduler do not advance at the same pace) --------
        const auto next_round = [&]() {
          int j = 0;
          if (lane == 0) j = atomicAdd(&s_pushc, 1);
          return __shfl_sync(FULL, j, 0);
        };
        int j = next_round(), jn = 0;
        int ncell_ = 0;
        if (32 * j + lane < tot) ncell_ = a.spc.cell[off + qa + 32 * j + lane];
        T *const sc0 = a.scc.base + ((so + qa) >> 5) * P::CHUNK_WORDS + lane;
        for (;; j = jn) {
          const bool done = j >= nchunks;  // (uniform) one more trip for what is left in the queue
          bool cross = false;
          int lc_old = 0;
          const int m = 32 * j + lane;
          const bool active = !done && m < tot, parked = m < np;
          if (!done) {
          const int r = j >> 3;
          T *sc = sc0 + (size_t)j * P::CHUNK_WORDS;
          const int cell = ncell_;
          jn = next_round();
          if (32 * jn + lane < tot) ncell_ = a.spc.cell[off + qa + 32 * jn + lane];
          T x[DIM], u[3], q = T(0);
          T *pb_ = buf + P::slot(m);
          if constexpr (P::FAST) {
            mbar_wait(&mbar_p[r], (pph >> r) & 1u);
            if (active) {
#pragma unroll
              for (int d = 0; d < DIM; d++) x[d] = pb_[32 * d];
#pragma unroll
              for (int d = 0; d < 3; d++) u[d] = pb_[32 * (DIM + d)];
              q = pb_[32 * (DIM + 3)];
            }
          } else {
            if (active) {
              const T *gp = a.spc.base + a.spc.blk(off + qa + m);
#pragma unroll
              for (int d = 0; d < DIM; d++) x[d] = gp[32 * d];
#pragma unroll
              for (int d = 0; d < 3; d++) u[d] = gp[32 * (DIM + d)];
              q = gp[32 * (DIM + 3)];
            }
          }
          if (active) {
            int l[DIM], di[3] = {0, 0, 0};
#pragma unroll
            for (int d = 0; d < DIM; d++) l[d] = cell_l(cell, d);
            lc_old = linear_cell<DIM>(l);
            Seg<DIM, T> seg[DIM + 1];  // (only the direct path keeps the pieces here)
#pragma unroll
            for (int d = 0; d < DIM; d++) {
              seg[0].p0[d] = x[d];
              seg[0].c[d] = l[d];
            }
            T vz = T(0);
            T en, qd = q;  // qd: the charge the deposit sees
            if constexpr (ACCEL) {  // beam initialisation: free streaming along x1, reduced charge
              en = accel_move<DIM, T>(x, u, l, a.k.dt_dx[0], a.accel.per_particle ? a.accel.u_frac * u[0] : a.accel.u_frac * a.accel.ufl1,
                                      seg[0].p1, di, &vz);
              qd = q * a.accel.q_frac;
            } else {
              en = push_move<DIM, ORDER, T>(x, u, l, eb, a.k, seg[0].p1, di, &vz);
            }
            ene_t = fmadd(q, en, ene_t);
            bool moved = false;
            int ncell = 0;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
              moved |= (unsigned)l[d] >= (unsigned)tlim[d];
              ncell |= (l[d] + 1) << (d * CELL_BITS);
              cross |= di[d] != 0;
            }
            if (DIM == 2) ncell |= 1 << (2 * CELL_BITS);
            int rank = 0;
            if (moved) ncell |= CELL_DEAD;
            else rank = atomicAdd(&hist[linear_cell<DIM>(l)], 1);
#pragma unroll
            for (int d = 0; d < DIM; d++) sc[32 * d] = x[d];
#pragma unroll
            for (int d = 0; d < 3; d++) sc[32 * (DIM + d)] = u[d];
            sc[32 * (DIM + 3)] = q;
            const long long rs = so + qa + m;
            a.scc.cell[rs] = ncell;
            a.scr_rank[rs] = rank;
            if (a.spc.tag) a.scc.tag[rs] = a.spc.tag[off + qa + m];
            if constexpr (P::FAST) {
              // park the end point of the trajectory (old cell's frame) where the momentum was
#pragma unroll
              for (int d = 0; d < DIM; d++) pb_[32 * (DIM + d)] = seg[0].p1[d];
              if (DIM == 2) pb_[32 * (DIM + 2)] = vz;
              if constexpr (ACCEL) pb_[32 * (DIM + 3)] = qd;  // (the deposit reads the charge from the staged slot)
              cross |= !parked;
            } else {
              if (DIM == 2) seg[0].vz = vz;
              const int nseg = split_trajectory<DIM, T>(seg, di);
              deposit_segments<DIM, ORDER, T>(seg, nseg, 0xfu, qd, jt, a.k, atomic_add);
            }
          }
          }  // !done
          if constexpr (P::FAST) {
            // trajectories that leave their cell (and every tail particle) are queued and split
            // 32 at a time
            const unsigned qm = __ballot_sync(FULL, cross);
            if (cross) queue[ncq + __popc(qm & lt_mask)] = m | (lc_old << 12) | (parked ? 0 : 1 << 20);
            ncq += __popc(qm);
            if (ncq >= 32 || (done && ncq > 0)) {
              const int take = min(ncq, 32);
              split_queued(ncq - take, take);
              ncq -= take;
            }
          }
          if (done) break;
        }
        if constexpr (P::FAST) pph ^= (1u << nrounds) - 1u;

        __syncthreads();  // every warp's trajectories and lists are complete
        if constexpr (P::ALIAS_J) {  // E/B is dead: its memory becomes the (zeroed) J tile; barriers follow before any flush
          for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) jt[i] = T(0);
        }

        if (!more) {
          // all particles of the tile are pushed: new-cell histogram -> cell offsets of the sorted
          // tile, room for the leavers in the mover buffer.  Done here, before the deposit, so that
          // every warp can go on to the sort as soon as its own cells are deposited.
          const int start = block_scan(hist[tid], &nlive, s_warp);
          hist[tid] = start;  // counts -> offsets
          if (tid == 0) {
            const int nmov = n - nlive;
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
          int *g_cs = a.cell_start + (size_t)tile * (G::NCELL + 1);
          g_cs[tid] = start;
          if (tid == 0) g_cs[G::NCELL] = nlive;
          __syncthreads();  // offsets and s_base visible to the sort
        }

        if constexpr (P::FAST) {
          // -- 2. deposit: one lane per cell walks the cell's trajectories and its list of foreign
          // pieces; the whole stencil is accumulated in registers and flushed once.  The cells are
          // dealt to the lanes in the order of their particle counts, so the 32 lanes of a warp
          // have (almost) the same number of trajectories to walk.
          int *bucket = cs2 + (cur ^ 1) * P::CS_STRIDE;  // (the next tile's offsets are stored at the tile's end)
          unsigned char *order = reinterpret_cast<unsigned char *>(bucket + 64);
          if (tid < 64) bucket[tid] = 0;
          __syncthreads();
          int cnt_own = (np > 0) ? max(0, min(cs[tid + 1], pb) - max(cs[tid], pa)) : 0;
          if (heads[tid] != 0) cnt_own++;
          const int bk = 63 - min(cnt_own, 63);  // fullest cells first
          const int r_in = atomicAdd(&bucket[bk], 1);
          __syncthreads();
          int btot;
          const int bbase = block_scan(tid < 64 ? bucket[tid] : 0, &btot, s_warp);
          __syncthreads();
          if (tid < 64) bucket[tid] = bbase;
          __syncthreads();
          order[bucket[bk] + r_in] = (unsigned char)tid;
          __syncthreads();
          const int c = order[tid];
          int i = (np > 0) ? max(cs[c], pa) : 0;
          const int ib = (np > 0) ? min(cs[c + 1], pb) : 0;
          int id = heads[c];
          if (id) heads[c] = 0;
          // -- 3. pieces in halo cells (owned by a neighbouring tile) follow as extra jobs of the
          // first threads, one piece each, through the same accumulate-and-flush code ------------
          const int nh = OSGPU_HALO_JOBS ? min(s_nh, P::HCAP) : 0;
          int he = tid;  // next halo piece of this thread
          bool own = true;
          for (;;) {
            int pc = 0;
            if (own) {
              if (!(i < ib || id)) { own = false; continue; }
            } else {
              if (he >= nh) break;
              pc = (int)hl[NC * P::HCAP + he];
            }
            // one walk over the job's trajectories: all three current components at once, or (3D cubic) one
            // component per walk, restarting from the same place.  (A macro, not a lambda: as a lambda the
            // single-pass instantiations came out 0.4 % slower -- profiles/r2_notes.md.)
#define OSGPU_WALK(COMP)                                                                                      \
  {                                                                                                           \
    T acc[P::NACC];                                                                                           \
    _Pragma("unroll") for (int k = 0; k < P::NACC; k++) acc[k] = T(0);                                        \
    bool halo = !own;                                                                                         \
    while (i < ib || id || halo) {                                                                            \
      const T *src;                                                                                           \
      int st;                                                                                                 \
      if (halo) { /* a piece in a halo cell */                                                                \
        src = hl + he;                                                                                        \
        st = P::HCAP;                                                                                         \
        halo = false;                                                                                         \
      } else if (i < ib) { /* a trajectory parked in its staged slot */                                       \
        src = buf + P::slot(i - pa);                                                                          \
        st = 32;                                                                                              \
        i++;                                                                                                  \
      } else { /* a listed piece */                                                                           \
        src = xl + (id - 1);                                                                                  \
        st = xcap;                                                                                            \
        id = (int)src[NC * st];                                                                               \
      }                                                                                                       \
      Seg<DIM, T> sg;                                                                                         \
      _Pragma("unroll") for (int d = 0; d < DIM; d++) {                                                       \
        sg.p0[d] = src[d * st];                                                                               \
        sg.p1[d] = src[(DIM + d) * st];                                                                       \
      }                                                                                                       \
      const T q = src[(DIM + 3) * st];                                                                        \
      if (DIM == 2) sg.vz = src[(DIM + 2) * st];                                                              \
      segment_current<DIM, ORDER, T, COMP>(sg, q, a.k, [&](int comp, int k1, int k2, int k3, T f1, T f2) {    \
        T &r = acc[P::acc_index(comp, k1, k2, k3)];                                                           \
        r = fmadd(f1, f2, r);                                                                                 \
      });                                                                                                     \
    }                                                                                                         \
    int lc[3];                                                                                                \
    if (own) {                                                                                                \
      lc[0] = c % G::TX;                                                                                      \
      lc[1] = (c / G::TX) % G::TY;                                                                            \
      lc[2] = c / (G::TX * G::TY);                                                                            \
    } else {                                                                                                  \
      _Pragma("unroll") for (int d = 0; d < 3; d++) lc[d] = ((pc >> (5 * d)) & 31) - 1; /* -1..T */           \
    }                                                                                                         \
    T *pj = jt + jtile_base<DIM, ORDER, T>(lc);                                                               \
    flush_stencil<DIM, ORDER, T, P, COMP>(acc, pj);                                                           \
  }
            if constexpr (OSGPU_WALK_LAMBDA_FOR(DIM, ORDER)) {
              [[maybe_unused]] const int i_job = i, id_job = id;
              auto walk = [&](auto comp_c) {
                constexpr int COMP = decltype(comp_c)::value;
                if constexpr (P::NPASS > 1) {
                  i = i_job;
                  id = id_job;
                }
                OSGPU_WALK(COMP)
              };
              if constexpr (P::NPASS == 1) {
                walk(std::integral_constant<int, -1>{});
              } else {
                walk(std::integral_constant<int, 0>{});
                walk(std::integral_constant<int, 1>{});
                walk(std::integral_constant<int, 2>{});
              }
            } else if constexpr (P::NPASS == 1) {
              OSGPU_WALK(-1)
            } else {
              const int i_job = i, id_job = id;
              OSGPU_WALK(0)
              i = i_job;
              id = id_job;
              OSGPU_WALK(1)
              i = i_job;
              id = id_job;
              OSGPU_WALK(2)
            }
#undef OSGPU_WALK
            if (own) own = false;
            else he += NTHREADS;
          }
        }
#if !OSGPU_HALO_JOBS
        if constexpr (P::FAST) {
          const int nh = min(s_nh, P::HCAP);
          for (int e = tid; e < nh; e += NTHREADS) {
            Seg<DIM, T> sg;
#pragma unroll
            for (int d = 0; d < DIM; d++) {
              sg.p0[d] = hl[d * P::HCAP + e];
              sg.p1[d] = hl[(DIM + d) * P::HCAP + e];
            }
            const T q = hl[(DIM + 3) * P::HCAP + e];
            if (DIM == 2) sg.vz = hl[(DIM + 2) * P::HCAP + e];
            const int pc = (int)hl[NC * P::HCAP + e];
#pragma unroll
            for (int d = 0; d < DIM; d++) sg.c[d] = ((pc >> (5 * d)) & 31) - 1;
            deposit_segment<DIM, ORDER, T>(sg, q, jt, a.k, atomic_add);
          }
        }
#endif
        qa = qb;
        if (more) {
          __syncthreads();  // the next piece reuses the buffer
          if (tid == 0) {
            s_nh = 0;
            s_nx = 0;
            s_pushc = 0;
          }
          if (P::FAST && tid == 32) {
            if constexpr (P::ALIAS_J) {  // this piece's J goes out, then E/B comes back for the next piece
              int org[3];
              tile_origin(tile, org);
              asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
              for (int c = 0; c < 3; c++)
                tma_reduce_add_3d(&maps.m[6 + c], jt + c * G::JVOL, org[0] + a.g.gl[0] + G::HALO_LO,
                                  org[1] + a.g.gl[1] + G::HALO_LO, org[2] + a.g.gl[2] + G::HALO_LO);
              asm volatile("cp.async.bulk.commit_group;" ::: "memory");
              asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
              stage_fields(tile);
            }
            stage_particles(off + qa, min(n - qa, P::TCAP));
          }
          __syncthreads();  // (halo list counter reset before the next piece adds to it)
          if constexpr (P::ALIAS_J) {
            mbar_wait(&mbar, phase);
            phase ^= 1;
          }
        }
      } while (more);
    }
    {  // this warp's share of the tile's energy sum: to its double in shared memory
      T e = ene_t;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) e += __shfl_down_sync(0xffffffffu, e, o);
      if (lane == 0) s_ene[wid] += (double)e;
      ene_t = T(0);
    }

    // ---- 4. the last warp to finish its deposit sends the J tile to HBM and starts the next
    // tile's loads; nobody waits: the other warps are already sorting ---------------------------
    __syncwarp();
    int last = 0;
    if (lane == 0) {
      __threadfence_block();  // this warp's J-tile updates and buffer reads before the arrival
      last = atomicAdd(&s_arrive, 1) == NTHREADS / 32 - 1;
    }
    if (last) {  // (one thread of the CTA)
      __threadfence_block();
      s_arrive = 0;
      tile_origin(tile, o0);  // (recomputed: not kept in registers through the push)
      if constexpr (P::TMA_J) {
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
        if constexpr (P::ALIAS_J) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // J is read before E/B lands on it
      }
      if (s_next[0] < a.tl.ntile) {
        stage_fields(s_next[0]);
        if (P::FAST) stage_particles(s_next_off, min(s_next[1], P::TCAP));
      }
    }

    // ---- 5. sort by new cell: scratch -> particle buffer, leavers -> movers -----------------
    const int base = s_base;
    // cell offsets of the CTA's next tile: loaded now, stored below
    int pre = 0, pre_last = 0;
    const bool pre_on = P::FAST && s_next[0] < a.tl.ntile && s_next[2] > 0;
    if (pre_on) {
      const int *g_cs = a.cell_start + (size_t)s_next[0] * (G::NCELL + 1);
      pre = g_cs[tid];
      if (tid == 0) pre_last = g_cs[G::NCELL];
    }
    // The warps take runs of PB * 32 particles from a counter (a warp that finished its deposit
    // early sorts more).  One particle per thread in flight: the scratch range sits in L2 and latency is
    // what costs, but two per thread spill the loaded values (12.57 -> 12.22 ms per step with one,
    // profiles/r2_notes.md).
#ifndef SORT_PB
#define SORT_PB 1
#endif
    constexpr int PB = SORT_PB;
    const T *scr0 = a.scc.base + (so >> 5) * P::CHUNK_WORDS + lane;
    T *dstb = a.spc.base + (off >> 5) * P::CHUNK_WORDS;
    for (;;) {
      int i0 = 0;
      if (lane == 0) i0 = atomicAdd(&s_sortc, PB * 32);
      i0 = __shfl_sync(0xffffffffu, i0, 0);
      if (i0 >= n) break;
      int cellv[PB], dst[PB];
      T vx[PB][DIM], vp[PB][3], vq[PB];
      long long vt[PB];
#pragma unroll
      for (int k = 0; k < PB; k++) {
        const int i = i0 + 32 * k + lane;
        cellv[k] = 0;
        dst[k] = 0;
        vt[k] = 0;
        if (i < n) {
          cellv[k] = a.scc.cell[so + i];
          dst[k] = a.scr_rank[so + i];
          if (a.spc.tag) vt[k] = a.scc.tag[so + i];
          const T *sp_ = scr0 + (size_t)((i0 >> 5) + k) * P::CHUNK_WORDS;
#pragma unroll
          for (int d = 0; d < DIM; d++) vx[k][d] = sp_[32 * d];
#pragma unroll
          for (int d = 0; d < 3; d++) vp[k][d] = sp_[32 * (DIM + d)];
          vq[k] = sp_[32 * (DIM + 3)];
        }
      }
#pragma unroll
      for (int k = 0; k < PB; k++) {
        const int i = i0 + 32 * k + lane;
        if (i >= n) continue;
        if (cellv[k] < 0) {
          if (base < 0) continue;  // mover buffer overflow: reported through errflag
          const int m = base + atomicAdd(&s_mcnt, 1);
#pragma unroll
          for (int d = 0; d < DIM; d++) {
            a.mvc.base[d * a.mvc.stride + m] = vx[k][d];
            a.mvc.cell[d * a.mvc.stride + m] = o0[d] + cell_l(cellv[k] & ~CELL_DEAD, d) + 1;
          }
#pragma unroll
          for (int d = 0; d < 3; d++) a.mvc.base[(DIM + d) * a.mvc.stride + m] = vp[k][d];
          a.mvc.base[(DIM + 3) * a.mvc.stride + m] = vq[k];
          if (a.spc.tag) a.mvc.tag[m] = vt[k];
        } else {
          int l[DIM];
#pragma unroll
          for (int d = 0; d < DIM; d++) l[d] = cell_l(cellv[k], d);
          const int s = hist[linear_cell<DIM>(l)] + dst[k];  // slot inside the tile
          T *dp = dstb + P::slot(s);
#pragma unroll
          for (int d = 0; d < DIM; d++) dp[32 * d] = vx[k][d];
#pragma unroll
          for (int d = 0; d < 3; d++) dp[32 * (DIM + d)] = vp[k][d];
          dp[32 * (DIM + 3)] = vq[k];
          a.spc.cell[off + s] = cellv[k];
          if (a.spc.tag) a.spc.tag[off + s] = vt[k];
        }
      }
    }
    if (P::TMA_J && last) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    __syncthreads();  // sort done (offsets in hist free), every warp's deposit done, J tile read by the TMA
    if constexpr (!P::TMA_J) {
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
        jt[i] = T(0);
      }
    } else if constexpr (!P::ALIAS_J) {
      for (int i = tid; i < 3 * G::JVOL; i += NTHREADS) jt[i] = T(0);
    }
    hist[tid] = 0;
    if (pre_on) {
      int *cs_next = cs2 + (cur ^ 1) * P::CS_STRIDE;
      cs_next[tid] = pre;
      if (tid == 0) cs_next[G::NCELL] = pre_last;
    }
    if (tid == 0) {
      s_mcnt = 0;
      s_nh = 0;
      s_nx = 0;
      s_pushc = 0;
      s_sortc = 0;
      s_tile = s_next[0]; s_n = s_next[1]; s_ns = s_next[2]; s_off = s_next_off;
    }
    cur ^= 1;
  }

  // ---- energy: one atomic per CTA -----------------------------------------------------------
  __syncthreads();
  if (tid == 0) {
    double t = 0.0;
#pragma unroll
    for (int w = 0; w < NTHREADS / 32; w++) t += s_ene[w];
    if (t != 0.0) atomicAdd(a.ene, t);
  }
}

template <int DIM, int ORDER, typename T, bool ACCEL>
static int launch(const TmaMaps &maps, const PushArgs<T> &a, int nblocks, cudaStream_t st) {
  using P = Plan<DIM, ORDER, T>;
  auto kern = push_deposit_kernel<DIM, ORDER, T, ACCEL>;
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
  if (dim == D && order == O) return a.accel.on ? launch<D, O, T, true>(maps, a, nblocks, st) : launch<D, O, T, false>(maps, a, nblocks, st);
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
  set_error("unsupported dim/order %d/%d", dim, order);
  return OSGPU_ERR_INVALID;
}
template int launch_push_deposit<float>(int, int, const TmaMaps &, const PushArgs<float> &, int,
                                        cudaStream_t);
template int launch_push_deposit<double>(int, int, const TmaMaps &, const PushArgs<double> &, int,
                                         cudaStream_t);

// ---------------------------------------------------------------------------------------------
// list pusher: one thread per particle, fields and current in global memory
// ---------------------------------------------------------------------------------------------
template <typename T> struct SixPlanes { const T *eb[6]; T *j[3]; };
template <int DIM, int ORDER, typename T>
__global__ void __launch_bounds__(128)
    push_list_kernel(MoverView<T> in, MoverView<T> out, SixPlanes<T> f, Grid g, PushConst<T> k, double *ene, int *errflag) {
  constexpr int LP = Stencil<ORDER>::LP;
  const int n = min(*in.count, in.cap);
  const long long sy = g.P0, sz = (long long)g.P0 * g.S[1];
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    T x[DIM], u[3];
    int ix[DIM], di[3] = {0, 0, 0};
    Seg<DIM, T> seg[DIM + 1];
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      x[d] = in.x[d][i];
      ix[d] = in.ix[d][i];
      seg[0].p0[d] = x[d];
      seg[0].c[d] = ix[d] - 1 + g.gl[d];  // padded grid coordinate of the cell
    }
#pragma unroll
    for (int d = 0; d < 3; d++) u[d] = in.p[d][i];
    const T q = in.q[i];
    bool inside = true;  // the interpolation stencil (one more point below for the staggered components) must lie on the grid
#pragma unroll
    for (int d = 0; d < DIM; d++) inside &= (seg[0].c[d] + LP - 1 >= 0) && (seg[0].c[d] + LP + ORDER < g.S[d]);
    if (!inside) { atomicOr(errflag, ERRF_BAD_CELL); continue; }
    T vz = T(0);
    const T en = push_move_grid<DIM, ORDER, T>(x, u, seg[0].c, f.eb, sy, sz, k, seg[0].p1, di, &vz);
    if (DIM == 2) seg[0].vz = vz;
    atomicAdd(ene, (double)(q * en));
    const int nseg = split_trajectory<DIM, T>(seg, di);
#pragma unroll 1
    for (int sidx = 0; sidx < nseg; sidx++) {
      Seg<DIM, T> sg = seg[0];
#pragma unroll
      for (int m = 1; m <= DIM; m++)
        if (sidx == m) sg = seg[m];
      int b[3] = {0, 0, 0};
#pragma unroll
      for (int d = 0; d < DIM; d++) b[d] = sg.c[d] + LP;
      segment_current<DIM, ORDER, T>(sg, q, k, [&](int comp, int k1, int k2, int k3, T f1, T f2) {
        const int q0 = b[0] + k1, q1 = b[1] + k2, q2 = b[2] + k3;
        if (q0 >= 0 && q0 < g.S[0] && q1 >= 0 && q1 < g.S[1] && q2 >= 0 && q2 < g.S[2]) atomicAdd(&f.j[comp][g.at(q0, q1, q2)], f1 * f2);
        else atomicOr(errflag, ERRF_BAD_CELL);
      });
    }
    const int m = atomicAdd(out.count, 1);
    if (m >= out.cap) { atomicOr(errflag, ERRF_MOVER_OVERFLOW); continue; }
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      out.x[d][m] = x[d];
      out.ix[d][m] = ix[d] + di[d];
    }
#pragma unroll
    for (int d = 0; d < 3; d++) out.p[d][m] = u[d];
    out.q[m] = q;
    if (out.tag) out.tag[m] = in.tag ? in.tag[i] : -1;
  }
}
template <typename T>
int launch_push_list(int dim, int order, const MoverView<T> &in, const MoverView<T> &out, const T *const *eb6, T *const *j3,
                     const Grid &g, const pic::PushConst<T> &k, double *ene, int *errflag, cudaStream_t st) {
  SixPlanes<T> f;
  for (int c = 0; c < 6; c++) f.eb[c] = eb6[c];
  for (int c = 0; c < 3; c++) f.j[c] = j3[c];
#define CASE(D, O) \
  if (dim == D && order == O) push_list_kernel<D, O, T><<<148 * 4, 128, 0, st>>>(in, out, f, g, k, ene, errflag);
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
  OSGPU_CUDA(cudaGetLastError());
  return OSGPU_OK;
}
template int launch_push_list<float>(int, int, const MoverView<float> &, const MoverView<float> &, const float *const *,
                                     float *const *, const Grid &, const pic::PushConst<float> &, double *, int *, cudaStream_t);
template int launch_push_list<double>(int, int, const MoverView<double> &, const MoverView<double> &, const double *const *,
                                      double *const *, const Grid &, const pic::PushConst<double> &, double *, int *, cudaStream_t);

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
    auto kern = push_deposit_kernel<D, O, T, false>;                                                 \
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
