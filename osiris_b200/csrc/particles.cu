// particles.cu -- particle bookkeeping kernels: tile binning (upload / counting sort by tile),
// re-insertion of particles that changed tile, periodic wrap, migration pack/unpack, download.
//
// Reference counterparts: sort_species / generate_sort_idx / rearrange_species
// (source/spec/os-species.f03:269-702), update_boundary_species and helpers
// (source/spec/os-spec-boundary.f03:538-816, 1099-1290), pack/unpack of migrating particles
// (source/spec/os-spec-comm.f90:309-759).  The reference re-sorts by cell every n_sort steps and
// scans every particle for boundary crossings each step; here particles are kept binned by cell
// tile at all times and only the particles that left their tile are touched.
#include "common.cuh"
#include "particles.cuh"

namespace osgpu {

using namespace pic;

constexpr int GRID_STRIDE_BLOCKS = 148 * 8;

// tile index and packed tile-local cell of a 1-based local cell index
__device__ __forceinline__ bool locate(const int *ix, int dim, const Tiling &tl, const int *nx,
                                       int *tile, int *cell) {
  int t = 0, mul = 1, c = 0;
  bool ok = true;
  for (int d = 0; d < 3; d++) {
    int i0 = (d < dim) ? ix[d] - 1 : 0;
    ok &= (i0 >= 0) && (i0 < ((d < dim) ? nx[d] : 1));
    const int td = i0 / tl.T[d];
    t += mul * td;
    mul *= tl.nt[d];
    c |= (i0 - td * tl.T[d] + 1) << (d * CELL_BITS);
  }
  *tile = t;
  *cell = c;
  return ok;
}

// linear cell inside the tile, the order the push kernel keeps a tile sorted in (push.cu: linear_cell)
__device__ __forceinline__ int tile_cell_index(int cell, const Tiling &tl) {
  return cell_l(cell, 0) + tl.T[0] * (cell_l(cell, 1) + tl.T[1] * cell_l(cell, 2));
}

// Upload = counting sort by (tile, cell): the buffer comes out the way the push kernel leaves it,
// every tile sorted by cell with its cell offsets known, so the first push after an upload already
// runs the per-cell register accumulation (an unsorted tile sends every particle through the
// piece lists).  Pass 1: particles per cell.
// `shift`: cells subtracted from ix first (moving window, move_window_species spec/os-species.f03:171-229).
// A particle outside 1..nx is an error on upload; with `to_movers` (window shift) it is handed to the
// mover buffer with its shifted ix and meets the boundary conditions / migration there.
template <typename T>
__global__ void __launch_bounds__(256)
    count_kernel(AosParticles<T> in, long long np, int dim, Tiling tl, Grid g, Shift3 shift, int to_movers,
                 int *cell_cnt, int *errflag) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < np;
       i += (long long)gridDim.x * blockDim.x) {
    int ix[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) ix[d] = in.ix[dim * i + d] - shift.n[d];
    int t, c;
    if (locate(ix, dim, tl, g.nx, &t, &c)) atomicAdd(&cell_cnt[(size_t)t * TILE_NCELL + tile_cell_index(c, tl)], 1);
    else if (!to_movers) atomicOr(errflag, ERRF_BAD_CELL);
  }
}

// Pass 2, one CTA per tile: cell counts -> cell offsets (exclusive scan), tile count; the counts
// are cleared and serve as the scatter cursors of pass 3.
__global__ void __launch_bounds__(TILE_NCELL)
    cell_scan_kernel(int *cell_cnt, int *cell_start, int *tile_cnt, int *tile_nsorted) {
  __shared__ int s_warp[TILE_NCELL / 32];
  const int t = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
  const int v = cell_cnt[(size_t)t * TILE_NCELL + tid];
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int u = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += u;
  }
  if (lane == 31) s_warp[wid] = inc;
  __syncthreads();
  int base = 0, tot = 0;
#pragma unroll
  for (int w = 0; w < TILE_NCELL / 32; w++) {
    const int u = s_warp[w];
    if (w < wid) base += u;
    tot += u;
  }
  cell_start[(size_t)t * (TILE_NCELL + 1) + tid] = base + inc - v;
  cell_cnt[(size_t)t * TILE_NCELL + tid] = 0;
  if (tid == 0) {
    cell_start[(size_t)t * (TILE_NCELL + 1) + TILE_NCELL] = tot;
    tile_cnt[t] = tot;
    tile_nsorted[t] = tot;
  }
}

// Pass 3: every particle to tile_off[tile] + cell_start[tile][cell] + (its turn inside the cell)
template <typename T>
__global__ void __launch_bounds__(256)
    scatter_kernel(AosParticles<T> in, long long np, long long tag_base, int dim, Tiling tl, Grid g, Shift3 shift,
                   int to_movers, MoverView<T> mv, int *errflag, const int *tile_off, const int *cell_start,
                   int *cursor, SpeciesView<T> sp) {
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < np;
       i += (long long)gridDim.x * blockDim.x) {
    int ix[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) ix[d] = in.ix[dim * i + d] - shift.n[d];
    int t, c;
    if (!locate(ix, dim, tl, g.nx, &t, &c)) {
      if (to_movers) {
        const int m = atomicAdd(mv.count, 1);
        if (m >= mv.cap) { atomicOr(errflag, ERRF_MOVER_OVERFLOW); continue; }
        for (int d = 0; d < dim; d++) { mv.x[d][m] = in.x[dim * i + d]; mv.ix[d][m] = ix[d]; }
        for (int d = 0; d < 3; d++) mv.p[d][m] = in.p[3 * i + d];
        mv.q[m] = in.q[i];
        if (mv.tag) mv.tag[m] = in.tag ? in.tag[i] : tag_base + i;
      }
      continue;
    }
    const int lc = tile_cell_index(c, tl);
    const long long s = (long long)tile_off[t] + cell_start[(size_t)t * (TILE_NCELL + 1) + lc] +
                        atomicAdd(&cursor[(size_t)t * TILE_NCELL + lc], 1);
    for (int d = 0; d < dim; d++) sp.r(d, s) = in.x[dim * i + d];
    for (int d = 0; d < 3; d++) sp.r(dim + d, s) = in.p[3 * i + d];
    sp.r(dim + 3, s) = in.q[i];
    sp.cell[s] = c;
    if (sp.tag) sp.tag[s] = in.tag ? in.tag[i] : tag_base + i;
  }
}

// one CTA per tile: live slots -> AoS output rows out_off[tile] ...
template <typename T>
__global__ void __launch_bounds__(256)
    gather_kernel(SpeciesView<T> sp, const int *tile_off, const int *tile_cnt, const long long *out_off,
                  int dim, Tiling tl, int *oix, T *ox, T *op, T *oq, long long *otag, int by_tag) {
  const int tile = blockIdx.x;
  const int n = tile_cnt[tile];
  const long long off = tile_off[tile], o = out_off[tile];
  int tc[3], t = tile;
  tc[0] = t % tl.nt[0]; t /= tl.nt[0];
  tc[1] = t % tl.nt[1];
  tc[2] = t / tl.nt[1];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long s = off + i, r = by_tag ? sp.tag[s] : o + i;
    const int cell = sp.cell[s];
    for (int d = 0; d < dim; d++) {
      ox[dim * r + d] = sp.r(d, s);
      oix[dim * r + d] = tc[d] * tl.T[d] + cell_l(cell, d) + 1;
    }
    for (int d = 0; d < 3; d++) op[3 * r + d] = sp.r(dim + d, s);
    oq[r] = sp.r(dim + 3, s);
    if (otag) otag[r] = sp.tag ? sp.tag[s] : -1;
  }
}

// pending movers (not yet re-inserted, e.g. after a bare advance_deposit) -> AoS output rows
template <typename T>
__global__ void __launch_bounds__(256)
    mover_gather_kernel(MoverView<T> mv, long long out_base, int dim, int *oix, T *ox, T *op, T *oq,
                        long long *otag, int by_tag) {
  const int n = min(*mv.count, mv.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const long long r = by_tag ? mv.tag[i] : out_base + i;
    for (int d = 0; d < dim; d++) {
      ox[dim * r + d] = mv.x[d][i];
      oix[dim * r + d] = mv.ix[d][i];
    }
    for (int d = 0; d < 3; d++) op[3 * r + d] = mv.p[d][i];
    oq[r] = mv.q[i];
    if (otag) otag[r] = mv.tag ? mv.tag[i] : -1;
  }
}

// Movers meet the boundaries of this rank, axis by axis, lower then upper side
// (update_boundary_species -> process_phys_boundary, spec/os-spec-boundary.f03:1238-1465):
//   PBC_WRAP     this rank is its own periodic neighbour: ix +- nx (:1280-1290)
//   PBC_ABSORB   physical wall, particle removed (:1293-1300; the lower wall of a moving window)
//   PBC_SPECULAR physical wall, p_d -> -p_d, ix -> 1 | nx, x -> -x (:1302-1325; the upper wall of a
//                moving window behaves like this, :1276)
//   PBC_NONE     another rank's side: the particle was migrated before, out of range here is an error
// then the arrivals per destination tile are counted.
template <typename T>
__global__ void __launch_bounds__(256)
    route_kernel(MoverView<T> mv, int dim, Tiling tl, Grid g, PartBc bc, int *tile_arr, int *errflag, MoverView<T> guests) {
  const int n = min(*mv.count, mv.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    int ix[3] = {1, 1, 1};
    if (mv.ix[0][i] == INT_MIN) continue;  // sent to another rank
    bool gone = false, guest = false;
    for (int d = 0; d < dim && !gone; d++) {
      int v = mv.ix[d][i];
      const int side = (v < 1) ? 0 : (v > g.nx[d] ? 1 : -1);
      if (side >= 0) {
        const int type = bc.t[side][d];
        if (type == PBC_WRAP) {
          v += side ? -g.nx[d] : g.nx[d];
        } else if (type == PBC_ABSORB) {
          if (side == 0 && ((bc.defer_lower >> d) & 1u)) guest = true;  // (keeps its index below 1 on this axis)
          else gone = true;
        } else if (type == PBC_SPECULAR) {
          mv.p[d][i] = -mv.p[d][i];
          mv.x[d][i] = -mv.x[d][i];
          v = side ? g.nx[d] : 1;
        }
        mv.ix[d][i] = v;
      }
      ix[d] = v;
    }
    if (guest && !gone) {  // pushed once more from the guest list, then met by the wall like any leaver
      const int k = atomicAdd(guests.count, 1);
      if (k < guests.cap) {
        for (int d = 0; d < dim; d++) { guests.x[d][k] = mv.x[d][i]; guests.ix[d][k] = mv.ix[d][i]; }
        for (int d = 0; d < 3; d++) guests.p[d][k] = mv.p[d][i];
        guests.q[k] = mv.q[i];
        if (guests.tag) guests.tag[k] = mv.tag[i];
      } else {
        atomicOr(errflag, ERRF_MOVER_OVERFLOW);
      }
      gone = true;
    }
    if (gone) { mv.ix[0][i] = INT_MIN; continue; }
    int t, c;
    if (locate(ix, dim, tl, g.nx, &t, &c)) atomicAdd(&tile_arr[t], 1);
    else atomicOr(errflag, ERRF_BAD_CELL);
  }
}

// host-layout particles appended to the mover buffer (injection, e.g. the plasma a moving window
// uncovers: inject -> create particles, spec/os-species.f03:218-226)
template <typename T>
__global__ void __launch_bounds__(256)
    inject_kernel(AosParticles<T> in, int np, long long tag_base, int dim, MoverView<T> mv, int base) {
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < np; i += gridDim.x * blockDim.x) {
    const int m = base + i;
    for (int d = 0; d < dim; d++) { mv.x[d][m] = in.x[dim * i + d]; mv.ix[d][m] = in.ix[dim * i + d]; }
    for (int d = 0; d < 3; d++) mv.p[d][m] = in.p[3 * i + d];
    mv.q[m] = in.q[i];
    if (mv.tag) mv.tag[m] = in.tag ? in.tag[i] : tag_base + i;
    if (i == 0) *mv.count = base + np;
  }
}

__global__ void __launch_bounds__(256)
    capacity_kernel(const int *tile_off, const int *tile_cnt, const int *tile_arr, int ntile, int *errflag) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < ntile && tile_cnt[t] + tile_arr[t] > tile_off[t + 1] - tile_off[t]) atomicOr(errflag, ERRF_TILE_OVERFLOW);
}

template <typename T>
__global__ void __launch_bounds__(256)
    insert_kernel(MoverView<T> mv, int dim, Tiling tl, Grid g, const int *tile_off, int *tile_cnt,
                  SpeciesView<T> sp) {
  const int n = min(*mv.count, mv.cap);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (mv.ix[0][i] == INT_MIN) continue;
    int ix[3] = {1, 1, 1};
    for (int d = 0; d < dim; d++) ix[d] = mv.ix[d][i];
    int t, c;
    if (!locate(ix, dim, tl, g.nx, &t, &c)) continue;
    const long long s = (long long)tile_off[t] + atomicAdd(&tile_cnt[t], 1);
    for (int d = 0; d < dim; d++) sp.r(d, s) = mv.x[d][i];
    for (int d = 0; d < 3; d++) sp.r(dim + d, s) = mv.p[d][i];
    sp.r(dim + 3, s) = mv.q[i];
    sp.cell[s] = c;
    if (sp.tag) sp.tag[s] = mv.tag[i];
  }
}

// one CTA per tile: copy the live range into the new layout (counting sort by tile, out of place
// like rearrange_species, os-species.f03:512-702)
template <typename T>
__global__ void __launch_bounds__(256)
    relayout_kernel(SpeciesView<T> src, const int *old_off, const int *tile_cnt, SpeciesView<T> dst,
                    const int *new_off, int dim) {
  const int tile = blockIdx.x;
  const int n = tile_cnt[tile];
  const long long so = old_off[tile], dn = new_off[tile];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    for (int d = 0; d < dim; d++) dst.r(d, dn + i) = src.r(d, so + i);
    for (int d = 0; d < 3; d++) dst.r(dim + d, dn + i) = src.r(dim + d, so + i);
    dst.r(dim + 3, dn + i) = src.r(dim + 3, so + i);
    dst.cell[dn + i] = src.cell[so + i];
    if (dst.tag) dst.tag[dn + i] = src.tag[so + i];
  }
}

// ---- migration -------------------------------------------------------------------------------
// wire record (32-bit words): [tag lo, tag hi][reals: x(D) p(3) q][ix(D)][pad to even]
template <typename T> __host__ __device__ inline int rec_words(int dim) {
  const int w = 2 + (dim + 4) * (int)(sizeof(T) / 4) + dim;
  return (w + 1) & ~1;
}
template <typename T> int record_words(int dim) { return rec_words<T>(dim); }
template int record_words<float>(int);
template int record_words<double>(int);

// movers leaving through the lower / upper face of `axis` go to the send buffers
// (check_boundary + pack, os-spec-boundary.f03:538-623, os-spec-comm.f90:454-522).  The upper
// message is re-indexed by the sender (ix - nx), the lower one by the receiver (ix + its nx).
template <typename T>
__global__ void __launch_bounds__(256)
    classify_kernel(MoverView<T> mv, int dim, int axis, int nx_axis, int send_lo, int send_up, uint32_t *buf_lo,
                    uint32_t *buf_up, int *cnt2, int cap_rec, int *errflag) {
  const int n = min(*mv.count, mv.cap);
  const int rw = rec_words<T>(dim);
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    if (mv.ix[0][i] == INT_MIN) continue;
    const int v = mv.ix[axis][i];
    int side = -1;
    if (v < 1 && send_lo) side = 0;
    else if (v > nx_axis && send_up) side = 1;
    if (side < 0) continue;
    const int k = atomicAdd(&cnt2[side], 1);
    if (k >= cap_rec) { atomicOr(errflag, ERRF_SEND_OVERFLOW); continue; }
    uint32_t *r = (side ? buf_up : buf_lo) + (size_t)k * rw;
    const long long tag = mv.tag ? mv.tag[i] : 0;
    r[0] = (uint32_t)(tag & 0xffffffffll);
    r[1] = (uint32_t)((unsigned long long)tag >> 32);
    T *f = reinterpret_cast<T *>(r + 2);
    for (int d = 0; d < dim; d++) f[d] = mv.x[d][i];
    for (int d = 0; d < 3; d++) f[dim + d] = mv.p[d][i];
    f[dim + 3] = mv.q[i];
    int *ii = reinterpret_cast<int *>(f + dim + 4);
    for (int d = 0; d < dim; d++) ii[d] = mv.ix[d][i] - ((side == 1 && d == axis) ? nx_axis : 0);
    mv.ix[0][i] = INT_MIN;  // no longer ours
  }
}

template <typename T>
__global__ void __launch_bounds__(256)
    unpack_kernel(const uint32_t *buf, int nrec, int dim, int axis, int add_axis, MoverView<T> mv, int *errflag) {
  const int rw = rec_words<T>(dim);
  __shared__ int s_base;
  // reserve a contiguous range per block
  const int first = blockIdx.x * blockDim.x;
  const int nb = min(blockDim.x, nrec - first);
  if (nb <= 0) return;
  if (threadIdx.x == 0) {
    int b = atomicAdd(mv.count, nb);
    if (b + nb > mv.cap) { atomicOr(errflag, ERRF_MOVER_OVERFLOW); b = -1; }
    s_base = b;
  }
  __syncthreads();
  if (s_base < 0 || (int)threadIdx.x >= nb) return;
  const int i = first + threadIdx.x, k = s_base + threadIdx.x;
  const uint32_t *r = buf + (size_t)i * rw;
  if (mv.tag) mv.tag[k] = (long long)(((unsigned long long)r[1] << 32) | r[0]);
  const T *f = reinterpret_cast<const T *>(r + 2);
  for (int d = 0; d < dim; d++) mv.x[d][k] = f[d];
  for (int d = 0; d < 3; d++) mv.p[d][k] = f[dim + d];
  mv.q[k] = f[dim + 3];
  const int *ii = reinterpret_cast<const int *>(f + dim + 4);
  for (int d = 0; d < dim; d++) mv.ix[d][k] = ii[d] + ((d == axis) ? add_axis : 0);
}

// ---- host-side launchers ---------------------------------------------------------------------
#define LAUNCHED(n)            \
  *launches += (n);            \
  OSGPU_CUDA(cudaGetLastError()); \
  return OSGPU_OK

static int blocks_for(long long n) {
  long long b = (n + 255) / 256;
  if (b < 1) b = 1;
  if (b > GRID_STRIDE_BLOCKS * 4) b = GRID_STRIDE_BLOCKS * 4;
  return (int)b;
}

template <typename T>
int bin_count(const AosParticles<T> &in, long long np, const Grid &g, const Tiling &tl, const Shift3 &shift,
              int to_movers, int *cell_cnt, int *cell_start, int *tile_cnt, int *tile_nsorted, int *errflag,
              cudaStream_t st, long long *launches) {
  if (np > 0) count_kernel<T><<<blocks_for(np), 256, 0, st>>>(in, np, g.dim, tl, g, shift, to_movers, cell_cnt, errflag);
  cell_scan_kernel<<<tl.ntile, TILE_NCELL, 0, st>>>(cell_cnt, cell_start, tile_cnt, tile_nsorted);
  LAUNCHED(2);
}
template <typename T>
int bin_scatter(const AosParticles<T> &in, long long np, long long tag_base, const Grid &g, const Tiling &tl,
                const Shift3 &shift, int to_movers, const MoverView<T> &mv, int *errflag, const int *tile_off,
                const int *cell_start, int *cursor, const SpeciesView<T> &sp, cudaStream_t st, long long *launches) {
  if (np == 0) return OSGPU_OK;
  scatter_kernel<T><<<blocks_for(np), 256, 0, st>>>(in, np, tag_base, g.dim, tl, g, shift, to_movers, mv, errflag,
                                                    tile_off, cell_start, cursor, sp);
  LAUNCHED(1);
}
template <typename T>
int gather_download(const SpeciesView<T> &sp, const int *tile_off, const int *tile_cnt, const long long *out_off,
                    const Grid &g, const Tiling &tl, int *oix, T *ox, T *op, T *oq, long long *otag, int by_tag,
                    cudaStream_t st, long long *launches) {
  gather_kernel<T><<<tl.ntile, 256, 0, st>>>(sp, tile_off, tile_cnt, out_off, g.dim, tl, oix, ox, op, oq, otag,
                                             by_tag);
  LAUNCHED(1);
}
template <typename T>
int movers_download(const MoverView<T> &mv, long long out_base, const Grid &g, int *oix, T *ox, T *op, T *oq,
                    long long *otag, int by_tag, cudaStream_t st, long long *launches) {
  mover_gather_kernel<T><<<GRID_STRIDE_BLOCKS, 256, 0, st>>>(mv, out_base, g.dim, oix, ox, op, oq, otag, by_tag);
  LAUNCHED(1);
}
template <typename T>
int movers_inject(const AosParticles<T> &in, int np, long long tag_base, const Grid &g, const MoverView<T> &mv,
                  int base, cudaStream_t st, long long *launches) {
  if (np == 0) return OSGPU_OK;
  inject_kernel<T><<<blocks_for(np), 256, 0, st>>>(in, np, tag_base, g.dim, mv, base);
  LAUNCHED(1);
}
template <typename T>
int movers_route(const MoverView<T> &mv, const Grid &g, const Tiling &tl, const PartBc &bc, const int *tile_off,
                 const int *tile_cnt, int *tile_arr, int *errflag, const MoverView<T> &guests, cudaStream_t st,
                 long long *launches) {
  OSGPU_CUDA(cudaMemsetAsync(tile_arr, 0, sizeof(int) * tl.ntile, st));
  route_kernel<T><<<GRID_STRIDE_BLOCKS, 256, 0, st>>>(mv, g.dim, tl, g, bc, tile_arr, errflag, guests);
  capacity_kernel<<<(tl.ntile + 255) / 256, 256, 0, st>>>(tile_off, tile_cnt, tile_arr, tl.ntile, errflag);
  LAUNCHED(2);
}
template <typename T>
int movers_insert(const MoverView<T> &mv, const Grid &g, const Tiling &tl, const int *tile_off, int *tile_cnt,
                  const SpeciesView<T> &sp, cudaStream_t st, long long *launches) {
  insert_kernel<T><<<GRID_STRIDE_BLOCKS, 256, 0, st>>>(mv, g.dim, tl, g, tile_off, tile_cnt, sp);
  LAUNCHED(1);
}
template <typename T>
int relayout(const SpeciesView<T> &src, const int *old_off, const int *tile_cnt, const SpeciesView<T> &dst,
             const int *new_off, const Grid &g, const Tiling &tl, cudaStream_t st, long long *launches) {
  relayout_kernel<T><<<tl.ntile, 256, 0, st>>>(src, old_off, tile_cnt, dst, new_off, g.dim);
  LAUNCHED(1);
}
template <typename T>
int movers_classify(const MoverView<T> &mv, const Grid &g, int axis, int send_lo, int send_up, uint32_t *buf_lo,
                    uint32_t *buf_up, int *cnt2, int cap_rec, int *errflag, cudaStream_t st,
                    long long *launches) {
  classify_kernel<T><<<GRID_STRIDE_BLOCKS, 256, 0, st>>>(mv, g.dim, axis, g.nx[axis], send_lo, send_up, buf_lo,
                                                         buf_up, cnt2, cap_rec, errflag);
  LAUNCHED(1);
}
template <typename T>
int movers_unpack(const uint32_t *buf, int nrec, const Grid &g, int axis, int add_axis, const MoverView<T> &mv,
                  int *errflag, cudaStream_t st, long long *launches) {
  if (nrec <= 0) return OSGPU_OK;
  unpack_kernel<T><<<(nrec + 255) / 256, 256, 0, st>>>(buf, nrec, g.dim, axis, add_axis, mv, errflag);
  LAUNCHED(1);
}

// ---------------------------------------------------------------------------------------------
// charge density (a diagnostic: one CTA per tile, one thread per particle, global atomics)
// ---------------------------------------------------------------------------------------------
template <int DIM, int ORDER, typename T>
__global__ void __launch_bounds__(256)
    charge_kernel(SpeciesView<T> sp, const int *tile_off, const int *tile_cnt, Grid g, Tiling tl, T *rho) {
  constexpr int NP = Stencil<ORDER>::NP, LP = Stencil<ORDER>::LP;
  const int tile = blockIdx.x;
  const int n = tile_cnt[tile];
  const long long off = tile_off[tile];
  int t = tile, o0[3];
  o0[0] = (t % tl.nt[0]) * tl.T[0]; t /= tl.nt[0];
  o0[1] = (t % tl.nt[1]) * tl.T[1]; t /= tl.nt[1];
  o0[2] = t * tl.T[2];
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long s = off + i;
    const int cell = sp.cell[s];
    T w[DIM][NP];
    int b[3] = {0, 0, 0};
#pragma unroll
    for (int d = 0; d < DIM; d++) {
      spline<ORDER, T>(sp.r(d, s), w[d]);
      b[d] = o0[d] + cell_l(cell, d) + g.gl[d] + LP;
    }
    const T q = sp.r(DIM + 3, s);
#pragma unroll
    for (int k3 = 0; k3 < (DIM == 3 ? NP : 1); k3++)
#pragma unroll
      for (int k2 = 0; k2 < NP; k2++)
#pragma unroll
        for (int k1 = 0; k1 < NP; k1++) {
          T v = (q * w[0][k1]) * w[1][k2];
          if (DIM == 3) v = v * w[DIM - 1][k3];
          atomicAdd(&rho[g.at(b[0] + k1, b[1] + k2, b[2] + k3)], v);
        }
  }
}
template <typename T>
int deposit_charge(const SpeciesView<T> &sp, const int *tile_off, const int *tile_cnt, const Grid &g, const Tiling &tl,
                   int order, T *rho, cudaStream_t st, long long *launches) {
#define CASE(D, O) \
  if (g.dim == D && order == O) charge_kernel<D, O, T><<<tl.ntile, 256, 0, st>>>(sp, tile_off, tile_cnt, g, tl, rho);
  CASE(2, 1) CASE(2, 2) CASE(2, 3) CASE(2, 4) CASE(3, 1) CASE(3, 2) CASE(3, 3) CASE(3, 4)
#undef CASE
  LAUNCHED(1);
}
// ---------------------------------------------------------------------------------------------
// phasespaces (a diagnostic): one CTA per tile, one thread per particle, global atomics.  Arithmetic in
// the reference's precisions: axis values and weights in p_diag_prec (float), the cell index from the
// particle-precision product (x - xmin) * rdx (deposit_1D / _2D, os-spec-phasespace.f03:1633-1775).
// ---------------------------------------------------------------------------------------------
struct PhaseArgs {
  osgpu_phasespace ps;
  double rqm;
  float dx[3];
};
// get_position (spec/os-spec-define.f03:1098-1123): xmin + dx * real((ix + ixmin) + x, p_single)
template <typename T>
__device__ __forceinline__ float phase_position(const PhaseArgs &a, int c, const int *ix, const T *x) {
  return (float)a.ps.xmin[c] + a.dx[c] * (float)((T)(ix[c] + a.ps.ixmin[c]) + x[c]);
}
// get_l (os-spec-phasespace.f03:314-409): angular momentum component c about the centre of the global box;
// in 2D the terms of the absent x3 drop out, exactly as the reference's select case does
template <int DIM, typename T>
__device__ __forceinline__ float phase_l(const PhaseArgs &a, int c, const int *ix, const T *x, const T *p) {
  auto arm = [&](int cx, int cp) { return (phase_position<T>(a, cx, ix, x) - (float)a.ps.center[cx]) * (float)p[cp]; };
  if (c == 0) {
    float l = arm(1, 2);
    if (DIM == 3) l = l - arm(2, 1);
    return l;
  }
  if (c == 1) {
    float l = -arm(0, 2);
    if (DIM == 3) l = l + arm(2, 0);
    return l;
  }
  return arm(0, 1) - arm(1, 0);
}
// get_phasespace_axis_spec (:3196-3238)
template <int DIM, typename T>
__device__ __forceinline__ float phase_axis(const PhaseArgs &a, int k, const int *ix, const T *x, const T *p) {
  const int c = a.ps.xp_dim[k] - 1;
  switch (a.ps.x_or_p[k]) {
  case 1: return phase_position<T>(a, c, ix, x);
  case 2: return (float)p[c];
  case 3: return sqrtf((float)(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + T(1)));
  case 4: return log10f(sqrtf((float)(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] + T(1))));
  case 5: return phase_l<DIM, T>(a, c, ix, x, p);
  default: {  // log_ke, never below tiny(1.0_p_single)
    const T u2 = (p[0] * p[0] + p[1] * p[1]) + p[2] * p[2];
    return log10f(fmaxf((float)(u2 / (sqrt(T(1) + u2) + T(1))), 1.17549435e-38f));
  }
  }
}
// get_phasespace_charge (:2627-2737); rqm is a particle-precision number in the reference
template <typename T>
__device__ __forceinline__ float phase_quantity(const PhaseArgs &a, T q, const T *p) {
  const float u2 = (float)(p[0] * p[0] + p[1] * p[1] + p[2] * p[2]);
  const float gamma = sqrtf(u2 + 1.0f);
  const T rqm = (T)a.rqm;
  const int pt = a.ps.ps_type;
  float lq;
  if (pt == 0) lq = (float)q;
  else if (pt == 1) lq = (float)(rqm * q);
  else if (pt == 2) lq = (float)(rqm * q) * u2 / (gamma + 1.0f);
  else if (pt <= 5) lq = ((float)(rqm * q) * u2 / (gamma + 1.0f)) * (float)p[pt - 3] / gamma;
  else if (pt == 6) lq = fabsf((float)q);
  else lq = (float)(q * p[pt - 7]) / gamma;
  if (a.ps.use_ene_bin) {
    const float kin = u2 / (gamma + 1.0f);
    if (kin <= (float)a.ps.ene_bin[0] || kin > (float)a.ps.ene_bin[1]) lq = 0.0f;
  }
  return lq;
}
template <int DIM, typename T>
__global__ void __launch_bounds__(256)
    phasespace_kernel(SpeciesView<T> sp, const int *tile_off, const int *tile_cnt, Tiling tl, PhaseArgs a, float *buf) {
  const int tile = blockIdx.x;
  const int n = tile_cnt[tile];
  const long long off = tile_off[tile];
  int t = tile, o0[3];
  o0[0] = (t % tl.nt[0]) * tl.T[0]; t /= tl.nt[0];
  o0[1] = (t % tl.nt[1]) * tl.T[1]; t /= tl.nt[1];
  o0[2] = t * tl.T[2];
  const int nd = a.ps.ndims;
  T rdx[3] = {T(0), T(0), T(0)};
  for (int k = 0; k < nd; k++) rdx[k] = (T)a.ps.n[k] / ((T)a.ps.max[k] - (T)a.ps.min[k]);
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const long long s = off + i;
    const int cell = sp.cell[s];
    int ix[3] = {1, 1, 1};
    T x[3] = {T(0), T(0), T(0)}, p[3];
#pragma unroll
    for (int d = 0; d < DIM; d++) { ix[d] = o0[d] + cell_l(cell, d) + 1; x[d] = sp.r(d, s); }
#pragma unroll
    for (int d = 0; d < 3; d++) p[d] = sp.r(DIM + d, s);
    const float lq = phase_quantity<T>(a, sp.r(DIM + 3, s), p);
    // deposit_1D / _2D / _3D (:1633-1902): nearest point and the one above it along every axis, linear
    // weights multiplied in the reference's order ((w1 * w2) * w3) * q
    int i1[3] = {1, 1, 1};
    float eps[3] = {0.0f, 0.0f, 0.0f};
    for (int k = 0; k < nd; k++) {
      const T nx1 = ((T)phase_axis<DIM, T>(a, k, ix, x, p) - (T)a.ps.min[k]) * rdx[k];
      i1[k] = (int)lrint(round((double)nx1));  // nint: halves away from zero
      eps[k] = (float)(nx1 - (T)i1[k]) + 0.5f;
    }
    const int n1 = a.ps.n[0], n2 = nd > 1 ? a.ps.n[1] : 1, n3 = nd > 2 ? a.ps.n[2] : 1;
    for (int b3 = 0; b3 < (nd > 2 ? 2 : 1); b3++) {
      const int j3 = nd > 2 ? i1[2] + b3 : 1;
      if (j3 < 1 || j3 > n3) continue;
      for (int b2 = 0; b2 < (nd > 1 ? 2 : 1); b2++) {
        const int j2 = nd > 1 ? i1[1] + b2 : 1;
        if (j2 < 1 || j2 > n2) continue;
        for (int b1 = 0; b1 < 2; b1++) {
          const int j1 = i1[0] + b1;
          if (j1 < 1 || j1 > n1) continue;
          float w = b1 ? eps[0] : 1.0f - eps[0];
          if (nd > 1) w = w * (b2 ? eps[1] : 1.0f - eps[1]);
          if (nd > 2) w = w * (b3 ? eps[2] : 1.0f - eps[2]);
          atomicAdd(&buf[(j1 - 1) + (long long)n1 * ((j2 - 1) + (long long)n2 * (j3 - 1))], w * lq);
        }
      }
    }
  }
}
template <typename T>
int deposit_phasespace(const SpeciesView<T> &sp, const int *tile_off, const int *tile_cnt, const Grid &g, const Tiling &tl,
                       const osgpu_phasespace &ps, double rqm, const double *dx, float *buf, cudaStream_t st,
                       long long *launches) {
  PhaseArgs a;
  a.ps = ps;
  a.rqm = rqm;
  for (int d = 0; d < 3; d++) a.dx[d] = (d < g.dim) ? (float)dx[d] : 0.0f;
  if (g.dim == 2) phasespace_kernel<2, T><<<tl.ntile, 256, 0, st>>>(sp, tile_off, tile_cnt, tl, a, buf);
  else phasespace_kernel<3, T><<<tl.ntile, 256, 0, st>>>(sp, tile_off, tile_cnt, tl, a, buf);
  LAUNCHED(1);
}

template <typename T>
__global__ void __launch_bounds__(256) plane_out_kernel(const T *plane, Grid g, T *out, int hs0, int hs1, int hs2, int hg0, int hg1, int hg2) {
  const long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (r >= (long long)hs0 * hs1 * hs2) return;
  const int i0 = (int)(r % hs0), i1 = (int)((r / hs0) % hs1), i2 = (int)(r / ((long long)hs0 * hs1));
  out[r] = plane[g.at(i0 - hg0 + g.gl[0], i1 - hg1 + g.gl[1], i2 - hg2 + g.gl[2])];
}
template <typename T>
int plane_to_host_layout(const Grid &g, const T *plane, T *out, const int *hs, const int *hg, cudaStream_t st,
                         long long *launches) {
  const long long n = (long long)hs[0] * hs[1] * hs[2];
  plane_out_kernel<T><<<(unsigned)((n + 255) / 256), 256, 0, st>>>(plane, g, out, hs[0], hs[1], hs[2], hg[0], hg[1], hg[2]);
  LAUNCHED(1);
}

#define INST(T)                                                                                               \
  template int bin_count<T>(const AosParticles<T> &, long long, const Grid &, const Tiling &, const Shift3 &, \
                            int, int *, int *, int *, int *, int *, cudaStream_t, long long *);               \
  template int bin_scatter<T>(const AosParticles<T> &, long long, long long, const Grid &, const Tiling &,     \
                              const Shift3 &, int, const MoverView<T> &, int *, const int *, const int *,     \
                              int *, const SpeciesView<T> &, cudaStream_t, long long *);                      \
  template int movers_inject<T>(const AosParticles<T> &, int, long long, const Grid &, const MoverView<T> &,   \
                                int, cudaStream_t, long long *);                                              \
  template int gather_download<T>(const SpeciesView<T> &, const int *, const int *, const long long *,         \
                                  const Grid &, const Tiling &, int *, T *, T *, T *, long long *, int,        \
                                  cudaStream_t, long long *);                                                 \
  template int movers_download<T>(const MoverView<T> &, long long, const Grid &, int *, T *, T *, T *,         \
                                  long long *, int, cudaStream_t, long long *);                               \
  template int movers_route<T>(const MoverView<T> &, const Grid &, const Tiling &, const PartBc &, const int *, \
                               const int *, int *, int *, const MoverView<T> &, cudaStream_t, long long *);   \
  template int movers_insert<T>(const MoverView<T> &, const Grid &, const Tiling &, const int *, int *,        \
                                const SpeciesView<T> &, cudaStream_t, long long *);                           \
  template int relayout<T>(const SpeciesView<T> &, const int *, const int *, const SpeciesView<T> &,           \
                           const int *, const Grid &, const Tiling &, cudaStream_t, long long *);             \
  template int movers_classify<T>(const MoverView<T> &, const Grid &, int, int, int, uint32_t *, uint32_t *,   \
                                  int *, int, int *, cudaStream_t, long long *);                              \
  template int movers_unpack<T>(const uint32_t *, int, const Grid &, int, int, const MoverView<T> &, int *,    \
                                cudaStream_t, long long *);                                                   \
  template int deposit_charge<T>(const SpeciesView<T> &, const int *, const int *, const Grid &, const Tiling &, \
                                 int, T *, cudaStream_t, long long *);                                        \
  template int plane_to_host_layout<T>(const Grid &, const T *, T *, const int *, const int *, cudaStream_t,   \
                                       long long *);                                                          \
  template int deposit_phasespace<T>(const SpeciesView<T> &, const int *, const int *, const Grid &, const Tiling &, \
                                     const osgpu_phasespace &, double, const double *, float *, cudaStream_t, long long *);
INST(float)
INST(double)

}  // namespace osgpu
