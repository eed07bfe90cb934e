// comm.cuh -- NCCL point-to-point plumbing for the domain-decomposed exchanges.
//
// The reference moves guard cells and migrating particles with nearest-neighbour
// mpi_isend/irecv, one dimension at a time (source/vdf/os-vdf-comm.f90:665-722,
// source/spec/os-spec-boundary.f03:1134-1225).  Here the same scheme runs over NCCL
// (ncclSend/ncclRecv grouped per dimension) on the simulation's stream, so the exchanges are
// ordered with the kernels without host synchronisation.  libnccl is loaded with dlopen so a
// process that already carries NCCL (torch) shares the one copy.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

namespace osgpu {

struct Nccl {
  void *lib = nullptr;
  void *comm = nullptr;
  int rank = 0, nranks = 1;
  int load();                                   // dlopen + resolve; OSGPU_ERR_NCCL on failure
  static int unique_id(void *id128);
  int init(const void *id128, int nranks, int rank);
  // a second communicator over the same ranks (ncclCommSplit), for exchanges that run on another stream at the
  // same time as this one's; collective.  out->comm stays null when the loaded libnccl has no ncclCommSplit.
  int clone(Nccl *out);
  void destroy();
  // one neighbour exchange: send `ns0` bytes to peer `to0` and `ns1` bytes to `to1` while
  // receiving from `from0` / `from1`; a peer of -1 skips that leg.  Sends are posted before
  // receives and in the opposite order (to0, to1 | from1-side first) so a two-rank periodic ring
  // pairs the messages correctly (os-vdf-comm.f90:703-707).
  int exchange(const void *s0, size_t ns0, int to0, const void *s1, size_t ns1, int to1, void *r0,
               size_t nr0, int from0, void *r1, size_t nr1, int from1, cudaStream_t st);
};

}  // namespace osgpu
