// comm.cu -- see comm.cuh
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>

#include "comm.cuh"
#include "common.cuh"

namespace osgpu {

namespace {
struct Api {
  ncclResult_t (*GetUniqueId)(ncclUniqueId *);
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int);
  ncclResult_t (*CommDestroy)(ncclComm_t);
  ncclResult_t (*Send)(const void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*Recv)(void *, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
  ncclResult_t (*GroupStart)();
  ncclResult_t (*GroupEnd)();
  const char *(*GetErrorString)(ncclResult_t);
  ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t *, void *);  // optional (NCCL >= 2.18)
  void *lib = nullptr;
} api;

int load_api() {
  if (api.lib) return OSGPU_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *h = nullptr;
  for (const char *n : names) {
    h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (h) break;
  }
  if (!h) {
    set_error("cannot load libnccl: %s", dlerror());
    return OSGPU_ERR_NCCL;
  }
#define SYM(field, name)                                   \
  *(void **)(&api.field) = dlsym(h, name);                 \
  if (!api.field) {                                        \
    set_error("libnccl lacks %s", name);                   \
    return OSGPU_ERR_NCCL;                                 \
  }
  SYM(GetUniqueId, "ncclGetUniqueId")
  SYM(CommInitRank, "ncclCommInitRank")
  SYM(CommDestroy, "ncclCommDestroy")
  SYM(Send, "ncclSend")
  SYM(Recv, "ncclRecv")
  SYM(GroupStart, "ncclGroupStart")
  SYM(GroupEnd, "ncclGroupEnd")
  SYM(GetErrorString, "ncclGetErrorString")
#undef SYM
  *(void **)(&api.CommSplit) = dlsym(h, "ncclCommSplit");
  api.lib = h;
  return OSGPU_OK;
}
}  // namespace

#define NCCL_TRY(call)                                                                  \
  do {                                                                                  \
    ncclResult_t r__ = (call);                                                          \
    if (r__ != ncclSuccess) {                                                           \
      set_error("NCCL error at %s:%d: %s", __FILE__, __LINE__, api.GetErrorString(r__)); \
      return OSGPU_ERR_NCCL;                                                            \
    }                                                                                   \
  } while (0)

int Nccl::load() { return load_api(); }

int Nccl::unique_id(void *id128) {
  OSGPU_TRY(load_api());
  static_assert(sizeof(ncclUniqueId) == OSGPU_NCCL_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NCCL_TRY(api.GetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return OSGPU_OK;
}

int Nccl::init(const void *id128, int nranks_, int rank_) {
  OSGPU_TRY(load_api());
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  ncclComm_t c;
  NCCL_TRY(api.CommInitRank(&c, nranks_, id, rank_));
  comm = c;
  rank = rank_;
  nranks = nranks_;
  return OSGPU_OK;
}

int Nccl::clone(Nccl *out) {
  *out = Nccl();
  out->lib = lib;
  out->rank = rank;
  out->nranks = nranks;
  if (!comm || !api.CommSplit) return OSGPU_OK;
  ncclComm_t c = nullptr;
  NCCL_TRY(api.CommSplit((ncclComm_t)comm, 0, rank, &c, nullptr));
  out->comm = c;
  return OSGPU_OK;
}

void Nccl::destroy() {
  if (comm) api.CommDestroy((ncclComm_t)comm);
  comm = nullptr;
}

int Nccl::exchange(const void *s0, size_t ns0, int to0, const void *s1, size_t ns1, int to1, void *r0,
                   size_t nr0, int from0, void *r1, size_t nr1, int from1, cudaStream_t st) {
  ncclComm_t c = (ncclComm_t)comm;
  NCCL_TRY(api.GroupStart());
  if (to0 >= 0 && ns0) NCCL_TRY(api.Send(s0, ns0, ncclChar, to0, c, st));
  if (to1 >= 0 && ns1) NCCL_TRY(api.Send(s1, ns1, ncclChar, to1, c, st));
  if (from0 >= 0 && nr0) NCCL_TRY(api.Recv(r0, nr0, ncclChar, from0, c, st));
  if (from1 >= 0 && nr1) NCCL_TRY(api.Recv(r1, nr1, ncclChar, from1, c, st));
  NCCL_TRY(api.GroupEnd());
  return OSGPU_OK;
}

}  // namespace osgpu
