// Multi-GPU exchange of the loss path: ONE NCCL all-reduce of the reduced loss scalar (north_star; the batch
// shards by utterance, forward_backward.cpp:38-52 has no cross-utterance term, so nothing else is exchanged).
//
// The communicator lives in the C library so that a host language without torch.distributed can shard a
// batch, and so that the collective is enqueued from the same call sequence as the kernels (a c10d
// all_reduce of one scalar costs ~40 us of host time per step, more than the rest of the sharded step's
// host work).  libnccl is not linked: it is resolved at run time (the process usually has it loaded already).
#include <dlfcn.h>

#include <cstdlib>
#include <cstring>
#include <new>

#include "common.cuh"

namespace e2e {
namespace {

struct NcclId { char internal[128]; };
typedef void* NcclComm;
typedef int (*fn_get_id)(NcclId*);
typedef int (*fn_init_rank)(NcclComm*, int, NcclId, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, NcclComm, cudaStream_t);
typedef int (*fn_destroy)(NcclComm);
typedef const char* (*fn_err)(int);

struct NcclApi {
  void* so = nullptr;
  fn_get_id get_id = nullptr;
  fn_init_rank init_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_destroy destroy = nullptr;
  fn_err err = nullptr;
};

NcclApi* nccl_api() {
  static NcclApi api;
  static bool tried = false;
  if (!tried) {
    tried = true;
    const char* names[] = {getenv("E2E_CTC_NCCL_LIB"), "libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      if (!n || !*n) continue;
      void* h = dlopen(n, RTLD_NOW | RTLD_NOLOAD);     // already in the process (torch loads it)?
      if (!h) h = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (h) { api.so = h; break; }
    }
    if (api.so) {
      api.get_id = (fn_get_id)dlsym(api.so, "ncclGetUniqueId");
      api.init_rank = (fn_init_rank)dlsym(api.so, "ncclCommInitRank");
      api.all_reduce = (fn_all_reduce)dlsym(api.so, "ncclAllReduce");
      api.destroy = (fn_destroy)dlsym(api.so, "ncclCommDestroy");
      api.err = (fn_err)dlsym(api.so, "ncclGetErrorString");
      if (!api.get_id || !api.init_rank || !api.all_reduce || !api.destroy) api.so = nullptr;
    }
  }
  return api.so ? &api : nullptr;
}

int nccl_fail(const NcclApi* a, const char* what, int rc) {
  set_error("%s failed: %s", what, (a && a->err) ? a->err(rc) : "NCCL error");
  return E2E_ERR_CUDA;
}

}  // namespace
}  // namespace e2e

struct e2e_ctc_comm {
  e2e::NcclComm comm = nullptr;
  int nranks = 0, rank = 0;
};

using namespace e2e;

extern "C" {

int e2e_ctc_comm_unique_id(void* out_id128) {
  NcclApi* a = nccl_api();
  if (!a) { set_error("libnccl not found (set E2E_CTC_NCCL_LIB)"); return E2E_ERR_UNSUPPORTED; }
  if (!out_id128) { set_error("null id buffer"); return E2E_ERR_INVALID_ARGUMENT; }
  const int rc = a->get_id(reinterpret_cast<NcclId*>(out_id128));
  return rc == 0 ? E2E_OK : nccl_fail(a, "ncclGetUniqueId", rc);
}

int e2e_ctc_comm_create(const void* id128, int32_t nranks, int32_t rank, e2e_ctc_comm** out) {
  NcclApi* a = nccl_api();
  if (!a) { set_error("libnccl not found (set E2E_CTC_NCCL_LIB)"); return E2E_ERR_UNSUPPORTED; }
  if (!id128 || !out || nranks < 1 || rank < 0 || rank >= nranks) { set_error("bad communicator arguments"); return E2E_ERR_INVALID_ARGUMENT; }
  *out = nullptr;
  e2e_ctc_comm* c = new (std::nothrow) e2e_ctc_comm();
  if (!c) { set_error("out of host memory"); return E2E_ERR_CUDA; }
  NcclId id;
  memcpy(&id, id128, sizeof(id));
  const int rc = a->init_rank(&c->comm, nranks, id, rank);      // collective over the ranks, current CUDA device
  if (rc != 0) { delete c; return nccl_fail(a, "ncclCommInitRank", rc); }
  c->nranks = nranks; c->rank = rank;
  *out = c;
  return E2E_OK;
}

void e2e_ctc_comm_destroy(e2e_ctc_comm* c) {
  if (!c) return;
  NcclApi* a = nccl_api();
  if (a && c->comm) a->destroy(c->comm);
  delete c;
}

int e2e_ctc_comm_allreduce_sum(e2e_ctc_comm* c, void* buf, int64_t count, int32_t dtype, void* cuda_stream) {
  NcclApi* a = nccl_api();
  if (!a || !c || !c->comm) { set_error("no communicator"); return E2E_ERR_INVALID_ARGUMENT; }
  if (!buf || count < 1) { set_error("bad all-reduce buffer"); return E2E_ERR_INVALID_ARGUMENT; }
  int nt;
  switch (dtype) {       // ncclDataType_t
    case E2E_F32: nt = 7; break;
    case E2E_F64: nt = 8; break;
    case E2E_F16: nt = 6; break;
    case E2E_BF16: nt = 9; break;
    default: set_error("bad dtype"); return E2E_ERR_INVALID_ARGUMENT;
  }
  const int rc = a->all_reduce(buf, buf, (size_t)count, nt, /*ncclSum*/ 0, c->comm, reinterpret_cast<cudaStream_t>(cuda_stream));
  return rc == 0 ? E2E_OK : nccl_fail(a, "ncclAllReduce", rc);
}

}  // extern "C"
