// Native NCCL binding of the library's three collectives (include/dmk_grid.h: dmk_nccl_*).
//
// The reference exchanges through MPI: FFTW-MPI's transposed plans move the half spectrum between the z- and y-decomposed
// layouts (/root/reference/src/grid/fft.cpp:48-58), and MPI_Allreduce sums the homogenised stress, the divergence RMS and the
// tangent (src/grid/spectral.cpp:160-165, mech_base.cpp:395-402).  Here the same three operations are NCCL calls issued by the
// library itself on the context's stream -- ncclAllReduce(sum | max) of a few doubles and an all-to-all of ncclSend/ncclRecv
// pairs inside one group -- so that a host without Python (the reference's C++ driver) gets the multi-GPU path by handing
// over a ncclUniqueId it distributed with its own MPI_Bcast.  libnccl is resolved at run time (dlopen: the copy already in the
// process, e.g. torch's, else DMK_NCCL_LIB, else libnccl.so.2), so the library keeps loading on single-GPU hosts without NCCL.
// The callback table dmk_comm (bound to torch.distributed by solver_cpp_b200/parallel.py) remains the alternative.
#include <dlfcn.h>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <cuda_runtime.h>
#include "ctx.h"
#include "dmk_common.cuh"

namespace {
// the stable subset of nccl.h (2.x ABI): opaque communicator, 128-byte id, enums by value
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
enum { NCCL_SUCCESS = 0, NCCL_UINT8 = 1, NCCL_FLOAT64 = 8, NCCL_SUM = 0, NCCL_MAX = 2 };
struct Api {
  void* h = nullptr;
  int (*GetUniqueId)(ncclUniqueId*) = nullptr;
  int (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  int (*CommDestroy)(ncclComm_t) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  int (*GroupStart)() = nullptr;
  int (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  int (*GetVersion)(int*) = nullptr;
};
Api g_api;

bool load_api() {
  if (g_api.h) return true;
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);          // the copy the process already uses (torch's)
  if (!h) { const char* e = getenv("DMK_NCCL_LIB"); if (e) h = dlopen(e, RTLD_NOW | RTLD_GLOBAL); }
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) { dmk_set_error("libnccl.so.2 not found (set DMK_NCCL_LIB)"); return false; }
#define SYM(name) *(void**)(&g_api.name) = dlsym(h, "nccl" #name); if (!g_api.name) { dmk_set_error("libnccl: symbol nccl" #name " missing"); return false; }
  SYM(GetUniqueId) SYM(CommInitRank) SYM(CommDestroy) SYM(AllReduce) SYM(Send) SYM(Recv) SYM(GroupStart) SYM(GroupEnd) SYM(GetErrorString) SYM(GetVersion)
#undef SYM
  g_api.h = h;
  return true;
}

struct NcclBinding {
  ncclComm_t comm;
  dmk_ctx* ctx;
};

int fail(int rc, const char* what) {
  static char msg[256];
  snprintf(msg, sizeof msg, "%s: %s", what, g_api.GetErrorString ? g_api.GetErrorString(rc) : "nccl error");
  dmk_set_error(msg);
  return 1;
}

// dmk_comm callbacks (user = NcclBinding*), all on the context's stream
int cb_allreduce(void* user, double* d_buf, int n, int op) {
  NcclBinding* b = (NcclBinding*)user;
  const int rc = g_api.AllReduce(d_buf, d_buf, (size_t)n, NCCL_FLOAT64, op, b->comm, b->ctx->stream);
  return rc == NCCL_SUCCESS ? 0 : fail(rc, "ncclAllReduce");
}
int cb_sum(void* user, double* d_buf, int n) { return cb_allreduce(user, d_buf, n, NCCL_SUM); }
int cb_max(void* user, double* d_buf, int n) { return cb_allreduce(user, d_buf, n, NCCL_MAX); }
int cb_alltoall(void* user, const void* d_send, void* d_recv, size_t bytes_per_peer) {
  NcclBinding* b = (NcclBinding*)user;
  const int P = b->ctx->P;
  int rc = g_api.GroupStart();
  if (rc != NCCL_SUCCESS) return fail(rc, "ncclGroupStart");
  for (int p = 0; p < P && rc == NCCL_SUCCESS; p++) {
    rc = g_api.Send((const char*)d_send + (size_t)p * bytes_per_peer, bytes_per_peer, NCCL_UINT8, p, b->comm, b->ctx->stream);
    if (rc == NCCL_SUCCESS) rc = g_api.Recv((char*)d_recv + (size_t)p * bytes_per_peer, bytes_per_peer, NCCL_UINT8, p, b->comm, b->ctx->stream);
  }
  const int rc2 = g_api.GroupEnd();
  if (rc != NCCL_SUCCESS) return fail(rc, "ncclSend/ncclRecv");
  return rc2 == NCCL_SUCCESS ? 0 : fail(rc2, "ncclGroupEnd");
}
}  // namespace

extern "C" int dmk_nccl_unique_id(unsigned char id[128]) {
  if (!load_api()) return DMK_ERR_STATE;
  ncclUniqueId u;
  const int rc = g_api.GetUniqueId(&u);
  if (rc != NCCL_SUCCESS) { fail(rc, "ncclGetUniqueId"); return DMK_ERR_STATE; }
  memcpy(id, u.internal, 128);
  return DMK_OK;
}

// collective over all ranks of the context's world (every rank calls it with the same id); replaces the callback table
extern "C" int dmk_nccl_init(dmk_ctx* c, const unsigned char id[128]) {
  if (!c || !id) { dmk_set_error("dmk_nccl_init: null argument"); return DMK_ERR_ARG; }
  if (!load_api()) return DMK_ERR_STATE;
  if (c->nccl) { dmk_set_error("dmk_nccl_init: already initialised"); return DMK_ERR_STATE; }
  ncclUniqueId u;
  memcpy(u.internal, id, 128);
  NcclBinding* b = new NcclBinding{nullptr, c};
  const int rc = g_api.CommInitRank(&b->comm, c->P, u, c->rank);
  if (rc != NCCL_SUCCESS) { delete b; fail(rc, "ncclCommInitRank"); return DMK_ERR_STATE; }
  c->nccl = b;
  dmk_comm t;
  t.user = b; t.alltoall = cb_alltoall; t.allreduce_sum = cb_sum; t.allreduce_max = cb_max;
  c->comm = t; c->has_comm = true;
  return DMK_OK;
}

extern "C" int dmk_nccl_version(int* version) {
  if (!load_api()) return DMK_ERR_STATE;
  return g_api.GetVersion(version) == NCCL_SUCCESS ? DMK_OK : DMK_ERR_STATE;
}

void dmk_nccl_teardown(dmk_ctx* c) {
  if (!c->nccl) return;
  NcclBinding* b = (NcclBinding*)c->nccl;
  if (c->has_comm && c->comm.user == b) c->has_comm = false;
  if (g_api.CommDestroy) g_api.CommDestroy(b->comm);
  delete b;
  c->nccl = nullptr;
}
