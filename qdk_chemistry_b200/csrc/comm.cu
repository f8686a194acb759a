// NCCL plumbing for the row-sharded solver (one process per GPU).
//
// Replaces the MPI calls of the reference's distributed path: MPI_Allgatherv of vector
// blocks (solvers/davidson.hpp:129-167, asci/iteration.hpp:206-215), MPI_Allreduce of the
// K inner products (davidson.hpp:411,421) and the halo exchange of pgespmv
// (sparsexx/spblas/pspmbv.hpp:316-405), which becomes an all-gather of the trial vector
// over NVLink/NVSwitch. libnccl is resolved at run time (the copy torch already loaded),
// so the single-GPU path has no NCCL dependency at all.
#include <dlfcn.h>
#include <nccl.h>

#include "common.cuh"

namespace b2ci {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t,
                            cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi& api() {
  static NcclApi a;
  if (a.handle) return a;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  for (auto nm : names) {
    a.handle = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (a.handle) break;
  }
  if (!a.handle) throw Error(std::string("cannot load libnccl: ") + dlerror());
#define B2_SYM(field, sym)                                              \
  a.field = reinterpret_cast<decltype(a.field)>(dlsym(a.handle, sym));  \
  if (!a.field) throw Error(std::string("libnccl lacks symbol ") + sym);
  B2_SYM(GetUniqueId, "ncclGetUniqueId")
  B2_SYM(CommInitRank, "ncclCommInitRank")
  B2_SYM(CommDestroy, "ncclCommDestroy")
  B2_SYM(AllReduce, "ncclAllReduce")
  B2_SYM(AllGather, "ncclAllGather")
  B2_SYM(Broadcast, "ncclBroadcast")
  B2_SYM(GroupStart, "ncclGroupStart")
  B2_SYM(GroupEnd, "ncclGroupEnd")
  B2_SYM(GetErrorString, "ncclGetErrorString")
#undef B2_SYM
  return a;
}

#define B2_NCCL(expr)                                                                     \
  do {                                                                                    \
    ncclResult_t _r = (expr);                                                             \
    if (_r != ncclSuccess)                                                                \
      throw Error(std::string(#expr) + ": " + api().GetErrorString(_r));                  \
  } while (0)

void comm_unique_id(void* id128) {
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  ncclUniqueId id;
  B2_NCCL(api().GetUniqueId(&id));
  memcpy(id128, &id, 128);
}

void comm_init(b2ci_ctx* ctx, const void* id128, int rank, int nranks) {
  if (nranks < 1 || rank < 0 || rank >= nranks) throw Error("b2ci_comm_init: bad rank/nranks");
  if (ctx->nccl_comm) throw Error("b2ci_comm_init: communicator already initialised");
  ctx->rank = rank;
  ctx->nranks = nranks;
  if (nranks == 1) return;
  ncclUniqueId id;
  memcpy(&id, id128, 128);
  ncclComm_t comm;
  B2_CUDA(cudaSetDevice(ctx->device));
  B2_NCCL(api().CommInitRank(&comm, nranks, id, rank));
  ctx->nccl_comm = comm;
}

void comm_destroy(b2ci_ctx* ctx) {
  if (ctx->nccl_comm) {
    api().CommDestroy((ncclComm_t)ctx->nccl_comm);
    ctx->nccl_comm = nullptr;
  }
}

// all-gather of unequal contiguous row blocks: one broadcast per owner inside a group
void comm_allgather_rows(b2ci_ctx* ctx, const double* local, double* full,
                         const std::vector<int64_t>& off) {
  if (ctx->nranks == 1) {
    if (local != full)
      B2_CUDA(cudaMemcpyAsync(full, local, size_t(off.empty() ? 0 : off[1] - off[0]) * 8,
                              cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  ncclComm_t comm = (ncclComm_t)ctx->nccl_comm;
  bool equal = true;
  const int64_t c0 = off[1] - off[0];
  for (int r = 0; r < ctx->nranks; ++r) equal = equal && (off[r + 1] - off[r] == c0);
  if (equal) {
    B2_NCCL(api().AllGather(local, full, size_t(c0), ncclDouble, comm, ctx->stream));
    return;
  }
  B2_NCCL(api().GroupStart());
  for (int r = 0; r < ctx->nranks; ++r) {
    const size_t cnt = size_t(off[r + 1] - off[r]);
    if (!cnt) continue;
    B2_NCCL(api().Broadcast(r == ctx->rank ? (const void*)local : (const void*)(full + off[r]),
                            full + off[r], cnt, ncclDouble, r, comm, ctx->stream));
  }
  B2_NCCL(api().GroupEnd());
}

void comm_allreduce_sum(b2ci_ctx* ctx, double* dev_buf, int64_t n) {
  if (ctx->nranks == 1 || n == 0) return;
  B2_NCCL(api().AllReduce(dev_buf, dev_buf, size_t(n), ncclDouble, ncclSum,
                          (ncclComm_t)ctx->nccl_comm, ctx->stream));
}

void comm_allgather_i64_host(b2ci_ctx* ctx, int64_t local, std::vector<int64_t>& all) {
  all.assign(ctx->nranks, 0);
  if (ctx->nranks == 1) { all[0] = local; return; }
  DevBuf<int64_t> s(1), r(ctx->nranks);
  B2_CUDA(cudaMemcpyAsync(s, &local, 8, cudaMemcpyHostToDevice, ctx->stream));
  B2_NCCL(api().AllGather(s, r, 1, ncclInt64, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(all.data(), r, size_t(ctx->nranks) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
}

void comm_allgather_bytes(b2ci_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
  if (ctx->nranks == 1) {
    if (send != recv) B2_CUDA(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  B2_NCCL(api().AllGather(send, recv, bytes_per_rank, ncclChar, (ncclComm_t)ctx->nccl_comm, ctx->stream));
}

void comm_allreduce_sum_i64_host(b2ci_ctx* ctx, int64_t* vals, int n) {
  if (ctx->nranks == 1 || n == 0) return;
  DevBuf<int64_t> d(n);
  B2_CUDA(cudaMemcpyAsync(d, vals, size_t(n) * 8, cudaMemcpyHostToDevice, ctx->stream));
  B2_NCCL(api().AllReduce(d, d, size_t(n), ncclInt64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
  B2_CUDA(cudaMemcpyAsync(vals, d, size_t(n) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
}

}  // namespace b2ci
