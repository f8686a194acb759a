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

// ---------------------------------------------------------------------------------------------
// Peer-to-peer exchange of the trial vector (the sigma step's only communication).
//
// Every rank owns two exchange buffers of N doubles (plain cudaMalloc, exported with CUDA IPC)
// and a flag word per peer. One sigma = epoch e:
//   k_push : copies this rank's block into buffer e&1 of EVERY rank (coalesced stores over
//            NVLink / NVSwitch), fences, and the last CTA publishes flag[rank] = e on every peer
//   k_wait : one warp spins until all peers' flags reached e
//   k_spmv : unchanged, reads the local buffer
// No NCCL launch, no staging. Two buffers make the write-after-read hazard impossible: a peer can
// only push epoch e+2 (same buffer as e) after it saw my flag e+1, which I publish after my SpMV
// of epoch e on the same stream. If IPC or peer access is unavailable on any rank, all ranks
// fall back to ncclAllGather together.
namespace {
struct P2P {
  bool tried = false, ok = false;
  size_t cap = 0;  // doubles per buffer
  double* xbuf[2] = {nullptr, nullptr};
  unsigned long long* flags = nullptr;  // [nranks], written by the peers
  unsigned int* done = nullptr;         // CTA counter of k_push
  double** d_peer_x[2] = {nullptr, nullptr};
  unsigned long long** d_peer_flag = nullptr;
  std::vector<void*> opened;
  unsigned long long epoch = 0;
  double* fallback = nullptr;  // library-owned gather buffer of the NCCL path
  size_t fallback_cap = 0;
  unsigned int* err_host = nullptr;  // mapped pinned word: k_wait gave up on peer (value - 1)
  unsigned int* err_dev = nullptr;
  unsigned long long timeout_ns = 600ull * 1000000000ull;
};
struct IpcPack { cudaIpcMemHandle_t h[3]; };

constexpr int PUSH_THREADS = 256;
__global__ void __launch_bounds__(PUSH_THREADS)
k_push(const double* __restrict__ local, int64_t nloc, int64_t row0, double* const* __restrict__ peer_x,
       unsigned long long* const* __restrict__ peer_flag, int nranks, int rank, unsigned long long epoch,
       unsigned int* __restrict__ done) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < nloc; i += stride) {
    const double v = local[i];
    for (int p = 0; p < nranks; ++p) peer_x[p][row0 + i] = v;
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int prev = atomicAdd(done, 1u);
    if (prev == gridDim.x - 1) {  // every CTA's stores are fenced: publish
      *done = 0;
      __threadfence_system();
      for (int p = 0; p < nranks; ++p)
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(peer_flag[p] + rank), "l"(epoch) : "memory");
    }
  }
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A peer may legitimately be late (an imbalanced H build ahead of the first sigma), so the wait is
// long (B2CI_P2P_TIMEOUT_S, default 600 s) and a timeout does not kill the context: it raises a word
// in mapped host memory, which the host turns into an error at its next exchange / synchronisation.
__global__ void k_wait(const unsigned long long* __restrict__ flags, int nranks, unsigned long long epoch,
                       unsigned long long timeout_ns, unsigned int* __restrict__ err) {
  const int r = threadIdx.x;
  if (r >= nranks) return;
  const unsigned long long t0 = global_ns();
  for (;;) {
    unsigned long long f;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(f) : "l"(flags + r) : "memory");
    if (f >= epoch) break;
    if (global_ns() - t0 > timeout_ns) {
      *err = 1u + unsigned(r);
      __threadfence_system();
      break;
    }
    __nanosleep(128);
  }
}

P2P* p2p_state(b2ci_ctx* ctx) {
  if (!ctx->p2p) ctx->p2p = new P2P;
  return static_cast<P2P*>(ctx->p2p);
}
// collective when buffers were exported: a peer may still have them open (cudaFree of an exported
// allocation before the importer's cudaIpcCloseMemHandle is undefined), so every rank closes what it
// imported, all ranks meet, and only then are the exported buffers freed
void p2p_release(b2ci_ctx* ctx, P2P* s, bool collective) {
  cudaStreamSynchronize(ctx->stream);
  for (void* q : s->opened) cudaIpcCloseMemHandle(q);
  s->opened.clear();
  if (collective && ctx->nccl_comm && !getenv("B2CI_P2P_NO_RELEASE_BARRIER")) {
    int64_t one = 1;
    try { comm_allreduce_sum_i64_host(ctx, &one, 1); } catch (...) {}
  }
  for (int b = 0; b < 2; ++b) {
    if (s->xbuf[b]) cudaFree(s->xbuf[b]);
    if (s->d_peer_x[b]) cudaFree(s->d_peer_x[b]);
    s->xbuf[b] = nullptr;
    s->d_peer_x[b] = nullptr;
  }
  if (s->flags) cudaFree(s->flags);
  if (s->done) cudaFree(s->done);
  if (s->d_peer_flag) cudaFree(s->d_peer_flag);
  s->flags = nullptr; s->done = nullptr; s->d_peer_flag = nullptr;
  s->cap = 0;
  s->ok = false;
  cudaGetLastError();
}
// collective: (re)create the exchange buffers for vectors of n doubles
void p2p_setup(b2ci_ctx* ctx, P2P* s, size_t n) {
  const int nr = ctx->nranks, me = ctx->rank;
  cudaStream_t st = ctx->stream;
  p2p_release(ctx, s, s->ok);  // s->ok is all-or-nothing, so every rank takes the same branch
  s->tried = true;
  s->epoch = 0;
  int64_t good = 1;
  const size_t cap = (n + 1023) & ~size_t(1023);
  IpcPack mine;
  memset(&mine, 0, sizeof(mine));
  if (getenv("B2CI_NO_P2P")) good = 0;
  if (good) {
    bool a = cudaMalloc(&s->xbuf[0], cap * 8) == cudaSuccess && cudaMalloc(&s->xbuf[1], cap * 8) == cudaSuccess &&
             cudaMalloc(&s->flags, size_t(nr) * 8) == cudaSuccess && cudaMalloc(&s->done, 4) == cudaSuccess;
    a = a && cudaMemset(s->flags, 0, size_t(nr) * 8) == cudaSuccess && cudaMemset(s->done, 0, 4) == cudaSuccess;
    a = a && cudaIpcGetMemHandle(&mine.h[0], s->xbuf[0]) == cudaSuccess &&
        cudaIpcGetMemHandle(&mine.h[1], s->xbuf[1]) == cudaSuccess &&
        cudaIpcGetMemHandle(&mine.h[2], s->flags) == cudaSuccess;
    if (!a) { cudaGetLastError(); good = 0; }
  }
  // exchange the handles (every rank takes part whatever its own outcome)
  std::vector<IpcPack> all(nr);
  {
    DevBuf<char> send(sizeof(IpcPack)), recv(sizeof(IpcPack) * size_t(nr));
    B2_CUDA(cudaMemcpyAsync(send, &mine, sizeof(IpcPack), cudaMemcpyHostToDevice, st));
    comm_allgather_bytes(ctx, send, recv, sizeof(IpcPack));
    B2_CUDA(cudaMemcpyAsync(all.data(), recv, sizeof(IpcPack) * size_t(nr), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
  }
  comm_allreduce_sum_i64_host(ctx, &good, 1);
  good = good == nr ? 1 : 0;
  std::vector<double*> px0(nr, nullptr), px1(nr, nullptr);
  std::vector<unsigned long long*> pf(nr, nullptr);
  if (good) {
    for (int r = 0; r < nr && good; ++r) {
      if (r == me) { px0[r] = s->xbuf[0]; px1[r] = s->xbuf[1]; pf[r] = s->flags; continue; }
      void* q[3] = {nullptr, nullptr, nullptr};
      for (int k = 0; k < 3; ++k) {
        if (cudaIpcOpenMemHandle(&q[k], all[r].h[k], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
          cudaGetLastError();
          good = 0;
          break;
        }
        s->opened.push_back(q[k]);
      }
      px0[r] = static_cast<double*>(q[0]);
      px1[r] = static_cast<double*>(q[1]);
      pf[r] = static_cast<unsigned long long*>(q[2]);
    }
  }
  if (good) {
    bool a = cudaMalloc(&s->d_peer_x[0], size_t(nr) * 8) == cudaSuccess &&
             cudaMalloc(&s->d_peer_x[1], size_t(nr) * 8) == cudaSuccess &&
             cudaMalloc(&s->d_peer_flag, size_t(nr) * 8) == cudaSuccess;
    a = a && cudaMemcpy(s->d_peer_x[0], px0.data(), size_t(nr) * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(s->d_peer_x[1], px1.data(), size_t(nr) * 8, cudaMemcpyHostToDevice) == cudaSuccess &&
        cudaMemcpy(s->d_peer_flag, pf.data(), size_t(nr) * 8, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!a) { cudaGetLastError(); good = 0; }
  }
  // all or nothing: one rank without peer access sends everybody to the NCCL path
  int64_t agree = good;
  comm_allreduce_sum_i64_host(ctx, &agree, 1);
  if (agree == nr) {
    s->ok = true;
    s->cap = cap;
  } else {
    p2p_release(ctx, s, true);  // a peer may have opened this rank's buffers before another rank failed
  }
  if (s->ok && !s->err_host) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&s->err_host), sizeof(unsigned int), cudaHostAllocMapped) == cudaSuccess &&
        cudaHostGetDevicePointer(reinterpret_cast<void**>(&s->err_dev), s->err_host, 0) == cudaSuccess) {
      *s->err_host = 0u;
    } else {
      cudaGetLastError();
      throw Error("p2p exchange: cannot allocate the mapped error word");
    }
    if (const char* env = getenv("B2CI_P2P_TIMEOUT_S")) s->timeout_ns = (unsigned long long)(std::max(1.0, atof(env)) * 1e9);
  }
  ctx->timers["comm.p2p"] = s->ok ? 1. : 0.;
}
}  // namespace

const double* comm_exchange_rows(b2ci_ctx* ctx, const double* local, const std::vector<int64_t>& off,
                                 double* fallback_full) {
  const int nr = ctx->nranks, me = ctx->rank;
  if (nr == 1) return local;
  const size_t n = size_t(off[nr]);
  P2P* s = p2p_state(ctx);
  if (!s->tried || (s->ok && s->cap < n)) p2p_setup(ctx, s, n);  // collective: n is the same everywhere
  if (s->ok) {
    if (*s->err_host) {
      const unsigned int who = *s->err_host - 1u;
      *s->err_host = 0u;
      throw Error("sigma exchange: rank " + std::to_string(who) + " did not publish its block of the trial vector within " +
                  std::to_string(s->timeout_ns / 1000000000ull) + " s (B2CI_P2P_TIMEOUT_S)");
    }
    const unsigned long long e = ++s->epoch;
    const int b = int(e & 1ull);
    const int64_t nloc = off[me + 1] - off[me];
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (nloc + PUSH_THREADS - 1) / PUSH_THREADS)));
    k_push<<<grid, PUSH_THREADS, 0, ctx->stream>>>(local, nloc, off[me], s->d_peer_x[b], s->d_peer_flag, nr, me, e, s->done);
    k_wait<<<1, 32, 0, ctx->stream>>>(s->flags, nr, e, s->timeout_ns, s->err_dev);
    ctx->launches += 2;
    B2_CHECK_LAUNCH();
    return s->xbuf[b];
  }
  double* full = fallback_full;
  if (!full) {
    if (s->fallback_cap < n) {
      if (s->fallback) cudaFree(s->fallback);
      s->fallback = nullptr;
      s->fallback_cap = 0;
      B2_CUDA(cudaMalloc(&s->fallback, n * 8));
      s->fallback_cap = n;
    }
    full = s->fallback;
  }
  comm_allgather_rows(ctx, local, full, off);
  return full;
}

// sigma of a row block: the peers' blocks of the trial vector arrive (peer-to-peer pushes) while the rank
// multiplies its own columns; the other columns follow once every block is there
void sigma_sharded(b2ci_ctx* ctx, b2ci_csr* m, const std::vector<int64_t>& off, const double* x_local,
                   double* x_full_or_null, double* y_local) {
  const int nr = ctx->nranks, me = ctx->rank;
  if (nr == 1) {
    spmv_launch(ctx, m, x_local, y_local);
    return;
  }
  const size_t n = size_t(off[nr]);
  P2P* s = p2p_state(ctx);
  if (!s->tried || (s->ok && s->cap < n)) p2p_setup(ctx, s, n);  // collective: n is the same everywhere
  // The overlapped form is OFF by default (B2CI_SIGMA_OVERLAP=1 turns it on). Measured on Cr2 CAS(12,12): the two
  // partial products run at ~5.3 TB/s against 6.0 TB/s for the single pass, which costs more than the ~30-80 us
  // exchange it hides -- sigma 2.27 vs 1.60 ms on 2 GPUs, 0.508 vs 0.434 ms per Davidson iteration on 8.
  bool overlap = false;
  if (const char* env = getenv("B2CI_SIGMA_OVERLAP")) overlap = atoi(env) != 0;
  if (getenv("B2CI_NO_SIGMA_OVERLAP")) overlap = false;
  if (s->ok && overlap) {
    if (*s->err_host) {
      const unsigned int who = *s->err_host - 1u;
      *s->err_host = 0u;
      throw Error("sigma exchange: rank " + std::to_string(who) + " did not publish its block of the trial vector within " +
                  std::to_string(s->timeout_ns / 1000000000ull) + " s (B2CI_P2P_TIMEOUT_S)");
    }
    spmv_prepare_parts(ctx, m);
    const unsigned long long e = ++s->epoch;
    const int b = int(e & 1ull);
    const int64_t nloc = off[me + 1] - off[me];
    const int grid = int(std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count, (nloc + PUSH_THREADS - 1) / PUSH_THREADS)));
    k_push<<<grid, PUSH_THREADS, 0, ctx->stream>>>(x_local, nloc, off[me], s->d_peer_x[b], s->d_peer_flag, nr, me, e, s->done);
    ctx->launches++;
    spmv_launch_part(ctx, m, 0, x_local, y_local);
    k_wait<<<1, 32, 0, ctx->stream>>>(s->flags, nr, e, s->timeout_ns, s->err_dev);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    spmv_launch_part(ctx, m, 1, s->xbuf[b], y_local);
    if (x_full_or_null)
      B2_CUDA(cudaMemcpyAsync(x_full_or_null, s->xbuf[b], n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
    return;
  }
  const double* xg = comm_exchange_rows(ctx, x_local, off, x_full_or_null);
  if (x_full_or_null && xg != x_full_or_null)
    B2_CUDA(cudaMemcpyAsync(x_full_or_null, xg, n * 8, cudaMemcpyDeviceToDevice, ctx->stream));
  spmv_launch(ctx, m, xg, y_local);
}

void comm_destroy(b2ci_ctx* ctx) {
  if (ctx->p2p) {
    P2P* s = static_cast<P2P*>(ctx->p2p);
    p2p_release(ctx, s, s->ok);
    if (s->fallback) cudaFree(s->fallback);
    if (s->err_host) cudaFreeHost(s->err_host);
    delete s;
    ctx->p2p = nullptr;
  }
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

void comm_allreduce_sum_u64(b2ci_ctx* ctx, unsigned long long* dev_buf, int64_t n) {
  if (ctx->nranks == 1 || n == 0) return;
  B2_NCCL(api().AllReduce(dev_buf, dev_buf, size_t(n), ncclUint64, ncclSum, (ncclComm_t)ctx->nccl_comm, ctx->stream));
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
