// Shared declarations of the b2ci CUDA library (sm_100a). Internal header.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/b2ci.h"

namespace b2ci {

void set_error(const std::string& msg);

struct Error : std::runtime_error {
  int code;
  explicit Error(const std::string& m, int c = 1) : std::runtime_error(m), code(c) {}
};

#define B2_CUDA(expr)                                                                    \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      throw ::b2ci::Error(std::string(#expr) + ": " + cudaGetErrorString(_e) + " (" +    \
                          __FILE__ + ":" + std::to_string(__LINE__) + ")");              \
  } while (0)

#define B2_CHECK_LAUNCH() B2_CUDA(cudaGetLastError())

// Integral tables on the device, one contiguous allocation so the small ones can be staged
// into shared memory with a single bulk copy:
//   [ T n^2 | G2_red n^2 | V2_red n^2 | G_red n^3 | V_red n^3 | V n^4 | Vt n^2 * n2p ]
// Vt is V with the index pairs exchanged, Vt[(r + s n) + (p + q n) n2p] = V(p,q,r,s) (a copy,
// no arithmetic), n2p = n^2 rounded up to even: the n^2 integrals (pq|..) that one alpha single
// excitation p<-q needs are then one contiguous, 16-byte aligned slice -- the unit the H-build
// kernel stages into shared memory with bulk (TMA) copies.
struct IntsView {
  int n;
  int n2p;
  const double* Vt;
  const double* T;
  const double* G2;
  const double* V2;
  const double* G;   // G_red(k,i,j) at k + i n + j n^2
  const double* Vr;  // V_red(k,i,j)
  const double* V;   // (pq|rs) at p + q n + r n^2 + s n^3
};

inline size_t ints_small_doubles(int n) {  // T, G2, V2, G_red, V_red
  size_t n2 = size_t(n) * n;
  return 3 * n2 + 2 * n2 * n;
}
inline size_t ints_n2p(int n) { return (size_t(n) * n + 1) & ~size_t(1); }
inline size_t ints_vt_offset(int n) {  // even, so Vt is 16-byte aligned
  size_t n2 = size_t(n) * n;
  return (ints_small_doubles(n) + n2 * n2 + 1) & ~size_t(1);
}
inline size_t ints_total_doubles(int n) {
  size_t n2 = size_t(n) * n;
  return ints_vt_offset(n) + n2 * ints_n2p(n);
}
inline IntsView make_view(int n, const double* base) {
  size_t n2 = size_t(n) * n, n3 = n2 * n;
  IntsView v;
  v.n = n;
  v.T = base;
  v.G2 = base + n2;
  v.V2 = base + 2 * n2;
  v.G = base + 3 * n2;
  v.Vr = base + 3 * n2 + n3;
  v.V = base + 3 * n2 + 2 * n3;
  v.n2p = int(ints_n2p(n));
  v.Vt = base + ints_vt_offset(n);
  return v;
}

struct EventTimer {
  cudaEvent_t a = nullptr, b = nullptr;
};

struct NcclApi;  // comm.cu

}  // namespace b2ci

struct b2ci_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  int sm_count = 148;
  int norb = 0;
  double* ints_dev = nullptr;  // layout above
  b2ci::IntsView ints{};
  std::vector<double> ints_host;  // same layout, host copy (host-side evaluation / ASCI driver)
  int64_t launches = 0;
  int generator = 0;  // B2CI_GEN_*: pattern / threshold rules of the H build (b2ci_set_hamiltonian_generator)
  std::map<std::string, double> timers;
  // multi-GPU
  void* nccl_comm = nullptr;
  int rank = 0, nranks = 1;
  // structural slot arrays of the last thresholded H build (kept when rows had to be packed:
  // handing multi-GB blocks back to the pool between builds fragments it and makes it grow)
  void* slot_cache[2] = {nullptr, nullptr};
  size_t slot_cache_bytes[2] = {0, 0};
  // pinned staging for the scalars a build reads back (pageable targets would make every
  // cudaMemcpyAsync a synchronisation of its own)
  int64_t* pinned = nullptr;
  // grow-only workspace of the ASCI search (plain cudaMalloc): the candidate tables grow 8x per
  // ASCI iteration, and growing the stream-ordered pool by tens of GB costs ~1 s per step
  char* arena = nullptr;
  size_t arena_cap = 0, arena_off = 0;
  // peer-to-peer exchange state of the sharded sigma (comm.cu), NULL until first use
  void* p2p = nullptr;
};

struct b2ci_dets {
  int64_t n = 0;
  uint64_t* alpha = nullptr;  // device
  uint64_t* beta = nullptr;   // device
};

struct b2ci_csr {
  int64_t nrows = 0, ncols = 0, nnz = 0, row_begin = 0;
  int64_t* rowptr = nullptr;  // device, nrows + 1, local offsets (rowptr[0] == 0)
  int32_t* colind = nullptr;  // device, global column indices
  double* nzval = nullptr;    // device
  std::vector<int64_t> row_offsets;  // multi-GPU: row offsets of all ranks (lazy)
  size_t colind_cap = 0, nzval_cap = 0;  // allocated bytes when known (recycled through the context)
  void* loc_range = nullptr;  // device, nrows x int2: own-column sub-range of every row (sharded sigma, lazy)
  // rows sorted into length classes for the binned SpMV (skewed matrices only; lazy)
  void* bin_list = nullptr;   // device, nrows x int32: row indices, class after class
  int64_t bin_off[5] = {0, 0, 0, 0, 0};
  bool bins_tried = false;
};

namespace b2ci {

// Device memory comes from the device's stream-ordered pool (cudaMallocAsync) with the release
// threshold lifted in b2ci_ctx_create, so the multi-GB CSR / slot buffers of repeated builds are
// recycled instead of going back to the driver (a cudaMalloc + cudaFree pair of 14 GB costs
// ~100 ms, ten times the kernels it serves). Every C-ABI entry that takes a context opens a
// StreamScope; allocations and frees are ordered on that context's stream.
cudaStream_t& alloc_stream();  // thread-local, set by StreamScope
// Large blocks (>= 32 MiB) bypass the stream-ordered pool: growing that pool costs ~50 ms per GB
// (mapping granule by granule), a plain cudaMalloc ~7 ms per GB. They are cached per (device,
// stream) in a process-wide free list and handed out again to requests of similar size; reuse on
// the same stream is ordered by the stream itself. capi.cu.
void* big_cache_alloc(size_t bytes, cudaStream_t st);
bool big_cache_free(void* p, cudaStream_t st);   // false: not a cached block
void big_cache_trim();                            // cudaFree every idle block of the current device
constexpr size_t BIG_BLOCK_MIN = size_t(32) << 20;
inline void* dev_alloc(size_t bytes) {
  if (bytes >= BIG_BLOCK_MIN) return big_cache_alloc(bytes, alloc_stream());
  void* p = nullptr;
  cudaError_t e = cudaMallocAsync(&p, bytes, alloc_stream());
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error("device allocation of " + std::to_string(bytes) + " bytes failed: " + cudaGetErrorString(e));
  }
  return p;
}
inline void dev_free(void* p) {
  if (!p) return;
  if (big_cache_free(p, alloc_stream())) return;
  cudaFreeAsync(p, alloc_stream());
}
struct StreamScope {
  cudaStream_t prev;
  explicit StreamScope(const b2ci_ctx* ctx) : prev(alloc_stream()) {
    if (ctx) {
      cudaSetDevice(ctx->device);
      alloc_stream() = ctx->stream;
    }
  }
  ~StreamScope() { alloc_stream() = prev; }
};

// RAII device buffer
template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  explicit DevBuf(size_t count) { alloc(count); }
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), n(o.n) { o.p = nullptr; o.n = 0; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) { release(); p = o.p; n = o.n; o.p = nullptr; o.n = 0; }
    return *this;
  }
  ~DevBuf() { release(); }
  void alloc(size_t count) {
    release();
    n = count;
    if (count) p = static_cast<T*>(dev_alloc(count * sizeof(T)));
  }
  void release() {
    dev_free(p);
    p = nullptr;
    n = 0;
  }
  T* take() { T* q = p; p = nullptr; n = 0; return q; }
  operator T*() const { return p; }
};

// [0] colind-like (4 B / entry), [1] nzval-like (8 B / entry): take a cached block or allocate
inline void* big_alloc(b2ci_ctx* ctx, int which, size_t bytes, size_t* cap) {
  void* q = nullptr;
  if (ctx->slot_cache[which] && ctx->slot_cache_bytes[which] >= bytes) {
    q = ctx->slot_cache[which];
    *cap = ctx->slot_cache_bytes[which];
  } else {
    dev_free(ctx->slot_cache[which]);
    q = dev_alloc(bytes);
    *cap = bytes;
  }
  ctx->slot_cache[which] = nullptr;
  ctx->slot_cache_bytes[which] = 0;
  return q;
}
// give a block back: the context keeps the larger one, the other returns to the pool
inline void big_release(b2ci_ctx* ctx, int which, void* p, size_t cap) {
  if (!p) return;
  if (cap > ctx->slot_cache_bytes[which]) {
    dev_free(ctx->slot_cache[which]);
    ctx->slot_cache[which] = p;
    ctx->slot_cache_bytes[which] = cap;
  } else {
    dev_free(p);
  }
}
// bump allocation from the context's workspace; arena_reset() at the start of a phase
inline void arena_reserve(b2ci_ctx* ctx, size_t bytes) {
  if (bytes <= ctx->arena_cap) return;
  cudaStreamSynchronize(ctx->stream);
  if (ctx->arena) cudaFree(ctx->arena);
  ctx->arena = nullptr;
  ctx->arena_cap = 0;
  const size_t want = bytes + bytes / 8;
  void* p = nullptr;
  if (cudaMalloc(&p, want) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
      cudaGetLastError();
      throw Error("workspace allocation of " + std::to_string(bytes >> 20) + " MiB failed");
    }
    ctx->arena_cap = bytes;
  } else {
    ctx->arena_cap = want;
  }
  ctx->arena = static_cast<char*>(p);
}
inline void arena_reset(b2ci_ctx* ctx) { ctx->arena_off = 0; }
template <typename T>
inline T* arena_take(b2ci_ctx* ctx, size_t count) {
  const size_t bytes = (count * sizeof(T) + 255) & ~size_t(255);
  if (ctx->arena_off + bytes > ctx->arena_cap) throw Error("workspace overflow (internal sizing error)");
  T* p = reinterpret_cast<T*>(ctx->arena + ctx->arena_off);
  ctx->arena_off += bytes;
  return p;
}
constexpr int PINNED_WORDS = 64;
inline int64_t* pinned_words(b2ci_ctx* ctx) {
  if (!ctx->pinned) {
    if (cudaHostAlloc(reinterpret_cast<void**>(&ctx->pinned), PINNED_WORDS * 8, cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      throw Error("pinned staging allocation failed");
    }
  }
  return ctx->pinned;
}

struct ScopedTimer {  // CUDA-event timing of a phase on the context stream
  b2ci_ctx* ctx;
  std::string name;
  cudaEvent_t a, b;
  bool accumulate;
  ScopedTimer(b2ci_ctx* c, const char* nm, bool acc = false) : ctx(c), name(nm), accumulate(acc) {
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, ctx->stream);
  }
  ~ScopedTimer() {
    cudaEventRecord(b, ctx->stream);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    if (accumulate) ctx->timers[name] += ms; else ctx->timers[name] = ms;
    cudaEventDestroy(a);
    cudaEventDestroy(b);
  }
};

// Phase timing without synchronising inside the phase loop: event pairs are recorded on the
// context stream and resolved once, after the caller's final synchronisation.
struct DeferredTimers {
  b2ci_ctx* ctx;
  struct Rec { std::string name; cudaEvent_t a, b; };
  std::vector<Rec> recs;
  explicit DeferredTimers(b2ci_ctx* c) : ctx(c) {}
  size_t start(const char* name) {
    Rec r;
    r.name = name;
    cudaEventCreate(&r.a);
    cudaEventCreate(&r.b);
    cudaEventRecord(r.a, ctx->stream);
    recs.push_back(r);
    return recs.size() - 1;
  }
  void stop(size_t id) { cudaEventRecord(recs[id].b, ctx->stream); }
  void resolve() {  // after a stream synchronisation
    for (auto& r : recs) {
      float ms = 0.f;
      if (cudaEventSynchronize(r.b) == cudaSuccess && cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess)
        ctx->timers[r.name] += ms;
      cudaEventDestroy(r.a);
      cudaEventDestroy(r.b);
    }
    recs.clear();
  }
  ~DeferredTimers() { resolve(); }
};
struct DeferredScope {
  DeferredTimers& t;
  size_t id;
  DeferredScope(DeferredTimers& tt, const char* name) : t(tt), id(tt.start(name)) {}
  ~DeferredScope() { t.stop(id); }
};

#ifdef __CUDACC__
// ---- bulk (TMA) copy + mbarrier primitives (PTX ISA 8.0+, sm_90+; SASS: UBLKCP / SYNCS)
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return uint32_t(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}"
      ::"r"(smem_u32(bar)), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

#endif

// scan.cu : out has n + 1 entries, out[n] = total
void exclusive_scan_i32_to_i64(b2ci_ctx* ctx, const int32_t* in, int64_t* out, int64_t n);
void exclusive_scan_i32(b2ci_ctx* ctx, const int32_t* in, int32_t* out, int64_t n);

// spmv.cu
void spmv_launch(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y);
void spmv_prepare_parts(b2ci_ctx* ctx, b2ci_csr* m);
void spmv_launch_part(b2ci_ctx* ctx, const b2ci_csr* m, int part, const double* x, double* y);
// sharded sigma: exchange of the trial vector overlapped with the own-column part of the product
void sigma_sharded(b2ci_ctx* ctx, b2ci_csr* m, const std::vector<int64_t>& row_offsets, const double* x_local,
                   double* x_full_or_null, double* y_local);

// eig.cpp-ish (davidson.cu): symmetric eigensolver, lower triangle, ascending
void sym_eig_lower(int n, double* A, int lda, double* W);
void sym_eig_lowest(int n, const double* A, int lda, double* lambda, double* vec);

// comm.cu
void comm_allgather_rows(b2ci_ctx* ctx, const double* local, double* full,
                         const std::vector<int64_t>& row_offsets);
void comm_allreduce_sum(b2ci_ctx* ctx, double* dev_buf, int64_t n);
// Gather the row blocks of all ranks: returns a device pointer to the full vector. With peer
// access every rank writes its block straight into the others' buffers over NVLink (one small
// kernel + a flag wait); otherwise NCCL gathers into `fallback_full` (or a library buffer).
const double* comm_exchange_rows(b2ci_ctx* ctx, const double* local, const std::vector<int64_t>& off,
                                 double* fallback_full);
void comm_allreduce_sum_i64_host(b2ci_ctx* ctx, int64_t* host_vals, int n);
void comm_allreduce_sum_u64(b2ci_ctx* ctx, unsigned long long* dev_buf, int64_t n);
void comm_allgather_i64_host(b2ci_ctx* ctx, int64_t local, std::vector<int64_t>& all);
void comm_allgather_bytes(b2ci_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);

}  // namespace b2ci
