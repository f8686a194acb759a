// extern "C" surface declared in include/b2ci.h. Every entry catches C++ exceptions,
// records the message for b2ci_last_error() and returns a non-zero status -- there is no
// CPU fallback behind any of them.
#include <nvtx3/nvToolsExt.h>
#include <unordered_map>
#include <mutex>
#include <cstring>

#include "common.cuh"
#include "slater.cuh"

namespace b2ci {

static thread_local std::string g_last_error;
void set_error(const std::string& msg) { g_last_error = msg; }
namespace {
struct BigBlock { void* p; size_t bytes; int device; cudaStream_t stream; };
struct BigCache {
  std::mutex mu;
  std::vector<BigBlock> idle;
  std::unordered_map<void*, BigBlock> live;
};
BigCache& big_cache() {
  static BigCache* c = new BigCache;  // never destroyed: frees at exit would race the driver
  return *c;
}
}  // namespace
void* big_cache_alloc(size_t bytes, cudaStream_t st) {
  BigCache& C = big_cache();
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(C.mu);
  // best fit among idle blocks of this device and stream that are not wastefully large
  int best = -1;
  for (int i = 0; i < int(C.idle.size()); ++i) {
    const BigBlock& b = C.idle[i];
    if (b.device != dev || b.stream != st || b.bytes < bytes || b.bytes > bytes + bytes / 2 + (size_t(64) << 20)) continue;
    if (best < 0 || b.bytes < C.idle[best].bytes) best = i;
  }
  if (best >= 0) {
    BigBlock b = C.idle[best];
    C.idle.erase(C.idle.begin() + best);
    C.live[b.p] = b;
    return b.p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  if (e != cudaSuccess) {  // give the idle blocks back and retry once
    cudaGetLastError();
    cudaDeviceSynchronize();
    for (auto it = C.idle.begin(); it != C.idle.end();) {
      if (it->device == dev) { cudaFree(it->p); it = C.idle.erase(it); } else ++it;
    }
    e = cudaMalloc(&p, bytes);
  }
  if (e != cudaSuccess) {
    cudaGetLastError();
    throw Error("device allocation of " + std::to_string(bytes >> 20) + " MiB failed: " + cudaGetErrorString(e));
  }
  C.live[p] = BigBlock{p, bytes, dev, st};
  return p;
}
bool big_cache_free(void* p, cudaStream_t st) {
  BigCache& C = big_cache();
  std::lock_guard<std::mutex> g(C.mu);
  auto it = C.live.find(p);
  if (it == C.live.end()) return false;
  BigBlock b = it->second;
  C.live.erase(it);
  b.stream = st;  // later work on this stream is ordered behind whatever still reads the block
  C.idle.push_back(b);
  return true;
}
void big_cache_trim() {
  BigCache& C = big_cache();
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> g(C.mu);
  cudaDeviceSynchronize();
  for (auto it = C.idle.begin(); it != C.idle.end();) {
    if (it->device == dev) { cudaFree(it->p); it = C.idle.erase(it); } else ++it;
  }
}

cudaStream_t& alloc_stream() {
  static thread_local cudaStream_t s = nullptr;
  return s;
}

// implemented in the other translation units
void integrals_upload(b2ci_ctx* ctx, int norb, const double* T, const double* V);
void integrals_rotate(b2ci_ctx* ctx, const double* C, double* T_out, double* V_out);
void dets_from_words(b2ci_ctx* ctx, const uint64_t* words_host, int wpd, int64_t n, b2ci_dets* d);
void dets_to_words(b2ci_ctx* ctx, const b2ci_dets* d, uint64_t* words_host, int wpd);
void dets_generate_fci(b2ci_ctx* ctx, int norb, int na, int nb, b2ci_dets* d);
void dets_balanced_partition(b2ci_ctx* ctx, const b2ci_dets* dets, int nparts, int64_t nsamples, int64_t* offsets);
void hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end, double thr,
                b2ci_csr* out);
bool hbuild_csr_patched(b2ci_ctx* ctx, const b2ci_dets* od, const b2ci_csr* oH, const b2ci_dets* nd, double thr,
                        double min_overlap, b2ci_csr* out, int64_t* n_kept_out);
void csr_diagonal_dev(b2ci_ctx* ctx, const b2ci_csr* m, double* D_dev);
int davidson(b2ci_ctx* ctx, const b2ci_csr* m, int64_t max_m, double tol, double* X_host,
             int use_guess_policy, int64_t* niter_out, double* eig_out, double* trace);
void dense_ground_state(b2ci_ctx* ctx, const b2ci_csr* m, double* eigval, double* eigvec_host);
void comm_unique_id(void* id128);
void comm_init(b2ci_ctx* ctx, const void* id128, int rank, int nranks);
void comm_destroy(b2ci_ctx* ctx);
int asci_search(b2ci_ctx* ctx, const b2ci_asci_search_opts* o, const uint64_t* core_words, int wpd,
                const double* coeffs, int64_t ncdets, double E0, uint64_t* out_words, int64_t cap,
                int64_t* n_out, double* stats, uint64_t* cand_words, double* cand_cm, double* cand_hd,
                int64_t* cand_n, bool candidates_only, double* pt2_out = nullptr);

void form_rdms(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C_host, bool spin_dep, double* o1,
               double* o2, double* t1, double* t2, double* t3);

size_t entropy_intermediate_doubles(int n, bool need_s2);
void entropy_intermediates(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C_host, bool need_s2, double* out_host);
void host_entropies_from_intermediates(int n, bool need_s2, const double* I, double* s1, double* s2, double* mi);

namespace {
__global__ void k_i32_to_i64(const int32_t* __restrict__ in, int64_t n, int64_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = in[i];
}
__global__ void k_i64_to_i32(const int64_t* __restrict__ in, int64_t n, int32_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = int32_t(in[i]);
}
}  // namespace
}  // namespace b2ci

using namespace b2ci;

#define B2_TRY try {
// profiler range per C-ABI entry, named after the entry (NVTX3 is header-only: a no-op unless a tool such as
// nsys / ncu --nvtx injects itself). The reference marks the same phases with its h_build / ci_solver / davidson /
// asci_search loggers (selected_ci_diag.hpp:204-296).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
};
#define B2_TRY_CTX(ctx)                                            \
  try {                                                            \
    if (!(ctx)) throw b2ci::Error("null b2ci context");            \
    NvtxRange _nvtx(__func__);                                     \
    b2ci::StreamScope _scope(ctx);
#define B2_CATCH                                   \
  }                                                \
  catch (const b2ci::Error& e) {                   \
    b2ci::set_error(e.what());                     \
    return e.code ? e.code : 1;                    \
  }                                                \
  catch (const std::exception& e) {                \
    b2ci::set_error(e.what());                     \
    return 1;                                      \
  }

extern "C" {

const char* b2ci_last_error(void) { return b2ci::g_last_error.c_str(); }
const char* b2ci_version(void) { return "b2ci 0.1 (sm_100a)"; }

int b2ci_ctx_create(int device, void* stream, b2ci_ctx** out) {
  B2_TRY
  if (!out) throw Error("b2ci_ctx_create: null output");
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    throw Error(std::string("b2ci_ctx_create: no CUDA device available (") + cudaGetErrorString(e) +
                "); this library has no CPU fallback");
  if (device < 0 || device >= ndev) throw Error("b2ci_ctx_create: bad device index");
  B2_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  B2_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    throw Error("b2ci_ctx_create: kernels are built for sm_100a only; device is sm_" +
                std::to_string(prop.major) + std::to_string(prop.minor));
  // keep freed blocks in the device's stream-ordered pool (see common.cuh)
  cudaMemPool_t pool;
  B2_CUDA(cudaDeviceGetDefaultMemPool(&pool, device));
  uint64_t keep = UINT64_MAX;
  B2_CUDA(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep));
  b2ci_ctx* c = new b2ci_ctx;
  c->device = device;
  c->stream = (cudaStream_t)stream;
  c->sm_count = prop.multiProcessorCount;
  *out = c;
  return 0;
  B2_CATCH
}
int b2ci_ctx_destroy(b2ci_ctx* ctx) {
  B2_TRY
  if (!ctx) return 0;
  {
    StreamScope scope(ctx);
    comm_destroy(ctx);
    dev_free(ctx->ints_dev);
    for (int i = 0; i < 2; ++i) dev_free(ctx->slot_cache[i]);
    if (ctx->pinned) cudaFreeHost(ctx->pinned);
    if (ctx->arena) cudaFree(ctx->arena);
    cudaStreamSynchronize(ctx->stream);
  }
  delete ctx;
  return 0;
  B2_CATCH
}
int b2ci_ctx_synchronize(b2ci_ctx* ctx) {
  B2_TRY_CTX(ctx)
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
  B2_CATCH
}
int64_t b2ci_ctx_launch_count(const b2ci_ctx* ctx) { return ctx ? ctx->launches : 0; }
int b2ci_ctx_trim(b2ci_ctx* ctx) {
  B2_TRY_CTX(ctx)
  for (int i = 0; i < 2; ++i) {
    dev_free(ctx->slot_cache[i]);
    ctx->slot_cache[i] = nullptr;
    ctx->slot_cache_bytes[i] = 0;
  }
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  if (ctx->arena) cudaFree(ctx->arena);
  ctx->arena = nullptr;
  ctx->arena_cap = ctx->arena_off = 0;
  big_cache_trim();
  cudaMemPool_t pool;
  B2_CUDA(cudaDeviceGetDefaultMemPool(&pool, ctx->device));
  B2_CUDA(cudaMemPoolTrimTo(pool, 0));
  return 0;
  B2_CATCH
}

int b2ci_comm_unique_id(void* id128) {
  B2_TRY
  comm_unique_id(id128);
  return 0;
  B2_CATCH
}
int b2ci_comm_init(b2ci_ctx* ctx, const void* id128, int rank, int nranks) {
  B2_TRY_CTX(ctx)
  comm_init(ctx, id128, rank, nranks);
  return 0;
  B2_CATCH
}
int b2ci_comm_rank(const b2ci_ctx* ctx, int* rank, int* nranks) {
  if (rank) *rank = ctx->rank;
  if (nranks) *nranks = ctx->nranks;
  return 0;
}

int b2ci_integrals_rotate(b2ci_ctx* ctx, const double* C, double* T_out, double* V_out) {
  B2_TRY_CTX(ctx)
  integrals_rotate(ctx, C, T_out, V_out);
  return 0;
  B2_CATCH
}
int b2ci_integrals_upload(b2ci_ctx* ctx, int norb, const double* T, const double* V) {
  B2_TRY_CTX(ctx)
  integrals_upload(ctx, norb, T, V);
  return 0;
  B2_CATCH
}
int b2ci_integrals_download(b2ci_ctx* ctx, double* G_red, double* V_red, double* G2_red,
                            double* V2_red) {
  B2_TRY_CTX(ctx)
  if (!ctx->ints_dev) throw Error("integrals not uploaded");
  const size_t n = ctx->norb, n2 = n * n, n3 = n2 * n;
  const IntsView h = make_view(ctx->norb, ctx->ints_host.data());
  if (G_red) memcpy(G_red, h.G, n3 * 8);
  if (V_red) memcpy(V_red, h.Vr, n3 * 8);
  if (G2_red) memcpy(G2_red, h.G2, n2 * 8);
  if (V2_red) memcpy(V2_red, h.V2, n2 * 8);
  return 0;
  B2_CATCH
}

int b2ci_dets_upload(b2ci_ctx* ctx, const uint64_t* words, int wpd, int64_t n, b2ci_dets** out) {
  B2_TRY_CTX(ctx)
  if (n < 0 || (n > 0 && !words)) throw Error("b2ci_dets_upload: bad arguments");
  b2ci_dets* d = new b2ci_dets;
  try { dets_from_words(ctx, words, wpd, n, d); } catch (...) { delete d; throw; }
  *out = d;
  return 0;
  B2_CATCH
}
int b2ci_dets_generate_fci(b2ci_ctx* ctx, int norb, int na, int nb, b2ci_dets** out) {
  B2_TRY_CTX(ctx)
  b2ci_dets* d = new b2ci_dets;
  try { dets_generate_fci(ctx, norb, na, nb, d); } catch (...) { delete d; throw; }
  *out = d;
  return 0;
  B2_CATCH
}
int b2ci_dets_size(const b2ci_dets* d, int64_t* n) {
  *n = d->n;
  return 0;
}
int b2ci_dets_download(b2ci_ctx* ctx, const b2ci_dets* d, uint64_t* words, int wpd) {
  B2_TRY_CTX(ctx)
  dets_to_words(ctx, d, words, wpd);
  return 0;
  B2_CATCH
}
int b2ci_dets_free(b2ci_ctx* ctx, b2ci_dets* d) {
  if (!d) return 0;
  StreamScope scope(ctx);
  dev_free(d->alpha);
  dev_free(d->beta);
  delete d;
  return 0;
}

int b2ci_hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end,
                    double h_thresh, b2ci_csr** out) {
  B2_TRY_CTX(ctx)
  b2ci_csr* m = new b2ci_csr;
  try { hbuild_csr(ctx, dets, row_begin, row_end, h_thresh, m); } catch (...) { delete m; throw; }
  *out = m;
  return 0;
  B2_CATCH
}
int b2ci_set_hamiltonian_generator(b2ci_ctx* ctx, int generator) {
  B2_TRY_CTX(ctx)
  if (generator < 0 || generator > 2) throw Error("b2ci_set_hamiltonian_generator: unknown generator");
  ctx->generator = generator;
  return 0;
  B2_CATCH
}
int b2ci_hbuild_csr_patched(b2ci_ctx* ctx, const b2ci_dets* old_dets, const b2ci_csr* old_H,
                            const b2ci_dets* new_dets, double h_thresh, double min_overlap,
                            b2ci_csr** out, int64_t* n_kept) {
  B2_TRY_CTX(ctx)
  if (!out) throw Error("b2ci_hbuild_csr_patched: out is NULL");
  *out = nullptr;
  b2ci_csr* m = new b2ci_csr;
  bool built = false;
  try { built = hbuild_csr_patched(ctx, old_dets, old_H, new_dets, h_thresh, min_overlap, m, n_kept); }
  catch (...) { delete m; throw; }
  if (built) *out = m; else delete m;
  return 0;
  B2_CATCH
}
int b2ci_csr_upload(b2ci_ctx* ctx, int64_t n, int64_t nnz, const int64_t* rowptr,
                    const int64_t* colind, const double* nzval, b2ci_csr** out) {
  B2_TRY_CTX(ctx)
  if (n < 0 || nnz < 0 || !rowptr) throw Error("b2ci_csr_upload: bad arguments");
  if (n >= (int64_t(1) << 31)) throw Error("b2ci_csr_upload: dimension exceeds int32 column indices");
  if (rowptr[0] != 0 || rowptr[n] != nnz) throw Error("b2ci_csr_upload: rowptr must be 0-based and end at nnz");
  DevBuf<int64_t> rp(n + 1), ci64(nnz > 0 ? nnz : 1);
  DevBuf<int32_t> ci(nnz > 0 ? nnz : 1);
  DevBuf<double> nz(nnz > 0 ? nnz : 1);
  cudaStream_t st = ctx->stream;
  B2_CUDA(cudaMemcpyAsync(rp, rowptr, size_t(n + 1) * 8, cudaMemcpyHostToDevice, st));
  if (nnz) {
    B2_CUDA(cudaMemcpyAsync(ci64, colind, size_t(nnz) * 8, cudaMemcpyHostToDevice, st));
    B2_CUDA(cudaMemcpyAsync(nz, nzval, size_t(nnz) * 8, cudaMemcpyHostToDevice, st));
    k_i64_to_i32<<<unsigned((nnz + 255) / 256), 256, 0, st>>>(ci64, nnz, ci);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }
  B2_CUDA(cudaStreamSynchronize(st));
  b2ci_csr* m = new b2ci_csr;
  m->nrows = n; m->ncols = n; m->nnz = nnz; m->row_begin = 0;
  m->rowptr = rp.take(); m->colind = ci.take(); m->nzval = nz.take();
  *out = m;
  return 0;
  B2_CATCH
}
int b2ci_csr_info(const b2ci_csr* m, int64_t* nrows, int64_t* ncols, int64_t* nnz,
                  int64_t* row_begin) {
  if (nrows) *nrows = m->nrows;
  if (ncols) *ncols = m->ncols;
  if (nnz) *nnz = m->nnz;
  if (row_begin) *row_begin = m->row_begin;
  return 0;
}
int b2ci_csr_download(b2ci_ctx* ctx, const b2ci_csr* m, int64_t* rowptr, int64_t* colind,
                      double* nzval) {
  B2_TRY_CTX(ctx)
  cudaStream_t st = ctx->stream;
  if (rowptr) B2_CUDA(cudaMemcpyAsync(rowptr, m->rowptr, size_t(m->nrows + 1) * 8, cudaMemcpyDeviceToHost, st));
  if (m->nnz && colind) {
    // widen on the device in bounded chunks so the staging buffer stays small
    const int64_t chunk = int64_t(1) << 26;
    DevBuf<int64_t> tmp(std::min<int64_t>(chunk, m->nnz));
    for (int64_t off = 0; off < m->nnz; off += chunk) {
      const int64_t c = std::min<int64_t>(chunk, m->nnz - off);
      k_i32_to_i64<<<unsigned((c + 255) / 256), 256, 0, st>>>(m->colind + off, c, tmp);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      B2_CUDA(cudaMemcpyAsync(colind + off, tmp, size_t(c) * 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
    }
  }
  if (m->nnz && nzval) B2_CUDA(cudaMemcpyAsync(nzval, m->nzval, size_t(m->nnz) * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return 0;
  B2_CATCH
}
int b2ci_csr_device_ptrs(const b2ci_csr* m, const int64_t** rowptr, const int32_t** colind,
                         const double** nzval) {
  if (rowptr) *rowptr = m->rowptr;
  if (colind) *colind = m->colind;
  if (nzval) *nzval = m->nzval;
  return 0;
}
int b2ci_csr_free(b2ci_ctx* ctx, b2ci_csr* m) {
  if (!m) return 0;
  StreamScope scope(ctx);
  dev_free(m->rowptr);
  dev_free(m->loc_range);
  dev_free(m->bin_list);
  if (ctx && m->colind_cap) big_release(ctx, 0, m->colind, m->colind_cap); else dev_free(m->colind);
  if (ctx && m->nzval_cap) big_release(ctx, 1, m->nzval, m->nzval_cap); else dev_free(m->nzval);
  delete m;
  return 0;
}

int b2ci_spmv(b2ci_ctx* ctx, const b2ci_csr* m, const double* x_dev, double* y_dev) {
  B2_TRY_CTX(ctx)
  spmv_launch(ctx, m, x_dev, y_dev);
  return 0;
  B2_CATCH
}
int b2ci_spmv_host(b2ci_ctx* ctx, const b2ci_csr* m, const double* x, double* y) {
  B2_TRY_CTX(ctx)
  DevBuf<double> dx(m->ncols > 0 ? m->ncols : 1), dy(m->nrows > 0 ? m->nrows : 1);
  cudaStream_t st = ctx->stream;
  B2_CUDA(cudaMemcpyAsync(dx, x, size_t(m->ncols) * 8, cudaMemcpyHostToDevice, st));
  spmv_launch(ctx, m, dx, dy);
  B2_CUDA(cudaMemcpyAsync(y, dy, size_t(m->nrows) * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return 0;
  B2_CATCH
}
int b2ci_sigma_sharded(b2ci_ctx* ctx, const b2ci_csr* m, const double* x_local_dev,
                       double* x_full_dev, double* y_local_dev) {
  B2_TRY_CTX(ctx)
  if (ctx->nranks == 1) {
    spmv_launch(ctx, m, x_local_dev, y_local_dev);
    return 0;
  }
  b2ci_csr* mm = const_cast<b2ci_csr*>(m);
  if (mm->row_offsets.empty()) {
    std::vector<int64_t> counts;
    comm_allgather_i64_host(ctx, m->nrows, counts);
    mm->row_offsets.assign(ctx->nranks + 1, 0);
    for (int r = 0; r < ctx->nranks; ++r) mm->row_offsets[r + 1] = mm->row_offsets[r] + counts[r];
    if (mm->row_offsets[ctx->rank] != m->row_begin || mm->row_offsets[ctx->nranks] != m->ncols)
      throw Error("b2ci_sigma_sharded: row blocks of the ranks do not tile [0, ncols) in rank order");
  }
  sigma_sharded(ctx, mm, mm->row_offsets, x_local_dev, x_full_dev, y_local_dev);
  return 0;
  B2_CATCH
}
int b2ci_csr_set_row_partition(b2ci_ctx* ctx, b2ci_csr* m, const int64_t* row_offsets, int nranks) {
  B2_TRY_CTX(ctx)
  if (!m || !row_offsets || nranks != ctx->nranks) throw Error("b2ci_csr_set_row_partition: bad arguments");
  for (int r = 0; r < nranks; ++r)
    if (row_offsets[r + 1] < row_offsets[r]) throw Error("b2ci_csr_set_row_partition: offsets must ascend");
  if (row_offsets[0] != 0 || row_offsets[nranks] != m->ncols || row_offsets[ctx->rank] != m->row_begin ||
      row_offsets[ctx->rank + 1] != m->row_begin + m->nrows)
    throw Error("b2ci_csr_set_row_partition: offsets do not match this rank's row block");
  m->row_offsets.assign(row_offsets, row_offsets + nranks + 1);
  return 0;
  B2_CATCH
}
int b2ci_dets_balanced_partition(b2ci_ctx* ctx, const b2ci_dets* dets, int nparts, int64_t nsamples,
                                 int64_t* offsets) {
  B2_TRY_CTX(ctx)
  if (!dets) throw Error("b2ci_dets_balanced_partition: null determinant list");
  dets_balanced_partition(ctx, dets, nparts, nsamples, offsets);
  return 0;
  B2_CATCH
}
int b2ci_csr_diagonal(b2ci_ctx* ctx, const b2ci_csr* m, double* D) {
  B2_TRY_CTX(ctx)
  DevBuf<double> d(m->nrows > 0 ? m->nrows : 1);
  csr_diagonal_dev(ctx, m, d);
  B2_CUDA(cudaMemcpyAsync(D, d, size_t(m->nrows) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  return 0;
  B2_CATCH
}

int b2ci_davidson(b2ci_ctx* ctx, const b2ci_csr* m, int64_t max_m, double tol, double* X,
                  int use_guess_policy, int64_t* niter, double* eigval, double* trace) {
  B2_TRY_CTX(ctx)
  int64_t it = 0;
  double ev = 0.;
  const int rc = davidson(ctx, m, max_m, tol, X, use_guess_policy, &it, &ev, trace);
  if (niter) *niter = it;
  if (eigval) *eigval = ev;
  return rc;
  B2_CATCH
}

int b2ci_dense_ground_state(b2ci_ctx* ctx, const b2ci_csr* m, double* eigval, double* eigvec) {
  B2_TRY_CTX(ctx)
  if (!m || !eigval || !eigvec) throw Error("b2ci_dense_ground_state: null argument");
  dense_ground_state(ctx, m, eigval, eigvec);
  return 0;
  B2_CATCH
}

double b2ci_timer_ms(const b2ci_ctx* ctx, const char* name) {
  auto it = ctx->timers.find(name);
  return it == ctx->timers.end() ? -1.0 : it->second;
}

int b2ci_asci_search(b2ci_ctx* ctx, const b2ci_asci_search_opts* opts, const uint64_t* core_words,
                     int wpd, const double* core_coeffs, int64_t ncdets, double E0,
                     uint64_t* out_words, int64_t cap, int64_t* n_out, double* stats) {
  B2_TRY_CTX(ctx)
  return asci_search(ctx, opts, core_words, wpd, core_coeffs, ncdets, E0, out_words, cap, n_out, stats,
                     nullptr, nullptr, nullptr, nullptr, false);
  B2_CATCH
}
int b2ci_asci_candidates(b2ci_ctx* ctx, const b2ci_asci_search_opts* opts, const uint64_t* core_words,
                         int wpd, const double* core_coeffs, int64_t ncdets, double E0,
                         uint64_t* out_words, double* out_cmatel, double* out_hdiag, int64_t* n_out) {
  B2_TRY_CTX(ctx)
  return asci_search(ctx, opts, core_words, wpd, core_coeffs, ncdets, E0, nullptr, 0, nullptr, nullptr,
                     out_words, out_cmatel, out_hdiag, n_out, true);
  B2_CATCH
}

int b2ci_asci_pt2(b2ci_ctx* ctx, const uint64_t* det_words, int wpd, const double* coeffs,
                  int64_t ndets, double E_asci, double pt2_tol, double* ept2, int64_t* npt2) {
  B2_TRY_CTX(ctx)
  if (!ept2) throw Error("b2ci_asci_pt2: ept2 is NULL");
  b2ci_asci_search_opts o;
  o.ndets_max = ndets; o.h_el_tol = pt2_tol; o.rv_prune_tol = 0.; o.just_singles = 0; o.sort_output = 0;
  double acc[2] = {0., 0.};
  asci_search(ctx, &o, det_words, wpd, coeffs, ndets, E_asci, nullptr, 0, nullptr, nullptr, nullptr, nullptr,
              nullptr, nullptr, false, acc);
  *ept2 = acc[0];
  if (npt2) *npt2 = int64_t(acc[1]);
  return 0;
  B2_CATCH
}

int b2ci_form_rdms(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, double* ordm, double* trdm) {
  B2_TRY_CTX(ctx)
  form_rdms(ctx, dets, C, false, ordm, nullptr, trdm, nullptr, nullptr);
  return 0;
  B2_CATCH
}
int b2ci_form_rdms_spin_dep(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, double* ordm_aa,
                            double* ordm_bb, double* trdm_aaaa, double* trdm_bbbb, double* trdm_aabb) {
  B2_TRY_CTX(ctx)
  form_rdms(ctx, dets, C, true, ordm_aa, ordm_bb, trdm_aaaa, trdm_bbbb, trdm_aabb);
  return 0;
  B2_CATCH
}

int b2ci_form_entropies(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, double* s1, double* s2, double* mi) {
  B2_TRY_CTX(ctx)
  if (!s1) throw Error("b2ci_form_entropies: single_orbital_entropies output is NULL");
  const bool need_s2 = s2 || mi;
  std::vector<double> I(entropy_intermediate_doubles(ctx->norb, need_s2));
  entropy_intermediates(ctx, dets, C, need_s2, I.data());
  host_entropies_from_intermediates(ctx->norb, need_s2, I.data(), s1, s2, mi);
  return 0;
  B2_CATCH
}
int64_t b2ci_entropy_intermediate_count(int norb, int need_s2) {
  return int64_t(entropy_intermediate_doubles(norb, need_s2 != 0));
}
int b2ci_entropy_intermediates(b2ci_ctx* ctx, const b2ci_dets* dets, const double* C, int need_s2, double* out) {
  B2_TRY_CTX(ctx)
  entropy_intermediates(ctx, dets, C, need_s2 != 0, out);
  return 0;
  B2_CATCH
}
int b2ci_host_entropies_from_intermediates(int norb, int need_s2, const double* intermediates, double* s1,
                                           double* s2, double* mi) {
  B2_TRY
  if (norb < 1 || !intermediates || !s1) throw Error("b2ci_host_entropies_from_intermediates: bad arguments");
  host_entropies_from_intermediates(norb, need_s2 != 0, intermediates, s1, s2, mi);
  return 0;
  B2_CATCH
}

double b2ci_host_matrix_element(int norb, const double* T, const double* V, uint64_t bra_alpha,
                                uint64_t bra_beta, uint64_t ket_alpha, uint64_t ket_beta) {
  // host evaluation of the same __host__ __device__ code path the kernels use
  const size_t n = norb, n2 = n * n, n3 = n2 * n;
  std::vector<double> buf(ints_total_doubles(norb));
  IntsView I = make_view(norb, buf.data());
  memcpy(buf.data(), T, n2 * 8);
  memcpy(const_cast<double*>(I.V), V, n2 * n2 * 8);
  double* G = const_cast<double*>(I.G);
  double* Vr = const_cast<double*>(I.Vr);
  double* G2 = const_cast<double*>(I.G2);
  double* V2 = const_cast<double*>(I.V2);
  for (size_t j = 0; j < n; ++j)
    for (size_t i = 0; i < n; ++i)
      for (size_t k = 0; k < n; ++k) {
        G[k + i * n + j * n2] = V[k + k * n + i * n2 + j * n3] - V[k + j * n + i * n2 + k * n3];
        Vr[k + i * n + j * n2] = V[k + k * n + i * n2 + j * n3];
      }
  for (size_t j = 0; j < n; ++j)
    for (size_t i = 0; i < n; ++i) {
      G2[i + j * n] = 0.5 * (V[i + i * n + j * n2 + j * n3] - V[i + j * n + j * n2 + i * n3]);
      V2[i + j * n] = V[i + i * n + j * n2 + j * n3];
    }
  return matel(I, bra_alpha, bra_beta, ket_alpha, ket_beta);
}

int b2ci_host_sym_eig_lowest(int n, const double* A, int lda, double* lambda, double* vec) {
  B2_TRY
  if (n < 1 || !A || !lambda || !vec || lda < n) throw Error("b2ci_host_sym_eig_lowest: bad arguments");
  sym_eig_lowest(n, A, lda, lambda, vec);
  return 0;
  B2_CATCH
}

int b2ci_host_sym_eig_lower(int n, double* A, int lda, double* W) {
  B2_TRY
  sym_eig_lower(n, A, lda, W);
  return 0;
  B2_CATCH
}

}  // extern "C"
