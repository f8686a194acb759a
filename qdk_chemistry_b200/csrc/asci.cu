// ASCI connected-determinant search on the device.
//
// Replaces macis::asci_search / asci_contributions_constraint (external/macis/include/macis/
// asci/determinant_search.hpp:349-771, 808-1123) and the contribution emitters
// (asci/determinant_contributions.hpp:93-297 == asci/mask_constraints.hpp:400-670 term by
// term). The reference partitions the excited space by alpha-string constraints so every
// CPU thread can sort/accumulate its share; on the GPU that partition is unnecessary:
//
//   1. per core determinant: orbital energies (fast_diagonals.ipp:29-49) and <D|H|D>
//   2. count pass / fill pass: one CTA per core determinant enumerates its singles, same-spin
//      doubles and opposite-spin doubles, keeps |c*h| >= h_el_tol, and writes
//      (bitstring key, c*h, E0 - <Q|H|Q>) records parent-major (deterministic offsets)
//   3. stable LSD radix sort of (key, record index)              [radix.cu]
//   4. per unique key: sequential sum of c*h in parent order, h_diag of the first parent
//      (accumulate_asci_pairs, determinant_sort.hpp:115-136, in canonical order)
//   5. drop core determinants (rv = inf sentinel, determinant_search.hpp:659-660,970-974)
//      and |rv| <= rv_prune_tol (:676-681); radix-select the k-th largest |rv| and keep
//      every candidate >= it (ties retained, :994-1080); append the core determinants.
#include <chrono>
#include <cmath>
#include <cstring>

#include "common.cuh"
#include "slater.cuh"

namespace b2ci {

void radix_sort_pairs(b2ci_ctx* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals,
                      uint32_t* vals_alt, int64_t n, const std::vector<int>& shifts);
void iota_u32(b2ci_ctx* ctx, uint32_t* v, int64_t n);
double select_kth_largest(b2ci_ctx* ctx, const double* score, int64_t n, int64_t k, bool distributed = false);

namespace {

constexpr int GEN_THREADS = 256;
constexpr int MAX_PAIRS = 2016;  // C(64, 2)

__global__ void k_core_pre(IntsView I, const uint64_t* __restrict__ ca,
                           const uint64_t* __restrict__ cb, int64_t nc,
                           double* __restrict__ eps_a, double* __restrict__ eps_b,
                           double* __restrict__ root) {
  const int64_t t = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const int n = I.n;
  if (t >= nc * n) return;
  const int64_t c = t / n;
  const unsigned i = unsigned(t % n);
  const uint64_t a = ca[c], b = cb[c];
  eps_a[t] = orbital_energy(I, i, a, b);
  eps_b[t] = orbital_energy(I, i, b, a);
  if (i == 0) root[c] = me_diag(I, a, b);
}

struct GenArgs {
  IntsView I;
  const uint64_t* ca;
  const uint64_t* cb;
  const double* coeff;
  const double* eps_a;
  const double* eps_b;
  const double* root;
  double E0, tol;
  int just_singles;
  int packed;              // 1: key = beta << 32 | alpha (norb <= 32); 0: key = alpha, key2 = beta
  uint32_t nparts, part;   // key partition handled by this launch (nparts == 1: everything)
  int32_t* count;          // count pass
  const int64_t* base;     // fill pass
  uint64_t* key;
  uint64_t* key2;
  double* cm;
  double* hd;
};

// partition of the excited space by a hash of the determinant key: every contribution to one
// determinant lands in the same part, in parent order, so per-part sort + accumulate gives the
// same sums as one global pass (the device analogue of the reference's alpha-string constraints,
// asci/mask_constraints.hpp: each Q determinant belongs to exactly one constraint)
__host__ __device__ __forceinline__ uint32_t key_part(uint64_t key, uint32_t nparts) {
  uint64_t x = key * 0x9E3779B97F4A7C15ull;
  x ^= x >> 29;
  x *= 0xBF58476D1CE4E5B9ull;
  return uint32_t((x >> 33) % nparts);
}
__device__ __forceinline__ void pair_from_index(int p, int m, int& ii, int& jj) {
  // p-th pair (ii < jj) of m items in the order of the reference's nested loops
  int i = 0, rem = p;
  while (rem >= m - 1 - i) { rem -= m - 1 - i; ++i; }
  ii = i;
  jj = i + 1 + rem;
}

#define G2_(p, q) ldg(A.I.G2 + (p) + size_t(q) * n)
#define V2_(p, q) ldg(A.I.V2 + (p) + size_t(q) * n)

template <bool FILL>
__global__ void __launch_bounds__(GEN_THREADS)
k_generate(const GenArgs A) {
  __shared__ unsigned char occ_a[64], vir_a[64], occ_b[64], vir_b[64];
  __shared__ unsigned short po_a[MAX_PAIRS], pv_a[MAX_PAIRS], po_b[MAX_PAIRS], pv_b[MAX_PAIRS];
  __shared__ int s_cnt;
  __shared__ int s_red[GEN_THREADS / 32];
  const int64_t c = blockIdx.x;
  const size_t n = A.I.n, n2 = n * n;
  const uint64_t sa = A.ca[c], sb = A.cb[c];
  const int na = __popcll(sa), nb = __popcll(sb);
  const int nva = int(n) - na, nvb = int(n) - nb;
  if (threadIdx.x == 0) {
    int k = 0;
    for (uint64_t s = sa; s; s &= s - 1) occ_a[k++] = (unsigned char)lsb64(s);
    k = 0;
    for (uint64_t s = ~sa & low_mask(int(n)); s; s &= s - 1) vir_a[k++] = (unsigned char)lsb64(s);
    k = 0;
    for (uint64_t s = sb; s; s &= s - 1) occ_b[k++] = (unsigned char)lsb64(s);
    k = 0;
    for (uint64_t s = ~sb & low_mask(int(n)); s; s &= s - 1) vir_b[k++] = (unsigned char)lsb64(s);
    s_cnt = 0;
  }
  const int npo_a = na * (na - 1) / 2, npv_a = nva * (nva - 1) / 2;
  const int npo_b = nb * (nb - 1) / 2, npv_b = nvb * (nvb - 1) / 2;
  if (!A.just_singles) {
    for (int p = threadIdx.x; p < npo_a; p += GEN_THREADS) { int i, j; pair_from_index(p, na, i, j); po_a[p] = (unsigned short)((i << 8) | j); }
    for (int p = threadIdx.x; p < npv_a; p += GEN_THREADS) { int i, j; pair_from_index(p, nva, i, j); pv_a[p] = (unsigned short)((i << 8) | j); }
    for (int p = threadIdx.x; p < npo_b; p += GEN_THREADS) { int i, j; pair_from_index(p, nb, i, j); po_b[p] = (unsigned short)((i << 8) | j); }
    for (int p = threadIdx.x; p < npv_b; p += GEN_THREADS) { int i, j; pair_from_index(p, nvb, i, j); pv_b[p] = (unsigned short)((i << 8) | j); }
  }
  __syncthreads();

  const double coeff = A.coeff[c];
  const double root = A.root[c];
  const double* eps_a = A.eps_a + c * n;
  const double* eps_b = A.eps_b + c * n;
  const int64_t nSa = int64_t(na) * nva, nSb = int64_t(nb) * nvb;
  const int64_t nDa = A.just_singles ? 0 : int64_t(npo_a) * npv_a;
  const int64_t nDb = A.just_singles ? 0 : int64_t(npo_b) * npv_b;
  const int64_t nOS = A.just_singles ? 0 : nSa * nSb;
  const int64_t o1 = nSa, o2 = o1 + nSb, o3 = o2 + nDa, o4 = o3 + nDb, o5 = o4 + nOS;
  const int64_t total = o5 + 1;  // + the "no excitation" sentinel
  const int64_t base = FILL ? A.base[c] : 0;
  int mycount = 0;

  for (int64_t t = threadIdx.x; t < total; t += GEN_THREADS) {
    bool emit = false;
    uint64_t ea = sa, eb = sb;
    double cmv = 0., hdv = 0.;
    if (t < o2) {
      // ---- single excitation, alpha (t < o1) or beta
      const bool is_a = t < o1;
      const int64_t u = is_a ? t : t - o1;
      const int nv = is_a ? nva : nvb;
      const unsigned i = is_a ? occ_a[u / nv] : occ_b[u / nv];
      const unsigned a = is_a ? vir_a[u % nv] : vir_b[u % nv];
      const uint64_t same = is_a ? sa : sb, othr = is_a ? sb : sa;
      const double* eps = is_a ? eps_a : eps_b;
      double h_el = ldg(A.I.T + a + i * n);
      const double* G = A.I.G + a * n + i * n2;
      const double* Vr = A.I.Vr + a * n + i * n2;
      for (uint64_t s = same; s; s &= s - 1) h_el += ldg(G + lsb64(s));
      for (uint64_t s = othr; s; s &= s - 1) h_el += ldg(Vr + lsb64(s));
      if (!(fabs(coeff * h_el) < A.tol)) {
        const double sign = sx_sign(same, a, i);
        h_el *= sign;
        const double h_diag = root + eps[a] - eps[i] - G2_(a, i) - G2_(i, a);
        const uint64_t ex = same ^ (uint64_t(1) << i) ^ (uint64_t(1) << a);
        if (is_a) ea = ex; else eb = ex;
        cmv = coeff * h_el;
        hdv = A.E0 - h_diag;
        emit = true;
      }
    } else if (t < o4) {
      // ---- same-spin double, alpha (t < o3) or beta
      const bool is_a = t < o3;
      const int64_t u = is_a ? t - o2 : t - o3;
      const int npv = is_a ? npv_a : npv_b;
      const unsigned short po = is_a ? po_a[u / npv] : po_b[u / npv];
      const unsigned short pv = is_a ? pv_a[u % npv] : pv_b[u % npv];
      const unsigned i = is_a ? occ_a[po >> 8] : occ_b[po >> 8];
      const unsigned j = is_a ? occ_a[po & 0xFF] : occ_b[po & 0xFF];
      const unsigned a = is_a ? vir_a[pv >> 8] : vir_b[pv >> 8];
      const unsigned b = is_a ? vir_a[pv & 0xFF] : vir_b[pv & 0xFF];
      const uint64_t same = is_a ? sa : sb;
      const double* eps = is_a ? eps_a : eps_b;
      const double V_aibj = ldg(A.I.V + (a + i * n) * n2 + (b + j * n));
      const double V_ajbi = ldg(A.I.V + (a + j * n) * n2 + (b + i * n));
      const double G_aibj = V_aibj - V_ajbi;
      if (!(fabs(coeff * G_aibj) < A.tol)) {
        const uint64_t full_ex = (uint64_t(1) << i) | (uint64_t(1) << j) | (uint64_t(1) << a) | (uint64_t(1) << b);
        const uint64_t ex_spin = same ^ full_ex;
        unsigned x1, y1, x2, y2;
        double sign;
        dx_sign_indices(same, ex_spin, full_ex, x1, y1, x2, y2, sign);
        const double h_el = sign * G_aibj;
        const double h_diag = root + eps[a] + eps[b] - eps[i] - eps[j] + G2_(i, j) + G2_(j, i) +
                              G2_(a, b) + G2_(b, a) - G2_(a, i) - G2_(i, a) - G2_(b, i) - G2_(i, b) -
                              G2_(a, j) - G2_(j, a) - G2_(b, j) - G2_(j, b);
        if (is_a) ea = ex_spin; else eb = ex_spin;
        cmv = coeff * h_el;
        hdv = A.E0 - h_diag;
        emit = true;
      }
    } else if (t < o5) {
      // ---- opposite-spin double
      const int64_t u = t - o4;
      const int64_t ua = u / nSb, ub = u % nSb;
      const unsigned i = occ_a[ua / nva], a = vir_a[ua % nva];
      const unsigned j = occ_b[ub / nvb], b = vir_b[ub % nvb];
      const double V_aibj = ldg(A.I.V + a + i * n + (b + j * n) * n2);
      if (!(fabs(coeff * V_aibj) < A.tol)) {
        const double sign_a = sx_sign(sa, a, i);
        const double sign_b = sx_sign(sb, b, j);
        const double sign = sign_a * sign_b;
        const double h_el = sign * V_aibj;
        const double h_diag = root + eps_a[a] + eps_b[b] - eps_a[i] - eps_b[j] + V2_(i, j) + V2_(a, b) -
                              G2_(a, i) - G2_(i, a) - G2_(b, j) - G2_(j, b) - V2_(a, j) - V2_(i, b);
        ea = sa ^ (uint64_t(1) << i) ^ (uint64_t(1) << a);
        eb = sb ^ (uint64_t(1) << j) ^ (uint64_t(1) << b);
        cmv = coeff * h_el;
        hdv = A.E0 - h_diag;
        emit = true;
      }
    } else {
      // ---- the core determinant itself: infinite score marks it for removal
      cmv = INFINITY;
      hdv = 1.0;
      emit = true;
    }
    if (emit && A.nparts > 1)
      emit = key_part(A.packed ? ((eb << 32) | ea) : (ea ^ (eb * 0xD6E8FEB86659FD93ull)), A.nparts) == A.part;
    if (emit) {
      if (FILL) {
        const int64_t pos = base + atomicAdd(&s_cnt, 1);
        if (A.packed) A.key[pos] = (eb << 32) | ea;
        else { A.key[pos] = ea; A.key2[pos] = eb; }
        A.cm[pos] = cmv;
        A.hd[pos] = hdv;
      } else {
        ++mycount;
      }
    }
  }
  if (!FILL) {
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) mycount += __shfl_down_sync(0xffffffffu, mycount, d);
    if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = mycount;
    __syncthreads();
    if (threadIdx.x == 0) {
      int s = 0;
      for (int k = 0; k < GEN_THREADS / 32; ++k) s += s_red[k];
      A.count[c] = s;
    }
  }
}
#undef G2_
#undef V2_

__global__ void k_gather_u64(const uint64_t* __restrict__ src, const uint32_t* __restrict__ idx,
                             int64_t n, uint64_t* __restrict__ dst) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[idx[i]];
}
// head flags of equal-key segments (two-word keys supported)
__global__ void k_seg_flags(const uint64_t* __restrict__ k1, const uint64_t* __restrict__ k2,
                            int64_t n, int32_t* __restrict__ flag) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  bool head = (i == 0) || (k1[i] != k1[i - 1]);
  if (!head && k2) head = k2[i] != k2[i - 1];
  flag[i] = head ? 1 : 0;
}
// one thread per segment head: sequential (parent-ordered) accumulation
__global__ void k_seg_accumulate(const uint64_t* __restrict__ k1, const uint64_t* __restrict__ k2,
                                 const uint32_t* __restrict__ idx, const int32_t* __restrict__ flag,
                                 const int64_t* __restrict__ seg_of, int64_t n,
                                 const double* __restrict__ cm, const double* __restrict__ hd,
                                 uint64_t* __restrict__ sk1, uint64_t* __restrict__ sk2,
                                 double* __restrict__ scm, double* __restrict__ shd) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || !flag[i]) return;
  const int64_t s = seg_of[i];
  double acc = cm[idx[i]];
  for (int64_t j = i + 1; j < n && !flag[j]; ++j) acc += cm[idx[j]];
  sk1[s] = k1[i];
  if (sk2) sk2[s] = k2[i];
  scm[s] = acc;
  shd[s] = hd[idx[i]];
}
// candidate filter: finite rv and |rv| > prune tol
__global__ void k_score(const double* __restrict__ scm, const double* __restrict__ shd, int64_t n,
                        double prune, int32_t* __restrict__ keep, double* __restrict__ score) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const double rv = scm[i] / shd[i];
  const double a = fabs(rv);
  const bool k = (a > prune) && !isinf(rv);
  keep[i] = k ? 1 : 0;
  score[i] = k ? a : 0.0;
}
__global__ void k_compact(const int32_t* __restrict__ keep, const int64_t* __restrict__ pos, int64_t n,
                          const uint64_t* __restrict__ sk1, const uint64_t* __restrict__ sk2,
                          const double* __restrict__ score, uint64_t* __restrict__ ok1,
                          uint64_t* __restrict__ ok2, double* __restrict__ oscore) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n || !keep[i]) return;
  const int64_t p = pos[i];
  ok1[p] = sk1[i];
  if (ok2) ok2[p] = sk2[i];
  if (oscore) oscore[p] = score[i];
}
__global__ void k_keep_ge(const double* __restrict__ score, int64_t n, double kth,
                          int32_t* __restrict__ keep) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) keep[i] = score[i] >= kth ? 1 : 0;
}
__global__ void k_max_below(const double* __restrict__ score, int64_t n, double kth,
                            unsigned long long* __restrict__ out) {
  const int64_t stride = int64_t(gridDim.x) * blockDim.x;
  unsigned long long best = 0;
  for (int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const double s = score[i];
    if (s < kth) {
      const unsigned long long b = (unsigned long long)__double_as_longlong(s);
      best = b > best ? b : best;
    }
  }
  atomicMax(out, best);
}

// ASCI-PT2 (asci/pt2.hpp:399-410): sum of rv * c_times_matel = (sum c*h)^2 / (E0 - <Q|H|Q>) over the
// accumulated candidates whose c*h is finite (the core determinants carry the inf sentinel).
// Fixed grid, fixed tree: the partial sums are combined on the host in block order.
constexpr int PT2_BLOCKS = 592, PT2_THREADS = 256;
__global__ void __launch_bounds__(PT2_THREADS)
k_pt2_partial(const double* __restrict__ scm, const double* __restrict__ shd, int64_t n,
              double* __restrict__ part_sum, double* __restrict__ part_cnt) {
  __shared__ double ssum[PT2_THREADS], scnt[PT2_THREADS];
  double acc = 0., cnt = 0.;
  for (int64_t i = int64_t(blockIdx.x) * PT2_THREADS + threadIdx.x; i < n; i += int64_t(PT2_BLOCKS) * PT2_THREADS) {
    const double c = scm[i];
    if (!isinf(c)) { acc += (c / shd[i]) * c; cnt += 1.; }
  }
  ssum[threadIdx.x] = acc; scnt[threadIdx.x] = cnt;
  __syncthreads();
  for (int s = PT2_THREADS / 2; s > 0; s >>= 1) {
    if (threadIdx.x < s) { ssum[threadIdx.x] += ssum[threadIdx.x + s]; scnt[threadIdx.x] += scnt[threadIdx.x + s]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { part_sum[blockIdx.x] = ssum[0]; part_cnt[blockIdx.x] = scnt[0]; }
}

// keys of the core determinants behind the selected ones (packed: beta << 32 | alpha)
__global__ void k_append_core_keys(const uint64_t* __restrict__ ca, const uint64_t* __restrict__ cb, int64_t nc,
                                   int packed, uint64_t* __restrict__ k1, uint64_t* __restrict__ k2) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nc) return;
  if (packed) k1[i] = (cb[i] << 32) | ca[i];
  else { k1[i] = ca[i]; k2[i] = cb[i]; }
}
// all-gathered fixed-size slabs -> one list: rank r's first (prefix[r+1] - prefix[r]) entries, rank-major
__global__ void k_slab_compact(const uint64_t* __restrict__ g, int64_t slab, const int64_t* __restrict__ prefix, int nranks,
                               uint64_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= slab * nranks) return;
  const int64_t r = i / slab, j = i - r * slab;
  if (j < prefix[r + 1] - prefix[r]) out[prefix[r] + j] = g[i];
}
__global__ void k_fill_u64(uint64_t* p, int64_t n, uint64_t v) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}
unsigned grid1d(int64_t n, int threads = 256) { return unsigned((n + threads - 1) / threads); }

}  // namespace

int asci_search(b2ci_ctx* ctx, const b2ci_asci_search_opts* o, const uint64_t* core_words, int wpd,
                const double* coeffs, int64_t nc, double E0, uint64_t* out_words, int64_t cap,
                int64_t* n_out, double* stats, uint64_t* cand_words, double* cand_cm,
                double* cand_hd, int64_t* cand_n, bool candidates_only, double* pt2_out) {
  if (!ctx->ints_dev) throw Error("b2ci_asci_search: integrals not uploaded");
  if (!o || !core_words || !coeffs || nc < 1) throw Error("b2ci_asci_search: bad arguments");
  if (wpd != 1 && wpd != 2) throw Error("b2ci_asci_search: words_per_det must be 1 or 2");
  const int n = ctx->norb;
  if (wpd == 1 && n > 32) throw Error("b2ci_asci_search: wfn_t<64> holds at most 32 orbitals per spin");
  cudaStream_t st = ctx->stream;
  // host wall-clock trace (B2CI_ASCI_TRACE=1): time between marks includes allocations and waits
  const bool trace = getenv("B2CI_ASCI_TRACE") != nullptr;
  auto t_host0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(st);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[asci] %-28s +%9.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_host0).count());
    t_host0 = t;
  };
  auto& T = ctx->timers;
  T["asci_search.PAIR_DUR"] = T["asci_search.SORT_ACC_DUR"] = T["asci_search.TOPK_DUR"] = 0.;

  // the reference insists on spin-sorted core determinants (determinant_search.hpp:363-364)
  for (int64_t i = 1; i < nc; ++i) {
    uint64_t a0, b0, a1, b1;
    if (wpd == 1) { a0 = core_words[i - 1] & 0xFFFFFFFFull; b0 = core_words[i - 1] >> 32; a1 = core_words[i] & 0xFFFFFFFFull; b1 = core_words[i] >> 32; }
    else { a0 = core_words[2 * i - 2]; b0 = core_words[2 * i - 1]; a1 = core_words[2 * i]; b1 = core_words[2 * i + 1]; }
    if (a1 < a0 || (a1 == a0 && b1 < b0)) throw Error("ASCI Search Only Works with Sorted Wfns");
  }

  b2ci_dets core;
  void dets_from_words(b2ci_ctx*, const uint64_t*, int, int64_t, b2ci_dets*);
  dets_from_words(ctx, core_words, wpd, nc, &core);
  DevBuf<uint64_t> ca_own, cb_own;  // adopt for RAII
  ca_own.p = core.alpha; ca_own.n = nc;
  cb_own.p = core.beta; cb_own.n = nc;
  DevBuf<double> dcoeff(nc), eps_a(size_t(nc) * n), eps_b(size_t(nc) * n), root(nc);
  B2_CUDA(cudaMemcpyAsync(dcoeff, coeffs, size_t(nc) * 8, cudaMemcpyHostToDevice, st));

  // norb <= 32: the key is the wfn_t<64> word itself (beta << 32 | alpha), numeric order ==
  // bitset_less. 32 < norb <= 64 (wfn_t<128>): 128-bit keys held as two words (alpha = low word,
  // beta = high word, raw_bitset.hpp:94-106), sorted by two stable LSD phases (alpha digits, then
  // beta digits), which is the numeric order of the 128-bit word (bitset_less, raw_bitset.hpp:143-160).
  if (n > 64) throw Error("b2ci_asci_search: norb > 64 (wfn_t<256>) is not supported");
  const bool two = n > 32;

  GenArgs A;
  A.I = ctx->ints;
  A.ca = core.alpha; A.cb = core.beta; A.coeff = dcoeff;
  A.eps_a = eps_a; A.eps_b = eps_b; A.root = root;
  A.E0 = E0; A.tol = o->h_el_tol; A.just_singles = o->just_singles;
  A.packed = two ? 0 : 1;
  DevBuf<int32_t> count(nc);
  DevBuf<int64_t> base(nc + 1);

  // ---- how many key partitions: the unfiltered contribution count against the memory budget
  int64_t M_total = 0;
  {
    ScopedTimer t(ctx, "asci_search.PAIR_DUR", true);
    k_core_pre<<<grid1d(nc * n), 256, 0, st>>>(ctx->ints, core.alpha, core.beta, nc, eps_a, eps_b, root);
    A.nparts = 1; A.part = 0;
    A.count = count; A.base = nullptr; A.key = nullptr; A.key2 = nullptr; A.cm = nullptr; A.hd = nullptr;
    k_generate<false><<<unsigned(nc), GEN_THREADS, 0, st>>>(A);
    ctx->launches += 2;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, count, base, nc);
    B2_CUDA(cudaMemcpyAsync(&M_total, base.p + nc, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
  }
  mark("core upload + count pass");
  size_t free_b = 0, total_b = 0;
  B2_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const int64_t bytes_per = two ? 108 : 80;  // workspace per contribution (records, sort buffers, flags, segments)
  int64_t budget = std::min<int64_t>(int64_t(double(free_b) * 0.5) / bytes_per, (int64_t(1) << 31) - 1);
  if (const char* env = getenv("B2CI_ASCI_BUDGET")) budget = std::max<int64_t>(1024, atoll(env));
  // hash partitions are even to a few percent; 1.25 covers the imbalance
  int64_t nparts = std::max<int64_t>(1, (int64_t(double(M_total) * 1.25) + budget - 1) / budget);
  if (M_total <= budget) nparts = 1;
  if (const char* env = getenv("B2CI_ASCI_PARTS")) nparts = std::max<int64_t>(1, atoll(env));
  const int nranks = ctx->nranks, rank = ctx->rank;
  if (nranks > 1) {
    // the partition count decides which rank owns which key: it must be ONE number. Free memory (and
    // with it the budget) differs from rank to rank, so the ranks agree on the largest request and
    // check that they are searching from the same core set.
    std::vector<int64_t> all_parts, all_m;
    comm_allgather_i64_host(ctx, nparts, all_parts);
    comm_allgather_i64_host(ctx, M_total, all_m);
    for (int r = 0; r < nranks; ++r) {
      nparts = std::max(nparts, all_parts[r]);
      if (all_m[r] != M_total)
        throw Error("b2ci_asci_search: ranks disagree on the contribution count (" + std::to_string(M_total) + " here, " +
                    std::to_string(all_m[r]) + " on rank " + std::to_string(r) + "): different core sets or integrals");
    }
    nparts = ((std::max<int64_t>(nparts, nranks) + nranks - 1) / nranks) * nranks;
  }
  if (candidates_only && (nparts > 1 || nranks > 1))
    throw Error("b2ci_asci_candidates: the candidate table is only available for single-part searches");
  T["asci_search.nparts"] = double(nparts);

  // candidates of this rank that survive pruning: (key, |rv|), appended part by part
  DevBuf<uint64_t> cand_key, cand_key2;  // cand_key2: beta words of two-word keys
  DevBuf<double> cand_score;
  int64_t ncand = 0, cand_cap = 0;
  int64_t M_sum = 0, nseg_sum = 0;
  if (pt2_out) pt2_out[0] = pt2_out[1] = 0.;
  auto append_candidates = [&](const uint64_t* k, const uint64_t* k2, const double* sc, int64_t m) {
    if (ncand + m > cand_cap) {
      const int64_t ncap = std::max<int64_t>(ncand + m, cand_cap * 2);
      DevBuf<uint64_t> nk(ncap), nk2(two ? ncap : 1);
      DevBuf<double> ns(ncap);
      if (ncand) {
        B2_CUDA(cudaMemcpyAsync(nk, cand_key, size_t(ncand) * 8, cudaMemcpyDeviceToDevice, st));
        if (two) B2_CUDA(cudaMemcpyAsync(nk2, cand_key2, size_t(ncand) * 8, cudaMemcpyDeviceToDevice, st));
        B2_CUDA(cudaMemcpyAsync(ns, cand_score, size_t(ncand) * 8, cudaMemcpyDeviceToDevice, st));
      }
      cand_key = std::move(nk);
      cand_key2 = std::move(nk2);
      cand_score = std::move(ns);
      cand_cap = ncap;
    }
    if (m) {
      B2_CUDA(cudaMemcpyAsync(cand_key.p + ncand, k, size_t(m) * 8, cudaMemcpyDeviceToDevice, st));
      if (two) B2_CUDA(cudaMemcpyAsync(cand_key2.p + ncand, k2, size_t(m) * 8, cudaMemcpyDeviceToDevice, st));
      B2_CUDA(cudaMemcpyAsync(cand_score.p + ncand, sc, size_t(m) * 8, cudaMemcpyDeviceToDevice, st));
    }
    ncand += m;
  };

  for (int64_t part = rank; part < nparts; part += nranks) {
    int64_t M = 0;
    uint64_t *key = nullptr, *key2 = nullptr;
    double *cm = nullptr, *hd = nullptr;
    {
      ScopedTimer t(ctx, "asci_search.PAIR_DUR", true);
      A.nparts = uint32_t(nparts); A.part = uint32_t(part);
      if (nparts > 1) {
        A.count = count; A.base = nullptr; A.key = nullptr; A.cm = nullptr; A.hd = nullptr;
        k_generate<false><<<unsigned(nc), GEN_THREADS, 0, st>>>(A);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        exclusive_scan_i32_to_i64(ctx, count, base, nc);
      }
      B2_CUDA(cudaMemcpyAsync(&M, base.p + nc, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      if (M >= (int64_t(1) << 32)) throw Error("b2ci_asci_search: more than 2^32 contributions in one part");
      mark("part count");
      // workspace of this part: records (24 B), sort double buffers and indices (16 B), segment
      // flags and ids (12 B), accumulated segments (<= 24 B), 256-byte alignment slack; two-word
      // keys add the beta word of the records (8 B), a sort copy (8 B) and of the segments (8 B)
      const size_t Mz = size_t(M > 0 ? M : 1);
      const size_t need = Mz * (two ? 100 : 76) + 16 * 1024;
      if (need > ctx->arena_cap) {
        B2_CUDA(cudaMemGetInfo(&free_b, &total_b));
        if (need > free_b + ctx->arena_cap)
          throw Error("b2ci_asci_search: " + std::to_string(M) + " contributions need " +
                      std::to_string(need >> 20) + " MiB of device memory, " +
                      std::to_string((free_b + ctx->arena_cap) >> 20) + " MiB free");
      }
      arena_reserve(ctx, need);
      arena_reset(ctx);
      key = arena_take<uint64_t>(ctx, Mz);
      if (two) key2 = arena_take<uint64_t>(ctx, Mz);
      cm = arena_take<double>(ctx, Mz);
      hd = arena_take<double>(ctx, Mz);
      mark("alloc key/cm/hd");
      A.count = nullptr; A.base = base; A.key = key; A.key2 = key2; A.cm = cm; A.hd = hd;
      k_generate<true><<<unsigned(nc), GEN_THREADS, 0, st>>>(A);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      mark("generate");
    }
    M_sum += M;
    if (M == 0) continue;

    // ---- sort + accumulate
    int64_t nseg = 0;
    uint64_t *sk1 = nullptr, *sk2 = nullptr;
    double *scm = nullptr, *shd = nullptr;
    {
      ScopedTimer t(ctx, "asci_search.SORT_ACC_DUR", true);
      uint64_t* kalt = arena_take<uint64_t>(ctx, M);
      uint64_t* kwork = two ? arena_take<uint64_t>(ctx, M) : nullptr;
      uint32_t* idx = arena_take<uint32_t>(ctx, M);
      uint32_t* idx_alt = arena_take<uint32_t>(ctx, M);
      mark("alloc sort buffers");
      iota_u32(ctx, idx, M);
      const int ndig = (n + 7) / 8;
      std::vector<int> shifts;
      const uint64_t *k1s = key, *k2s = nullptr;  // the sorted key words
      if (!two) {
        for (int d = 0; d < ndig; ++d) shifts.push_back(8 * d);        // alpha bits
        for (int d = 0; d < ndig; ++d) shifts.push_back(32 + 8 * d);   // beta bits (high half)
        radix_sort_pairs(ctx, key, kalt, idx, idx_alt, M, shifts);
      } else {
        for (int d = 0; d < ndig; ++d) shifts.push_back(8 * d);
        // phase 1: by the alpha word (a copy: the records keep their order for the final gather)
        B2_CUDA(cudaMemcpyAsync(kwork, key, size_t(M) * 8, cudaMemcpyDeviceToDevice, st));
        radix_sort_pairs(ctx, kwork, kalt, idx, idx_alt, M, shifts);
        // phase 2: by the beta word, stable on top of phase 1
        k_gather_u64<<<grid1d(M), 256, 0, st>>>(key2, idx, M, kwork);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        radix_sort_pairs(ctx, kwork, kalt, idx, idx_alt, M, shifts);
        k_gather_u64<<<grid1d(M), 256, 0, st>>>(key, idx, M, kalt);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        k1s = kalt;
        k2s = kwork;
      }
      mark("radix sort");
      int32_t* flag = arena_take<int32_t>(ctx, M);
      int64_t* seg_of = arena_take<int64_t>(ctx, M + 1);
      k_seg_flags<<<grid1d(M), 256, 0, st>>>(k1s, k2s, M, flag);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, flag, seg_of, M);
      B2_CUDA(cudaMemcpyAsync(&nseg, seg_of + M, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      mark("segment flags + scan");
      sk1 = arena_take<uint64_t>(ctx, nseg);
      if (two) sk2 = arena_take<uint64_t>(ctx, nseg);
      scm = arena_take<double>(ctx, nseg);
      shd = arena_take<double>(ctx, nseg);
      // seg_of[i] (exclusive scan) is the segment id of a head at i
      k_seg_accumulate<<<grid1d(M), 256, 0, st>>>(k1s, k2s, idx, flag, seg_of, M, cm, hd, sk1, sk2, scm, shd);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    mark("accumulate");
    nseg_sum += nseg;

    if (pt2_out) {
      ScopedTimer t(ctx, "asci_search.TOPK_DUR", true);
      DevBuf<double> ps(PT2_BLOCKS), pc(PT2_BLOCKS);
      k_pt2_partial<<<PT2_BLOCKS, PT2_THREADS, 0, st>>>(scm, shd, nseg, ps, pc);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      std::vector<double> hs(PT2_BLOCKS), hc(PT2_BLOCKS);
      B2_CUDA(cudaMemcpyAsync(hs.data(), ps, PT2_BLOCKS * 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(hc.data(), pc, PT2_BLOCKS * 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      for (int b = 0; b < PT2_BLOCKS; ++b) { pt2_out[0] += hs[b]; pt2_out[1] += hc[b]; }
      continue;
    }
    if (candidates_only) {
      if (cand_n) *cand_n = nseg;
      if (cand_words) {
        std::vector<uint64_t> tmp(nseg), tmp2(two ? nseg : 0);
        B2_CUDA(cudaMemcpyAsync(tmp.data(), sk1, size_t(nseg) * 8, cudaMemcpyDeviceToHost, st));
        if (two) B2_CUDA(cudaMemcpyAsync(tmp2.data(), sk2, size_t(nseg) * 8, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(cand_cm, scm, size_t(nseg) * 8, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(cand_hd, shd, size_t(nseg) * 8, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        if (two)
          for (int64_t i = 0; i < nseg; ++i) { cand_words[2 * i] = tmp[i]; cand_words[2 * i + 1] = tmp2[i]; }
        else if (wpd == 1) memcpy(cand_words, tmp.data(), size_t(nseg) * 8);
        else
          for (int64_t i = 0; i < nseg; ++i) { cand_words[2 * i] = tmp[i] & 0xFFFFFFFFull; cand_words[2 * i + 1] = tmp[i] >> 32; }
      }
      return 0;
    }

    // ---- prune: finite rv (core determinants carry inf) and |rv| > rv_prune_tol
    {
      ScopedTimer t(ctx, "asci_search.TOPK_DUR", true);
      // records, sort buffers and flags are dead now: their workspace (52 B per contribution,
      // below the accumulated segments) is reused for the 36 B per segment of this phase
      arena_reset(ctx);
      int32_t* keep = arena_take<int32_t>(ctx, nseg);
      double* score = arena_take<double>(ctx, nseg);
      int64_t* pos = arena_take<int64_t>(ctx, nseg + 1);
      int64_t m = 0;
      k_score<<<grid1d(nseg), 256, 0, st>>>(scm, shd, nseg, o->rv_prune_tol, keep, score);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, keep, pos, nseg);
      B2_CUDA(cudaMemcpyAsync(&m, pos + nseg, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      if (m) {
        uint64_t* ck = arena_take<uint64_t>(ctx, m);
        uint64_t* ck2 = two ? arena_take<uint64_t>(ctx, m) : nullptr;
        double* cs = arena_take<double>(ctx, m);
        k_compact<<<grid1d(nseg), 256, 0, st>>>(keep, pos, nseg, sk1, sk2, score, ck, ck2, cs);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        append_candidates(ck, ck2, cs, m);  // survivors leave the workspace
        B2_CUDA(cudaStreamSynchronize(st));
      }
    }
    mark("prune + keep candidates");
  }

  if (pt2_out) {
    if (nranks > 1) {  // every rank summed its own key partitions
      DevBuf<double> d(2);
      B2_CUDA(cudaMemcpyAsync(d, pt2_out, 16, cudaMemcpyHostToDevice, st));
      comm_allreduce_sum(ctx, d, 2);
      B2_CUDA(cudaMemcpyAsync(pt2_out, d, 16, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
    }
    if (stats) { stats[0] = double(M_sum); stats[1] = double(nseg_sum); stats[5] = double(nparts); }
    return 0;
  }

  // ---- top-k over the surviving candidates (determinant_search.hpp:966-1114)
  const int64_t top_k = o->ndets_max - nc;
  // keep every candidate with score >= the k-th largest of `cs` (ties retained); returns count
  auto keep_top = [&](DevBuf<uint64_t>& ck, DevBuf<uint64_t>& ck2, DevBuf<double>& cs, int64_t m, int64_t k,
                      double& kth, double& below, bool want_scores) -> int64_t {
    kth = select_kth_largest(ctx, cs, m, k);
    DevBuf<int32_t> keep2(m);
    DevBuf<int64_t> pos2(m + 1);
    int64_t nk = 0;
    k_keep_ge<<<grid1d(m), 256, 0, st>>>(cs, m, kth, keep2);
    ctx->launches++;
    exclusive_scan_i32_to_i64(ctx, keep2, pos2, m);
    B2_CUDA(cudaMemcpyAsync(&nk, pos2.p + m, 8, cudaMemcpyDeviceToHost, st));
    DevBuf<unsigned long long> mb(1);
    B2_CUDA(cudaMemsetAsync(mb, 0, 8, st));
    k_max_below<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count * 4, (m + 255) / 256)), 256, 0, st>>>(cs, m, kth, mb);
    ctx->launches++;
    unsigned long long mbh = 0;
    B2_CUDA(cudaMemcpyAsync(&mbh, mb, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    memcpy(&below, &mbh, 8);
    DevBuf<uint64_t> sk(nk > 0 ? nk : 1), sk2(two && nk > 0 ? nk : 1);
    DevBuf<double> ss(want_scores && nk > 0 ? nk : 1);
    if (nk) {
      k_compact<<<grid1d(m), 256, 0, st>>>(keep2, pos2, m, ck, two ? ck2.p : nullptr, want_scores ? cs.p : nullptr,
                                          sk, two ? sk2.p : nullptr, want_scores ? ss.p : nullptr);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    B2_CUDA(cudaStreamSynchronize(st));
    ck = std::move(sk);
    ck2 = std::move(sk2);
    if (want_scores) cs = std::move(ss);
    return nk;
  };

  int64_t nkeep = ncand;
  double kth = 0., below = 0.;
  {
    ScopedTimer t(ctx, "asci_search.TOPK_DUR", true);
    const int64_t k_eff = top_k > 0 ? top_k : 1;  // top_k == 0: max_element over an empty range
                                                   // lands on the largest score (:1056-1062)
    if (nranks == 1) {
      if (o->ndets_max >= nc && ncand > top_k) nkeep = keep_top(cand_key, cand_key2, cand_score, ncand, k_eff, kth, below, false);
    } else {
      // distributed selection (determinant_search.hpp:1000-1053, util/dist_quickselect.hpp): the global k-th
      // score by a radix select whose per-digit histograms are summed over the ranks (kilobytes), then every
      // rank keeps its candidates at or above it and only those survivors are exchanged -- about
      // k / nranks keys per rank instead of each rank's local top k with scores.
      std::vector<int64_t> counts;
      comm_allgather_i64_host(ctx, ncand, counts);
      int64_t total_cand = 0;
      for (int64_t c : counts) total_cand += c;
      nkeep = ncand;
      if (o->ndets_max >= nc && total_cand > top_k) {
        kth = select_kth_largest(ctx, cand_score, ncand, k_eff, true);
        int64_t nk = 0;
        DevBuf<int32_t> keep2(ncand > 0 ? ncand : 1);
        DevBuf<int64_t> pos2(ncand + 1);
        DevBuf<unsigned long long> mb(1);
        B2_CUDA(cudaMemsetAsync(mb, 0, 8, st));
        if (ncand) {
          k_keep_ge<<<grid1d(ncand), 256, 0, st>>>(cand_score, ncand, kth, keep2);
          k_max_below<<<(unsigned)std::max<int64_t>(1, std::min<int64_t>(ctx->sm_count * 4, (ncand + 255) / 256)), 256, 0, st>>>(
              cand_score, ncand, kth, mb);
          ctx->launches += 2;
          B2_CHECK_LAUNCH();
        }
        exclusive_scan_i32_to_i64(ctx, keep2, pos2, ncand);
        B2_CUDA(cudaMemcpyAsync(&nk, pos2.p + ncand, 8, cudaMemcpyDeviceToHost, st));
        unsigned long long mbh = 0;
        B2_CUDA(cudaMemcpyAsync(&mbh, mb, 8, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaStreamSynchronize(st));
        DevBuf<uint64_t> sk(nk > 0 ? nk : 1), sk2(two && nk > 0 ? nk : 1);
        if (nk) {
          k_compact<<<grid1d(ncand), 256, 0, st>>>(keep2, pos2, ncand, cand_key, two ? cand_key2.p : nullptr, nullptr, sk,
                                                   two ? sk2.p : nullptr, nullptr);
          ctx->launches++;
          B2_CHECK_LAUNCH();
        }
        cand_key = std::move(sk);
        cand_key2 = std::move(sk2);
        nkeep = nk;
        // largest score below the cut, over all ranks (non-negative doubles order like their bit patterns)
        std::vector<int64_t> allb;
        comm_allgather_i64_host(ctx, int64_t(mbh), allb);
        int64_t mbmax = 0;
        for (int64_t v : allb) mbmax = std::max(mbmax, v);
        memcpy(&below, &mbmax, 8);
      }
      // exchange the survivors' keys: fixed-size slabs (the largest rank's count), padding dropped by count
      comm_allgather_i64_host(ctx, nkeep, counts);
      int64_t slab = 0, total = 0;
      std::vector<int64_t> prefix(size_t(nranks) + 1, 0);
      for (int r = 0; r < nranks; ++r) { slab = std::max(slab, counts[r]); total += counts[r]; prefix[size_t(r) + 1] = total; }
      if (total > 0) {
        DevBuf<uint64_t> sk(slab), gk(size_t(slab) * nranks), outk(total);
        DevBuf<uint64_t> sk2(two ? slab : 1), gk2(two ? size_t(slab) * nranks : 1), outk2(two ? total : 1);
        DevBuf<int64_t> dprefix(size_t(nranks) + 1);
        B2_CUDA(cudaMemcpyAsync(dprefix, prefix.data(), (size_t(nranks) + 1) * 8, cudaMemcpyHostToDevice, st));
        if (nkeep) {
          B2_CUDA(cudaMemcpyAsync(sk, cand_key, size_t(nkeep) * 8, cudaMemcpyDeviceToDevice, st));
          if (two) B2_CUDA(cudaMemcpyAsync(sk2, cand_key2, size_t(nkeep) * 8, cudaMemcpyDeviceToDevice, st));
        }
        comm_allgather_bytes(ctx, sk, gk, size_t(slab) * 8);
        k_slab_compact<<<grid1d(slab * nranks), 256, 0, st>>>(gk, slab, dprefix, nranks, outk);
        ctx->launches++;
        if (two) {
          comm_allgather_bytes(ctx, sk2, gk2, size_t(slab) * 8);
          k_slab_compact<<<grid1d(slab * nranks), 256, 0, st>>>(gk2, slab, dprefix, nranks, outk2);
          ctx->launches++;
        }
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaStreamSynchronize(st));
        cand_key = std::move(outk);
        cand_key2 = std::move(outk2);
      }
      nkeep = total;
    }
    B2_CUDA(cudaStreamSynchronize(st));
  }
  mark("top-k");
  if (stats) {
    stats[0] = double(M_sum); stats[1] = double(nseg_sum); stats[2] = kth; stats[3] = below; stats[4] = double(nkeep);
    stats[5] = double(nparts);
  }
  const int64_t total = nkeep + nc;
  if (n_out) *n_out = total;
  if (total > cap) throw Error("b2ci_asci_search: output capacity " + std::to_string(cap) + " < " + std::to_string(total), 4);
  if (o->sort_output) {
    // spin_comparator order (alpha-major, then beta) of the whole new list, made on the device: what
    // asci_iter does next with a host std::sort (asci/iteration.hpp:117-119). Stable LSD radix: the
    // minor key (beta) first, then the major key (alpha).
    ScopedTimer t(ctx, "asci_search.TOPK_DUR", true);
    DevBuf<uint64_t> k1(total), k1alt(total), k2(two ? total : 1), kw(two ? total : 1);
    DevBuf<uint32_t> idx(total), idx_alt(total);
    if (nkeep) {
      B2_CUDA(cudaMemcpyAsync(k1, cand_key, size_t(nkeep) * 8, cudaMemcpyDeviceToDevice, st));
      if (two) B2_CUDA(cudaMemcpyAsync(k2, cand_key2, size_t(nkeep) * 8, cudaMemcpyDeviceToDevice, st));
    }
    k_append_core_keys<<<grid1d(nc), 256, 0, st>>>(core.alpha, core.beta, nc, two ? 0 : 1, k1.p + nkeep, k2.p + (two ? nkeep : 0));
    ctx->launches++;
    B2_CHECK_LAUNCH();
    iota_u32(ctx, idx, total);
    const int ndig = (n + 7) / 8;
    std::vector<int> shifts;
    // The sorted keys go straight into the caller's array (no zero-filled staging vectors: at 1e7 determinants
    // their page faults and the extra pass were a third of this phase). wpd == 2 with one-word keys: the packed
    // words land in the first half and are expanded in place from the back.
    std::vector<uint64_t> h2;
    uint64_t* h1 = out_words;
    if (!two) {
      for (int d = 0; d < ndig; ++d) shifts.push_back(32 + 8 * d);  // beta: minor
      for (int d = 0; d < ndig; ++d) shifts.push_back(8 * d);       // alpha: major
      radix_sort_pairs(ctx, k1, k1alt, idx, idx_alt, total, shifts);
      B2_CUDA(cudaMemcpyAsync(h1, k1, size_t(total) * 8, cudaMemcpyDeviceToHost, st));
    } else {
      h2.resize(size_t(total));
      for (int d = 0; d < ndig; ++d) shifts.push_back(8 * d);
      B2_CUDA(cudaMemcpyAsync(kw, k2, size_t(total) * 8, cudaMemcpyDeviceToDevice, st));
      radix_sort_pairs(ctx, kw, k1alt, idx, idx_alt, total, shifts);            // by beta
      k_gather_u64<<<grid1d(total), 256, 0, st>>>(k1, idx, total, kw);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      radix_sort_pairs(ctx, kw, k1alt, idx, idx_alt, total, shifts);            // by alpha, stable
      k_gather_u64<<<grid1d(total), 256, 0, st>>>(k2, idx, total, k1alt);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      B2_CUDA(cudaMemcpyAsync(h1, kw, size_t(total) * 8, cudaMemcpyDeviceToHost, st));       // alpha words, packed
      B2_CUDA(cudaMemcpyAsync(h2.data(), k1alt, size_t(total) * 8, cudaMemcpyDeviceToHost, st));
    }
    B2_CUDA(cudaStreamSynchronize(st));
    mark("output sort + download");
    if (two) {
      for (int64_t i = total - 1; i >= 0; --i) { const uint64_t a = h1[i]; out_words[2 * i] = a; out_words[2 * i + 1] = h2[size_t(i)]; }
    } else if (wpd == 2) {
      for (int64_t i = total - 1; i >= 0; --i) { const uint64_t v = h1[i]; out_words[2 * i] = v & 0xFFFFFFFFull; out_words[2 * i + 1] = v >> 32; }
    }
    mark("words to caller");
    return 0;
  }
  if (nkeep) {
    std::vector<uint64_t> tmp(nkeep), tmp2(two ? nkeep : 0);
    B2_CUDA(cudaMemcpyAsync(tmp.data(), cand_key, size_t(nkeep) * 8, cudaMemcpyDeviceToHost, st));
    if (two) B2_CUDA(cudaMemcpyAsync(tmp2.data(), cand_key2, size_t(nkeep) * 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    if (two)
      for (int64_t i = 0; i < nkeep; ++i) { out_words[2 * i] = tmp[i]; out_words[2 * i + 1] = tmp2[i]; }
    else if (wpd == 1) memcpy(out_words, tmp.data(), size_t(nkeep) * 8);
    else
      for (int64_t i = 0; i < nkeep; ++i) { out_words[2 * i] = tmp[i] & 0xFFFFFFFFull; out_words[2 * i + 1] = tmp[i] >> 32; }
  }
  memcpy(out_words + size_t(nkeep) * wpd, core_words, size_t(nc) * wpd * 8);
  return 0;
}

}  // namespace b2ci
