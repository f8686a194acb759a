// Determinant-space Hamiltonian build: row-partitioned CSR assembly.
//
// Replaces SortedDoubleLoopHamiltonianGenerator::make_csr_hamiltonian_block_
// (external/macis/include/macis/hamiltonian_generator/sorted_double_loop.hpp:86-451) for
// the symmetric (bra == ket) case used by selected_ci_diag. The reference evaluates the
// upper triangle under OpenMP, mirrors it with atomic row cursors, sorts every row and
// finally drops |h| <= H_thresh. Here every row is owned by one warp:
//
//   1. alpha run-length encoding  (get_unique_alpha, sd_operations.hpp:449-468)
//   2. run adjacency: runs whose alpha strings differ by <= 4 bits (XOR + popcount)
//   3. count pass : warp per row, lanes scan the beta strings of each adjacent run with
//                   XOR/popcount, hits are queued in shared memory and evaluated 32 at a
//                   time (full lanes) with the Slater-Condon rules; |h| > thr is counted
//   4. exclusive scan -> rowptr
//   5. fill pass  : same walk, ballot-compacted ordered writes -> columns ascending, no
//                   per-row sort, no atomics, both triangles computed with the SAME
//                   (bra = lower index, ket = higher index) roles as the reference so the
//                   values are bit-identical to its mirrored entries.
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "slater.cuh"

namespace b2ci {

void radix_sort_pairs(b2ci_ctx* ctx, uint64_t* keys, uint64_t* keys_alt, uint32_t* vals,
                      uint32_t* vals_alt, int64_t n, const std::vector<int>& shifts);
void iota_u32(b2ci_ctx* ctx, uint32_t* v, int64_t n);

namespace {

constexpr int ROW_WARPS = 8;  // warps (rows) per CTA

__global__ void k_run_flags(const uint64_t* __restrict__ alpha, int64_t n,
                            int32_t* __restrict__ flag) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) flag[i] = (i == 0 || alpha[i] != alpha[i - 1]) ? 1 : 0;
}
__global__ void k_run_scatter(const uint64_t* __restrict__ alpha, int64_t n,
                              const int32_t* __restrict__ flag,
                              const int32_t* __restrict__ excl, int32_t* __restrict__ run_of,
                              int64_t* __restrict__ run_start, uint64_t* __restrict__ run_alpha) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t r = excl[i] + flag[i] - 1;
  run_of[i] = r;
  if (flag[i]) {
    run_start[r] = i;
    run_alpha[r] = alpha[i];
  }
  if (i == n - 1) run_start[r + 1] = n;  // r + 1 == number of runs
}

// adjacency between bit strings (alpha runs, or the beta template of a rectangular list):
// entry = (string index << 2) | (popcount / 2), ascending in the string index. skip_zero drops
// empty strings on either side (the reference skips alpha-empty determinants, beta-empty ones
// are kept). deg_cnt (count pass, optional): per string the number of neighbours at distance
// 0, 2 and 4.
template <bool FILL>
__global__ void __launch_bounds__(256)
k_string_adjacency(const uint64_t* __restrict__ str, int32_t nstr, int maxd, int skip_zero,
                   int32_t* __restrict__ cnt, int32_t* __restrict__ deg_cnt,
                   const int64_t* __restrict__ adj_ptr, uint32_t* __restrict__ adj,
                   int32_t r_lo = 0, int32_t r_hi = INT32_MAX) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nstr) return;
  if (r < r_lo || r >= r_hi) {  // a row block only needs the adjacency of the runs its rows belong to
    if (!FILL && lane == 0) {
      cnt[r] = 0;
      if (deg_cnt) { deg_cnt[3 * r] = 0; deg_cnt[3 * r + 1] = 0; deg_cnt[3 * r + 2] = 0; }
    }
    return;
  }
  const uint64_t a = str[r];
  int64_t out = FILL ? adj_ptr[r] : 0;
  int32_t c = 0, c0 = 0, c2 = 0, c4 = 0;
  if (!(skip_zero && a == 0)) {
    for (int32_t r0 = 0; r0 < nstr; r0 += 32) {
      const int32_t r2 = r0 + lane;
      bool ok = false;
      int d = 0;
      if (r2 < nstr) {
        const uint64_t a2 = str[r2];
        d = __popcll(a ^ a2);
        ok = !(skip_zero && a2 == 0) && d <= maxd;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (FILL && ok) adj[out + __popc(m & ((1u << lane) - 1u))] = (uint32_t(r2) << 2) | uint32_t(d >> 1);
      out += __popc(m);
      c += __popc(m);
      if (!FILL && deg_cnt) {
        c0 += __popc(__ballot_sync(0xffffffffu, ok && d == 0));
        c2 += __popc(__ballot_sync(0xffffffffu, ok && d == 2));
        c4 += __popc(__ballot_sync(0xffffffffu, ok && d == 4));
      }
    }
  }
  if (!FILL && lane == 0) {
    cnt[r] = c;
    if (deg_cnt) { deg_cnt[3 * r] = c0; deg_cnt[3 * r + 1] = c2; deg_cnt[3 * r + 2] = c4; }
  }
}

struct RowArgs {
  IntsView I;
  const uint64_t* alpha;
  const uint64_t* beta;
  const int32_t* run_of;
  const int64_t* run_start;
  const int64_t* adj_ptr;   // alpha-run adjacency, distance <= 2 (self included), ascending
  const uint32_t* adj;
  const int32_t* bgrp_of;   // beta group of every determinant
  const int64_t* bgrp_start;
  const uint32_t* bgrp_mem; // determinant indices grouped by beta string, ascending inside a group
  int64_t row_begin;
  int64_t nrows;
  double thr;
  int32_t* row_cnt;       // count pass output
  const int64_t* rowptr;  // fill pass input
  int32_t* colind;
  double* nzval;
  // rectangular blocks (BLK): the rows come from a second (bra) list; alpha / beta / run_start /
  // bgrp_* above describe the ket list. Used by the patched (incremental) build.
  const uint64_t* bra_alpha = nullptr;
  const uint64_t* bra_beta = nullptr;
  const int32_t* bra_run = nullptr;  // row of the bra-run x ket-run adjacency of every bra determinant
  const int32_t* bra_grp = nullptr;  // beta group of the ket list holding the bra's beta string, or -1
  const int32_t* rowmap = nullptr;   // index of a bra / ket determinant in the common (global) list:
  const int32_t* colmap = nullptr;   //   decides the (bra, ket) roles and is the column written
  // threshold rule of the pair-based generators (residue_arrays, dynamic_bit_masking:
  // connection_build_utils.hpp:163-199): the diagonal is always kept, an off-diagonal element is
  // dropped only when |h| < thr, and alpha-empty determinants are ordinary determinants
  int pair_rule = 0;
  // hit lists: the count pass can store the structural connections it finds (chunks of 32 column
  // slots chained per row), so that the fill pass evaluates them without scanning again
  int32_t* hit_cols = nullptr;        // [hit_capacity][32] ket indices
  int32_t* hit_next = nullptr;        // [hit_capacity] next chunk of the row, -1 = last
  int32_t* hit_head = nullptr;        // [nrows] first chunk of the row
  unsigned int* hit_cursor = nullptr; // chunks handed out so far (may run past hit_capacity: overflow)
  unsigned int hit_capacity = 0;
  const int64_t* bra_run_start = nullptr;  // BLK: first bra determinant of every bra run
  const int32_t* row_list = nullptr;       // optional: the launch's rows (the rest are taken by the tiled scan)
  const void* beta_s = nullptr;            // ket beta strings as 32-bit words (norb <= 32), else NULL
  const void* alpha_s = nullptr;           // ket alpha strings likewise (the class-binned fill)
  const double* diag = nullptr;            // the class-binned fill: diagonal element of every row (k_row_diag)
  const int32_t* struct_cnt = nullptr;  // gathers and the fills from connections: structural row lengths of the count pass
  int64_t row_stride = 1;               // sampling (estimate pass): row r of the launch is row r * row_stride
};

template <typename S> __device__ __forceinline__ int popc_s(S x);
template <> __device__ __forceinline__ int popc_s<uint32_t>(uint32_t x) { return __popc(x); }
template <> __device__ __forceinline__ int popc_s<uint64_t>(uint64_t x) { return __popcll(x); }

// Evaluate up to 32 queued column indices (one per lane) and count / emit the survivors.
// count pass with hit lists: one chunk per batch, chained to the row's previous chunk
__device__ __forceinline__ void store_hits(const RowArgs& A, int64_t row, const int32_t* q, int nvalid, int lane,
                                           int32_t& prev_chunk) {
  unsigned int c = 0;
  if (lane == 0) c = atomicAdd(A.hit_cursor, 1u);
  c = __shfl_sync(0xffffffffu, c, 0);
  if (c >= A.hit_capacity) return;  // overflow: the host sees the cursor and falls back to the scan
  if (lane < nvalid) A.hit_cols[size_t(c) * 32 + lane] = q[lane];
  if (lane == 0) {
    A.hit_next[c] = -1;
    if (prev_chunk >= 0) A.hit_next[prev_chunk] = int32_t(c);
    else A.hit_head[row] = int32_t(c);
  }
  prev_chunk = int32_t(c);
}

template <bool FILL, bool EVAL, bool BLK>
__device__ __forceinline__ void process_batch(const RowArgs& A, int64_t i, uint64_t ai,
                                              uint64_t bi, const int32_t* q, int nvalid, int lane,
                                              int64_t& out, int32_t& cnt) {
  const bool valid = lane < nvalid;
  int32_t j = 0;
  double v = 0.;
  bool keep = valid;
  if (valid) {
    j = q[lane];
    if (FILL || EVAL) {
      const uint64_t aj = A.alpha[j], bj = A.beta[j];
      if (BLK && A.colmap) j = A.colmap[j];
      // the reference computes the upper triangle with bra = lower index and mirrors it
      v = (i <= int64_t(j)) ? matel(A.I, ai, bi, aj, bj) : matel(A.I, aj, bj, ai, bi);
      if (EVAL) keep = A.pair_rule ? (i == int64_t(j) || !(fabs(v) < A.thr)) : fabs(v) > A.thr;
    }
  }
  const unsigned km = __ballot_sync(0xffffffffu, keep);
  if (FILL && keep) {
    const int64_t pos = out + __popc(km & ((1u << lane) - 1u));
    A.colind[pos] = j;
    A.nzval[pos] = v;
  }
  out += __popc(km);
  cnt += __popc(km);
}

// Row i = (alpha_i, beta_i) of an arbitrary list connects to
//   (a) the determinants of its own alpha run with beta distance <= 4,
//   (b) the determinants of runs one alpha single excitation away with beta distance <= 2,
//   (c) the determinants with the SAME beta string whose alpha string is a double excitation.
// (a) and (b) are found by scanning the beta strings of the <= 1 + n_occ n_virt adjacent runs;
// (c) comes from the list of determinants grouped by beta string -- no scan of the (thousands
// of) runs two alpha excitations away. Runs are contiguous index ranges and group members are
// stored in ascending index, so the two streams are merged on the fly: before run r' is
// scanned, the group members below its first index are emitted. Columns come out ascending.
template <bool FILL, bool EVAL, bool BLK = false, typename S = uint64_t>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_rows(const RowArgs A) {
  __shared__ int32_t queue[ROW_WARPS][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t ridx = int64_t(blockIdx.x) * ROW_WARPS + w;
  if (ridx >= A.nrows) return;
  const int64_t row = A.row_list ? int64_t(A.row_list[ridx]) : ridx;
  const int64_t il = A.row_begin + row * A.row_stride;  // index in the bra list
  const uint64_t ai = BLK ? A.bra_alpha[il] : A.alpha[il], bi = BLK ? A.bra_beta[il] : A.beta[il];
  const int32_t r = BLK ? A.bra_run[il] : A.run_of[il];
  const int64_t i = (BLK && A.rowmap) ? int64_t(A.rowmap[il]) : il;  // index in the common list
  int32_t* q = queue[w];
  int qn = 0;
  int32_t cnt = 0;
  int64_t out = FILL ? A.rowptr[row] : 0;
  const unsigned lt = (1u << lane) - 1u;
  int32_t prev_chunk = -1;
  const bool keep_hits = !FILL && !EVAL && A.hit_cols != nullptr;
  // enqueue the hits of one ballot; evaluate / count / write 32 at a time with full lanes
  auto push = [&](bool hit, int64_t j) {
    const unsigned m = __ballot_sync(0xffffffffu, hit);
    if (!m) return;
    if (hit) q[qn + __popc(m & lt)] = int32_t(j);
    qn += __popc(m);
    __syncwarp();
    if (qn >= 32) {
      if (keep_hits) store_hits(A, row, q, 32, lane, prev_chunk);
      process_batch<FILL, EVAL, BLK>(A, i, ai, bi, q, 32, lane, out, cnt);
      const int rest = qn - 32;
      const int32_t t = (lane < rest) ? q[32 + lane] : 0;
      __syncwarp();
      if (lane < rest) q[lane] = t;
      qn = rest;
      __syncwarp();
    }
  };
  if (ai != 0 || A.pair_rule) {
    const int32_t g = BLK ? A.bra_grp[il] : A.bgrp_of[il];
    int64_t bpos = g >= 0 ? A.bgrp_start[g] : 0;
    const int64_t bend = g >= 0 ? A.bgrp_start[g + 1] : 0;
    int64_t nextj = bpos < bend ? int64_t(A.bgrp_mem[bpos]) : INT64_MAX;
    // (c): members of the beta group with index < bound
    auto flush_group = [&](int64_t bound) {
      while (nextj < bound) {
        const int64_t p = bpos + lane;
        const int64_t j = p < bend ? int64_t(A.bgrp_mem[p]) : INT64_MAX;
        const bool in = j < bound;
        bool hit = false;
        if (in) {
          const uint64_t aj = A.alpha[j];
          hit = (aj != 0 || A.pair_rule) && __popcll(ai ^ aj) == 4;
        }
        const int nin = __popc(__ballot_sync(0xffffffffu, in));  // a prefix: members ascend
        push(hit, j);
        bpos += nin;
        nextj = bpos < bend ? int64_t(A.bgrp_mem[bpos]) : INT64_MAX;
        if (nin < 32) break;
      }
    };
    const int64_t e0 = A.adj_ptr[r], e1 = A.adj_ptr[r + 1];
    for (int64_t e = e0; e < e1; ++e) {
      const uint32_t pk = A.adj[e];
      const int da = int(pk & 3u) * 2;
      const int64_t ks = A.run_start[pk >> 2], ke = A.run_start[(pk >> 2) + 1];
      flush_group(ks);
      {
        // the hot loop of a general build (876 steps per row at 2e5 determinants, 61 % of the kernel's
        // instructions): 32-bit positions, one pointer, no bounds test in the full steps
        const int32_t len = int32_t(ke - ks), j32 = int32_t(ks);
        // strings narrowed to 32 bits when norb <= 32 (A.beta_s): half the POPC work and traffic
        const S* __restrict__ bp = (sizeof(S) == 8 ? reinterpret_cast<const S*>(A.beta) : static_cast<const S*>(A.beta_s)) + ks + lane;
        const S bs = S(bi);
        const int lim = 4 - da;
        int32_t base = 0;
        for (; base + 32 <= len; base += 32) {
          const bool hit = popc_s<S>(bs ^ __ldg(bp + base)) <= lim;
          push(hit, j32 + base + lane);
        }
        if (base < len) {
          bool hit = false;
          if (base + lane < len) hit = popc_s<S>(bs ^ __ldg(bp + base)) <= lim;
          push(hit, j32 + base + lane);
        }
      }
      // group members inside this run are at alpha distance <= 2: already covered by the scan
      while (nextj < ke) {
        const int64_t p = bpos + lane;
        const int64_t j = p < bend ? int64_t(A.bgrp_mem[p]) : INT64_MAX;
        const int nin = __popc(__ballot_sync(0xffffffffu, j < ke));
        bpos += nin;
        nextj = bpos < bend ? int64_t(A.bgrp_mem[bpos]) : INT64_MAX;
        if (nin < 32) break;
      }
    }
    flush_group(INT64_MAX);
    if (qn > 0) {
      if (keep_hits) store_hits(A, row, q, qn, lane, prev_chunk);
      process_batch<FILL, EVAL, BLK>(A, i, ai, bi, q, qn, lane, out, cnt);
    }
  }
  if (lane == 0 && A.row_cnt) A.row_cnt[row] = cnt;
}


// ------------------------------------------------------------------ tiled scan of general lists
// The warp-per-row scan above spends one instruction stream per (row, 32 strings): every row of an
// alpha run re-reads the beta strings of the same adjacent runs. Here a warp owns a UNIT of up to
// 32 consecutive rows of ONE alpha run, one row per lane: the strings of an adjacent run are staged
// 32 at a time in shared memory and every staged string is tested against the 32 row strings at once
// (broadcast shared-memory read, then per string one XOR + POPC + compare and a predicated OR into
// the lane's hit word; votes, counts and stores once per step of 32 strings after a warp bit
// transpose) -- the adjacency walk, the run bounds and the loads are paid once per unit instead of
// once per row, and the test is the only per-(row, string) work. With norb <= 32 the strings are
// 32-bit words (half the POPC work; POPC issues at 16 lanes / clock / SM on B200, measured by
// scripts/micro/popc_rate.cu, and bounds this kernel). A lane's scan hits ascend in the ket index by
// construction (runs ascend, strings ascend); the beta-group members of class (c) come first in the
// unit's stream and k_tile_gather merges them into each row: rows need no sort.
//
// Connections leave the kernel as a dense stream of (lane mask, ket index) entries per unit -- one
// entry per tested string that connects to at least one row of the unit, 8 bytes, so the store is
// bounded by 8 bytes per connection whatever the rows look like -- in chained chunks of TILE_CH
// entries (one atomic per chunk, warp-uniform). k_tile_gather then deals every row's hits to its
// slot range of the CSR column array, where k_rows_hits_binned evaluates them in place. If the store
// overflows the counts are still right and the fill falls back to scanning again.
constexpr int TILE_CH = 512;  // entries per chunk (4 KiB)
constexpr int TILE_WARPS = 8;
struct TileArgs {
  RowArgs A;
  const void* beta_s;        // ket beta strings as 32- or 64-bit words
  const int32_t* unit_row0;  // first local row of every tiled unit
  const int32_t* unit_len;   // its rows (<= 32)
  int32_t nunits;
  uint2* tile_ent;           // [capacity][TILE_CH] (lane mask, ket index)
  int32_t* tile_next;        // [capacity]
  int32_t* unit_head;        // [nunits] first chunk (-1: none)
  int32_t* unit_nent;        // [nunits] entries of the unit (all chunks full but the last)
  int32_t* unit_nent_c;      // [nunits] of them class (c) entries (they come first)
  int32_t* row_cnt_c;        // [nrows] class (c) connections of every row (tiled rows only)
  unsigned int* cursor;
  unsigned int capacity;
};

// 32 x 32 bit transpose across a warp: in, lane l holds brev(row (31 - l)); out, lane l holds the word whose
// bit k is bit l of row k (five butterfly steps, Hacker's Delight 7-3 in warp form)
__device__ __forceinline__ unsigned warp_bit_transpose(unsigned x, int lane) {
  unsigned m = 0x0000FFFFu;
#pragma unroll
  for (int j = 16; j; j >>= 1) {
    const unsigned y = __shfl_xor_sync(0xffffffffu, x, j);
    if (!(lane & j)) x ^= (x ^ (y >> j)) & m;
    else x ^= ((y ^ (x >> j)) & m) << j;
    m ^= m << (j >> 1);
  }
  return x;
}
template <typename S, bool BLK>
__global__ void __launch_bounds__(TILE_WARPS * 32)
k_rows_tile(const TileArgs T) {
  __shared__ __align__(16) S slab[TILE_WARPS][32];
  const RowArgs& A = T.A;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t u = int64_t(blockIdx.x) * TILE_WARPS + w;
  if (u >= T.nunits) return;
  const int32_t row0 = T.unit_row0[u], ulen = T.unit_len[u];
  const bool valid = lane < ulen;
  const int64_t row = row0 + (valid ? lane : 0);
  const int64_t il = A.row_begin + row;
  const uint64_t ai = BLK ? A.bra_alpha[il] : A.alpha[il];  // the same for every row of the unit
  const S bi = S(BLK ? A.bra_beta[il] : A.beta[il]);
  const int32_t r = BLK ? A.bra_run[il] : A.run_of[il];
  const S* __restrict__ beta_s = static_cast<const S*>(T.beta_s);
  const unsigned lt = (1u << lane) - 1u;
  int32_t cnt = 0, cnt_c = 0;           // this lane's row: all connections / those of class (c)
  int32_t chunk = -1, pos = TILE_CH;    // warp-uniform stream state
  int32_t nent = 0;
  uint2* __restrict__ ent = nullptr;    // current chunk
  auto new_chunk = [&]() {              // warp-uniform
    unsigned int c = 0;
    if (lane == 0) c = atomicAdd(T.cursor, 1u);
    c = __shfl_sync(0xffffffffu, c, 0);
    const int32_t nc = c < T.capacity ? int32_t(c) : -1;  // overflow: counting goes on, nothing is stored
    if (lane == 0) {
      if (chunk >= 0) T.tile_next[chunk] = nc;
      else if (nent == 0) T.unit_head[u] = nc;
      if (nc >= 0) T.tile_next[nc] = -1;
    }
    chunk = nc;
    pos = 0;
    ent = nc >= 0 ? T.tile_ent + size_t(nc) * TILE_CH : nullptr;
  };
  int32_t nent_c = 0;
  if (ai != 0 || A.pair_rule) {
    // ---- class (c) first: same beta string, alpha double excitation. All rows of the unit share the alpha
    // string, so row by row the warp walks the row's beta group with one member per LANE (a per-lane walk of 32
    // different groups ran with 3 of 32 lanes busy and was a third of this kernel's instructions); the hits of a
    // step leave together as single-row entries. They ascend per row; k_tile_gather merges them with the row's
    // scan hits, which follow in the stream.
    const int32_t g_l = valid ? (BLK ? A.bra_grp[il] : A.bgrp_of[il]) : -1;
    for (int l = 0; l < ulen; ++l) {
      const int32_t g = __shfl_sync(0xffffffffu, g_l, l);
      if (g < 0) continue;
      const int64_t b0 = A.bgrp_start[g], b1 = A.bgrp_start[g + 1];
      int32_t nrow = 0;
      for (int64_t p0 = b0; p0 < b1; p0 += 32) {
        const int64_t p = p0 + lane;
        bool hit = false;
        int32_t j = 0;
        if (p < b1) {
          j = int32_t(A.bgrp_mem[p]);
          const uint64_t aj = A.alpha[j];
          hit = (aj != 0 || A.pair_rule) && __popcll(ai ^ aj) == 4;
        }
        const unsigned m = __ballot_sync(0xffffffffu, hit);
        if (!m) continue;
        const int nh = __popc(m);
        // room for nh entries in the current chunk (else finish it: a chunk is never left partly empty in the
        // middle of the stream, so pad by splitting the step)
        int done = 0;
        while (done < nh) {
          if (pos == TILE_CH) new_chunk();
          const int take = min(nh - done, TILE_CH - pos);
          const int my = __popc(m & lt);
          if (hit && my >= done && my < done + take && ent) ent[pos + my - done] = make_uint2(1u << l, unsigned(j));
          pos += take;
          nent += take;
          done += take;
        }
        nrow += nh;
      }
      if (lane == l) cnt_c = nrow;
    }
    nent_c = nent;
    cnt = cnt_c;
    // ---- classes (a) and (b): the beta strings of the adjacent runs (own run: distance <= 4, runs one alpha
    // single away: distance <= 2), every staged string against the 32 row strings
    const int64_t e0 = A.adj_ptr[r], e1 = A.adj_ptr[r + 1];
    for (int64_t e = e0; e < e1; ++e) {
      const uint32_t pk = A.adj[e];
      const int32_t ks = int32_t(A.run_start[pk >> 2]), ke = int32_t(A.run_start[(pk >> 2) + 1]);
      const int lim = valid ? 4 - int(pk & 3u) * 2 : -1;  // idle lanes never hit
      const S* __restrict__ bp = beta_s + ks;
      const int32_t rlen = ke - ks;
      S mine = lane < rlen ? bp[lane] : S(0);
      for (int32_t base = 0; base < rlen; base += 32) {
        __syncwarp();
        slab[w][lane] = mine;
        __syncwarp();
        // the next step's strings are in flight while this step is tested
        mine = base + 32 + lane < rlen ? bp[base + 32 + lane] : S(0);
        const int nst = rlen - base < 32 ? rlen - base : 32;
        const int32_t jb = ks + base;
        // bit t of rowbits: string t of the step connects to this lane's row. The per-string work is XOR, POPC,
        // compare and one predicated OR; votes, counts and stores are paid once per step of 32 strings.
        unsigned rowbits = 0u;
        int q = 0;
        for (; q + 8 <= nst; q += 8) {
          S st[8];
          if (sizeof(S) == 4) {
            const uint4 v0 = *reinterpret_cast<const uint4*>(&slab[w][q]);
            const uint4 v1 = *reinterpret_cast<const uint4*>(&slab[w][q + 4]);
            st[0] = S(v0.x); st[1] = S(v0.y); st[2] = S(v0.z); st[3] = S(v0.w);
            st[4] = S(v1.x); st[5] = S(v1.y); st[6] = S(v1.z); st[7] = S(v1.w);
          } else {
#pragma unroll
            for (int t = 0; t < 8; t += 2) {
              const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(&slab[w][q + t]);
              st[t] = S(v.x); st[t + 1] = S(v.y);
            }
          }
          unsigned g = 0u;
#pragma unroll
          for (int t = 0; t < 8; ++t)
            if (popc_s<S>(bi ^ st[t]) <= lim) g |= 1u << t;
          rowbits |= g << q;
        }
        for (int t = q; t < nst; ++t)
          if (popc_s<S>(bi ^ slab[w][t]) <= lim) rowbits |= 1u << t;
        if (!__any_sync(0xffffffffu, rowbits != 0u)) continue;
        cnt += __popc(rowbits);
        // lane t <- the row mask of string t (bit l: row l connects), then the strings that connect to some row
        // leave as one contiguous run of entries
        const unsigned smask = warp_bit_transpose(__brev(__shfl_sync(0xffffffffu, rowbits, 31 - lane)), lane);
        const unsigned nz = __ballot_sync(0xffffffffu, smask != 0u);
        const int nh = __popc(nz), my = __popc(nz & lt);
        int done = 0;
        while (done < nh) {
          if (pos == TILE_CH) new_chunk();
          const int take = min(nh - done, TILE_CH - pos);
          if (smask != 0u && my >= done && my < done + take && ent) ent[pos + my - done] = make_uint2(smask, unsigned(jb + lane));
          pos += take;
          nent += take;
          done += take;
        }
      }
    }
  }
  if (lane == 0) { T.unit_nent[u] = nent; T.unit_nent_c[u] = nent_c; }
  if (valid && A.row_cnt) A.row_cnt[row] = cnt;
  if (valid && T.row_cnt_c) T.row_cnt_c[row] = cnt_c;
}

// deal the entries of every unit to its rows: hits of row (unit, lane) = ket indices of the entries whose mask has
// bit `lane`, written to the row's slot range of `hits` in ascending order. The stream holds the class (c) entries
// first: a lane parks its (c) hits at the END of its range and merges them in while it walks the scan entries --
// in place, because the write position never passes the first unread parked element.
__global__ void __launch_bounds__(TILE_WARPS * 32)
k_tile_gather(const TileArgs T, const int64_t* __restrict__ slot_ptr, int32_t* __restrict__ hits) {
  __shared__ int32_t jbuf[TILE_WARPS][32];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t u = int64_t(blockIdx.x) * TILE_WARPS + w;
  if (u >= T.nunits) return;
  const int32_t row0 = T.unit_row0[u], ulen = T.unit_len[u];
  const bool valid = lane < ulen;
  int32_t* __restrict__ out = hits + (valid ? slot_ptr[row0 + lane] : 0);
  const int32_t n_tot = valid ? T.A.struct_cnt[row0 + lane] : 0;
  const int32_t n_c = valid ? T.row_cnt_c[row0 + lane] : 0;
  int32_t* __restrict__ park = out + (n_tot - n_c);
  int32_t remaining = T.unit_nent[u];
  int32_t rem_c = T.unit_nent_c[u];
  int32_t c = remaining > 0 ? T.unit_head[u] : -1;
  int32_t o = 0, cp = 0, parked = 0;
  int32_t cnext = INT32_MAX;  // next parked element (value), loaded when the (c) part is complete
  bool c_loaded = false;
  while (remaining > 0 && c >= 0) {
    const int32_t nin = remaining < TILE_CH ? remaining : TILE_CH;
    const uint2* __restrict__ src = T.tile_ent + size_t(c) * TILE_CH;
    for (int32_t b = 0; b < nin;) {
      int nb = nin - b < 32 ? nin - b : 32;
      const bool cpart = rem_c > 0;        // warp-uniform: a block never straddles the (c) / scan boundary
      if (cpart && rem_c < nb) nb = rem_c;
      // lane l takes entry (31 - l) of the block: after the transpose lane l holds, bit k, "entry k is mine"
      const int kk = 31 - lane;
      const uint2 e = kk < nb ? src[b + kk] : make_uint2(0u, 0u);
      __syncwarp();
      jbuf[w][kk] = int32_t(e.y);
      __syncwarp();
      unsigned mine = warp_bit_transpose(__brev(e.x), lane);
      if (cpart) {
        while (mine) {
          const int k = __ffs(mine) - 1;
          mine &= mine - 1;
          park[parked++] = jbuf[w][k];
        }
        rem_c -= nb;
      } else {
        if (!c_loaded) {  // first scan block: the parked list is complete (own writes, same thread)
          c_loaded = true;
          cnext = n_c > 0 ? park[0] : INT32_MAX;
        }
        while (mine) {
          const int k = __ffs(mine) - 1;
          mine &= mine - 1;
          const int32_t jk = jbuf[w][k];
          while (cnext < jk) {
            out[o++] = cnext;
            ++cp;
            cnext = cp < n_c ? park[cp] : INT32_MAX;
          }
          out[o++] = jk;
        }
      }
      b += nb;
    }
    remaining -= nin;
    c = T.tile_next[c];
  }
  // parked elements left over are already where they belong (o == n_tot - n_c + cp)
}

__global__ void k_unit_flags(int64_t nrows, int64_t row_begin, const int32_t* __restrict__ run_ix,
                             const int64_t* __restrict__ run_first, int32_t* __restrict__ flag) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  const int64_t il = row_begin + r;
  const int32_t run = run_ix[il];
  flag[r] = (r == 0 || run != run_ix[il - 1] || ((il - run_first[run]) & 31) == 0) ? 1 : 0;
}
// per row: is it the first row of a unit (flag), which unit (excl + flag - 1); a unit of fewer than
// min_rows rows is left to the warp-per-row scan (one lane per row would idle the other lanes)
__global__ void k_unit_classify(int64_t nrows, const int32_t* __restrict__ flag, const int32_t* __restrict__ excl,
                                int32_t min_rows, int32_t* __restrict__ tile_flag, int32_t* __restrict__ scan_flag) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  // length of the unit r belongs to: distance between the unit starts around r (<= 32 rows either way)
  int64_t a = r, b = r + 1;
  while (!flag[a]) --a;
  while (b < nrows && !flag[b]) ++b;
  const bool big = b - a >= min_rows;
  tile_flag[r] = (big && flag[r]) ? 1 : 0;
  scan_flag[r] = big ? 0 : 1;
  (void)excl;
}
__global__ void k_unit_lists(int64_t nrows, const int32_t* __restrict__ flag, const int32_t* __restrict__ tile_flag,
                             const int32_t* __restrict__ tile_excl, const int32_t* __restrict__ scan_flag,
                             const int32_t* __restrict__ scan_excl, int32_t* __restrict__ unit_row0,
                             int32_t* __restrict__ unit_len, int32_t* __restrict__ unit_head,
                             int32_t* __restrict__ scan_rows) {
  const int64_t r = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (r >= nrows) return;
  if (tile_flag[r]) {
    int64_t b = r + 1;
    while (b < nrows && !flag[b]) ++b;
    const int32_t t = tile_excl[r];
    unit_row0[t] = int32_t(r);
    unit_len[t] = int32_t(b - r);
    unit_head[t] = -1;  // k_rows_tile sets it when the unit produces its first entry
  }
  if (scan_flag[r]) scan_rows[scan_excl[r]] = int32_t(r);
}
// hits of the warp-per-row scan (chunks of 32 chained per row) to the row's slot range, for the rows of a list
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_hits_gather(const RowArgs A, const int64_t* __restrict__ slot_ptr, int32_t* __restrict__ hits) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t ridx = int64_t(blockIdx.x) * ROW_WARPS + w;
  if (ridx >= A.nrows) return;
  const int64_t row = A.row_list ? int64_t(A.row_list[ridx]) : ridx;
  int32_t remaining = A.struct_cnt[row];
  int32_t chunk = remaining > 0 ? A.hit_head[row] : -1;
  int32_t* __restrict__ out = hits + slot_ptr[row];
  while (remaining > 0) {
    const int nvalid = remaining < 32 ? remaining : 32;
    if (lane < nvalid) out[lane] = A.hit_cols[size_t(chunk) * 32 + lane];
    out += nvalid;
    remaining -= nvalid;
    chunk = A.hit_next[chunk];
  }
}
__global__ void k_narrow_u32(const uint64_t* __restrict__ in, int64_t n, uint32_t* __restrict__ out) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < n) out[i] = uint32_t(in[i]);
}

// fill pass from contiguous hit lists held IN PLACE in the row's slot range of colind
template <bool EVAL, bool BLK>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_rows_hits_flat(const RowArgs A) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = int64_t(blockIdx.x) * ROW_WARPS + w;
  if (row >= A.nrows) return;
  const int64_t il = A.row_begin + row;
  const uint64_t ai = BLK ? A.bra_alpha[il] : A.alpha[il], bi = BLK ? A.bra_beta[il] : A.beta[il];
  const int64_t i = (BLK && A.rowmap) ? int64_t(A.rowmap[il]) : il;
  const int32_t nhit = A.struct_cnt[row];
  const int64_t slot = A.rowptr[row];
  int64_t out = slot;
  int32_t cnt = 0;
  for (int32_t t0 = 0; t0 < nhit; t0 += 32) {
    const int nvalid = nhit - t0 < 32 ? nhit - t0 : 32;
    // survivors are written at or below the position they were read from, after the whole batch was read
    const int32_t jraw = lane < nvalid ? A.colind[slot + t0 + lane] : 0;
    __syncwarp();
    const bool valid = lane < nvalid;
    int32_t j = jraw;
    double v = 0.;
    bool keep = valid;
    if (valid) {
      const uint64_t aj = A.alpha[j], bj = A.beta[j];
      if (BLK && A.colmap) j = A.colmap[j];
      v = (i <= int64_t(j)) ? matel(A.I, ai, bi, aj, bj) : matel(A.I, aj, bj, ai, bi);
      if (EVAL) keep = A.pair_rule ? (i == int64_t(j) || !(fabs(v) < A.thr)) : fabs(v) > A.thr;
    }
    const unsigned km = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t pos = out + __popc(km & ((1u << lane) - 1u));
      A.colind[pos] = j;
      A.nzval[pos] = v;
    }
    out += __popc(km);
    cnt += __popc(km);
  }
  if (lane == 0 && A.row_cnt) A.row_cnt[row] = cnt;
}

// ket string j as an S word: from the narrowed copy when there is one
template <typename S>
__device__ __forceinline__ S ket_string(const uint64_t* __restrict__ wide, const void* __restrict__ narrow, int32_t j) {
  if (sizeof(S) == 4) return static_cast<const S*>(narrow)[j];
  return S(wide[j]);
}
// The same fill with the lanes of a warp kept on ONE excitation class at a time. A row's connections arrive in
// column order, classes interleaved: evaluated 32 at a time as they come (k_rows_hits_flat) the warp walks every
// branch of the Slater-Condon dispatch with a third of its lanes (ncu: 11 of 32 threads per instruction, 20 warp
// instructions per matrix element). Here a batch is only CLASSIFIED (two POPCs); (position, column) pairs queue
// per class in shared memory and a class is evaluated whenever 32 of its connections are waiting -- opposite-spin
// doubles, same-spin doubles (spin picked by select) and singles (likewise) each as straight-line code. Values
// land at their structural position; a dropped element is marked by colind = -1 and k_compact_rows_holes packs
// the rows that the thresholded build packs anyway. Same functions, same operands: the same bits.
template <int C, bool EVAL, bool BLK, typename S>
__device__ __forceinline__ int eval_queued(const RowArgs& A, const uint2 e, const bool act, const S ai, const S bi,
                                           const int64_t i, const int64_t row, const int64_t slot) {
  if (!act) return 0;
  int32_t j = int32_t(e.y);
  const S aj = ket_string<S>(A.alpha, A.alpha_s, j), bj = ket_string<S>(A.beta, A.beta_s, j);
  if (BLK && A.colmap) j = A.colmap[j];
  const bool fwd = i <= int64_t(j);  // bra = the lower determinant index, as the reference's upper triangle
  const S bra_a = fwd ? ai : aj, bra_b = fwd ? bi : bj, ket_a = fwd ? aj : ai, ket_b = fwd ? bj : bi;
  const S ex_a = ai ^ aj, ex_b = bi ^ bj;
  double v;
  if (C == 0) {
    v = me22<S>(A.I, bra_a, ket_a, ex_a, bra_b, ket_b, ex_b);
  } else if (C == 1) {
    const bool sa = ex_a != 0;
    v = me4<S>(A.I, sa ? bra_a : bra_b, sa ? ket_a : ket_b, sa ? ex_a : ex_b);
  } else {
    const int ca = popc_w<S>(ex_a), cb = popc_w<S>(ex_b);
    if (ca + cb == 2) {
      const bool sa = ca == 2;
      v = me2<S>(A.I, sa ? bra_a : bra_b, sa ? ket_a : ket_b, sa ? ex_a : ex_b, sa ? bra_a : bra_b, sa ? bra_b : bra_a);
    } else if (ca + cb == 0 && A.diag) {
      v = A.diag[row];  // k_row_diag: one thread per row instead of one lane of this warp
    } else {
      v = matel(A.I, uint64_t(bra_a), uint64_t(bra_b), uint64_t(ket_a), uint64_t(ket_b));
    }
  }
  bool keep = true;
  if (EVAL) keep = A.pair_rule ? (i == int64_t(j) || !(fabs(v) < A.thr)) : fabs(v) > A.thr;
  const int64_t pos = slot + e.x;
  A.nzval[pos] = v;
  if (!keep) A.colind[pos] = -1;
  else if (BLK && A.colmap) A.colind[pos] = j;
  return keep ? 0 : 1;
}
// diagonal elements of the launch's rows, one thread per row (the double loops over the occupied orbitals are
// ~3,000 dependent instructions: inside the fill they ran on one lane of the row's warp and were a fifth of it)
template <bool BLK>
__global__ void k_row_diag(const RowArgs A, double* __restrict__ diag) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= A.nrows) return;
  const int64_t il = A.row_begin + row;
  diag[row] = me_diag(A.I, BLK ? A.bra_alpha[il] : A.alpha[il], BLK ? A.bra_beta[il] : A.beta[il]);
}
constexpr int BIN_WARPS = 8;
template <bool EVAL, bool BLK, typename S>
__global__ void __launch_bounds__(BIN_WARPS * 32)
k_rows_hits_binned(const RowArgs A) {
  __shared__ uint2 queue[BIN_WARPS][3][64];
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t row = int64_t(blockIdx.x) * BIN_WARPS + w;
  if (row >= A.nrows) return;
  const int64_t il = A.row_begin + row;
  const S ai = S(BLK ? A.bra_alpha[il] : A.alpha[il]), bi = S(BLK ? A.bra_beta[il] : A.beta[il]);
  const int64_t i = (BLK && A.rowmap) ? int64_t(A.rowmap[il]) : il;
  const int32_t nhit = A.struct_cnt[row];
  const int64_t slot = A.rowptr[row];
  const unsigned lt = (1u << lane) - 1u;
  int qn0 = 0, qn1 = 0, qn2 = 0;
  int32_t ndrop = 0;
  for (int32_t t0 = 0; t0 < nhit; t0 += 32) {
    const int32_t t = t0 + lane;
    int cls = 3;
    int32_t j = 0;
    if (t < nhit) {
      j = A.colind[slot + t];
      const int ca = popc_w<S>(ai ^ ket_string<S>(A.alpha, A.alpha_s, j)), cb = popc_w<S>(bi ^ ket_string<S>(A.beta, A.beta_s, j));
      cls = (ca == 2 && cb == 2) ? 0 : ((ca + cb == 4) ? 1 : 2);
    }
    const unsigned m0 = __ballot_sync(0xffffffffu, cls == 0), m1 = __ballot_sync(0xffffffffu, cls == 1),
                   m2 = __ballot_sync(0xffffffffu, cls == 2);
    const uint2 e = make_uint2(unsigned(t), unsigned(j));
    if (cls == 0) queue[w][0][qn0 + __popc(m0 & lt)] = e;
    else if (cls == 1) queue[w][1][qn1 + __popc(m1 & lt)] = e;
    else if (cls == 2) queue[w][2][qn2 + __popc(m2 & lt)] = e;
    qn0 += __popc(m0); qn1 += __popc(m1); qn2 += __popc(m2);
    __syncwarp();
    // (a lane may evaluate one connection of every class in the same step: count, do not OR)
    if (qn0 >= 32) { qn0 -= 32; ndrop += eval_queued<0, EVAL, BLK, S>(A, queue[w][0][qn0 + lane], true, ai, bi, i, row, slot); }
    if (qn1 >= 32) { qn1 -= 32; ndrop += eval_queued<1, EVAL, BLK, S>(A, queue[w][1][qn1 + lane], true, ai, bi, i, row, slot); }
    if (qn2 >= 32) { qn2 -= 32; ndrop += eval_queued<2, EVAL, BLK, S>(A, queue[w][2][qn2 + lane], true, ai, bi, i, row, slot); }
    __syncwarp();
  }
  if (qn0 > 0) ndrop += eval_queued<0, EVAL, BLK, S>(A, queue[w][0][lane], lane < qn0, ai, bi, i, row, slot);
  if (qn1 > 0) ndrop += eval_queued<1, EVAL, BLK, S>(A, queue[w][1][lane], lane < qn1, ai, bi, i, row, slot);
  if (qn2 > 0) ndrop += eval_queued<2, EVAL, BLK, S>(A, queue[w][2][lane], lane < qn2, ai, bi, i, row, slot);
  if (EVAL) {  // this lane's drops -> the row's
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) ndrop += __shfl_xor_sync(0xffffffffu, ndrop, d);
  }
  if (lane == 0 && A.row_cnt) A.row_cnt[row] = nhit - ndrop;
}

// beta groups: determinant indices sorted by beta string (stable), group boundaries
__global__ void k_group_scatter(const uint64_t* __restrict__ key_sorted, const uint32_t* __restrict__ idx_sorted,
                                int64_t n, const int32_t* __restrict__ flag, const int32_t* __restrict__ excl,
                                int32_t* __restrict__ grp_of, int64_t* __restrict__ grp_start, int32_t ngroups) {
  const int64_t p = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (p >= n) return;
  const int32_t g = excl[p] + flag[p] - 1;
  grp_of[idx_sorted[p]] = g;
  if (flag[p]) grp_start[g] = p;
  if (p == n - 1) grp_start[ngroups] = n;
}

// ------------------------------------------------------------------ rectangular lists
// A determinant list is "rectangular" when it is the product of R alpha strings and one common
// sequence of Nb beta strings (index = r * Nb + k) -- every list generate_hilbert_space makes.
// Then no beta scan is needed at all: with alpha-run adjacency A(r) and beta adjacency lists
// B2(k) (distance <= 2) and B4(k) (distance <= 4), row (r, k) is exactly
//   { (r', k') : r' in A(r), k' in B_{4 - d_alpha(r, r')}(k) },
// already in ascending column order. Every lane evaluates a real matrix element.
__global__ void k_check_rect(const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
                             int64_t n, const int32_t* __restrict__ nruns_dev, int* __restrict__ bad) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int64_t nruns = *nruns_dev;
  if (nruns <= 0 || n % nruns != 0 || n / nruns >= (int64_t(1) << 29)) {
    if (i == 0) atomicOr(bad, 1);
    return;
  }
  const int64_t nb = n / nruns;
  const int64_t r = i / nb, k = i % nb;
  bool ok = beta[i] == beta[k] && alpha[i] == alpha[r * nb];
  if (k == 0 && r > 0) ok = ok && alpha[i] != alpha[i - 1];
  if (!ok) atomicOr(bad, 1);
}

// Per-pair metadata, computed once per (alpha run pair) / (beta template pair) instead of once
// per matrix element. Orientation everywhere: bra = lower determinant index, as the reference
// (it evaluates the upper triangle and mirrors it).
//   same-spin double : the VALUE  sign * (V(v1,o1,v2,o2) - V(v1,o2,v2,o1))  (matrix_elements.hpp:113-121)
//   single           : (hole, particle, sign) and the leading part of the single-excitation sum
//                      T(v,o) + sum_{p in occ_same(bra), ascending} G_red(p,v,o)   (:176-182);
//                      the kernel appends the other-spin terms V_red(p,v,o) in ascending p,
//                      i.e. the additions happen in exactly the reference's order.
//                      An opposite-spin double through two singles is one integral load:
//                      sign_a * sign_b * V(v1,o1,v2,o2)  (:140-151).
//   meta = o | v << 8 | (sign < 0) << 16 | dead << 17
// dead (h_thresh > 0 only): a same-spin double with |value| <= thr (dropped outright), or an
// alpha single whose integrals V(v,o,*,*) are all <= thr, i.e. every opposite-spin double
// through it would be dropped, so only its k' = k element is enumerated.
__global__ void k_dead_ov(IntsView I, double thr, unsigned char* __restrict__ dead /* n*n */) {
  const int n = I.n;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n * n) return;
  const size_t n2 = size_t(n) * n;
  bool all_small = thr > 0.0;
  for (size_t pq = 0; pq < n2 && all_small; ++pq) all_small = fabs(I.V[t + pq * n2]) <= thr;
  dead[t] = all_small ? 1 : 0;
}
__device__ __forceinline__ double single_lead_sum(const IntsView& I, uint64_t occ_same, unsigned o,
                                                  unsigned v) {
  const size_t n = I.n;
  double h = ldg(I.T + v + o * n);
  const double* G = I.G + v * n + o * n * n;
  for (uint64_t s = occ_same; s; s &= s - 1) h += ldg(G + lsb64(s));
  return h;
}
// one warp per string: meta / val of every adjacency entry
__global__ void __launch_bounds__(256)
k_pair_meta(IntsView I, const uint64_t* __restrict__ str, int32_t nstr,
            const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj, double thr,
            const unsigned char* __restrict__ dead_ov, uint32_t* __restrict__ meta,
            double* __restrict__ val) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nstr) return;
  const uint64_t s = str[r];
  for (int64_t e = adj_ptr[r] + lane; e < adj_ptr[r + 1]; e += 32) {
    const uint32_t pk = adj[e];
    const int64_t r2 = pk >> 2;
    const int dc = int(pk & 3u);
    const uint64_t s2 = str[r2];
    const uint64_t bra = r < r2 ? s : s2, ket = r < r2 ? s2 : s;
    uint32_t m = 0;
    double v = 0.;
    if (dc == 2) {
      v = me4(I, bra, ket, bra ^ ket);
      if (thr > 0.0 && !(fabs(v) > thr)) m = 1u << 17;
    } else if (dc == 1) {
      unsigned o, vv;
      double sg;
      sx_sign_indices(bra, ket, bra ^ ket, o, vv, sg);
      const bool dd = dead_ov && dead_ov[vv + o * I.n] != 0;
      m = o | (vv << 8) | (sg < 0 ? (1u << 16) : 0u) | (dd ? (1u << 17) : 0u);
      v = single_lead_sum(I, bra, o, vv);
    }
    meta[e] = m;
    val[e] = v;
  }
}

// Compacted alpha-run adjacency: entries that enumerate at least one column, in ascending run
// order, plus the list of the run's single excitations (leading sum + meta) in the same order.
struct __align__(8) ARec {
  uint32_t r2t;   // run index << 2 | kind: 0 self, 1 live single, 2 unit (double, or dead single)
  uint32_t meta;  // o | v << 8 | sign << 16 | is_single << 18   (singles)
};
template <bool FILL>
__global__ void __launch_bounds__(256)
k_adj_compact(int32_t nstr, const int64_t* __restrict__ adj_ptr, const uint32_t* __restrict__ adj,
              const uint32_t* __restrict__ meta, const double* __restrict__ val,
              int32_t* __restrict__ ecnt, int32_t* __restrict__ scnt, int32_t* __restrict__ run_cnt,
              const int64_t* __restrict__ cptr, ARec* __restrict__ crec, double* __restrict__ cval,
              const int64_t* __restrict__ sptr, double* __restrict__ slead, uint32_t* __restrict__ smeta) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= nstr) return;
  const unsigned lt = (1u << lane) - 1u;
  int64_t out = FILL ? cptr[r] : 0, sout = FILL ? sptr[r] : 0;
  uint32_t nself = 0, nlive = 0, nunit = 0, nsing = 0;
  const int64_t e0 = adj_ptr[r], e1 = adj_ptr[r + 1];
  for (int64_t eb = e0; eb < e1; eb += 32) {
    const int64_t e = eb + lane;
    const bool ev = e < e1;
    const uint32_t pk = ev ? adj[e] : 0u;
    const uint32_t m = ev ? meta[e] : 0u;
    const int dc = int(pk & 3u);
    const bool dead = (m >> 17) & 1u;
    const bool is_self = ev && dc == 0, is_live = ev && dc == 1 && !dead;
    const bool is_unit = ev && ((dc == 1 && dead) || (dc == 2 && !dead));
    const bool is_sing = ev && dc == 1;
    const unsigned ms = __ballot_sync(0xffffffffu, is_self);
    const unsigned ml = __ballot_sync(0xffffffffu, is_live);
    const unsigned mu = __ballot_sync(0xffffffffu, is_unit);
    const unsigned mg = __ballot_sync(0xffffffffu, is_sing);
    if (FILL && (is_self || is_live || is_unit)) {
      ARec rec;
      rec.r2t = (pk & ~3u) | (is_self ? 0u : (is_live ? 1u : 2u));
      rec.meta = (m & 0x1FFFFu) | (dc == 1 ? (1u << 18) : 0u);
      const int64_t pos = out + __popc((ms | ml | mu) & lt);
      crec[pos] = rec;
      cval[pos] = val[e];
    }
    if (FILL && is_sing) {
      const int64_t pos = sout + __popc(mg & lt);
      slead[pos] = val[e];
      smeta[pos] = m & 0x1FFFFu;
    }
    out += __popc(ms | ml | mu);
    sout += __popc(mg);
    nself += __popc(ms);
    nlive += __popc(ml);
    nunit += __popc(mu);
    nsing += __popc(mg);
  }
  if (!FILL && lane == 0) {
    ecnt[r] = int32_t(nself + nlive + nunit);
    scnt[r] = int32_t(nsing);
    run_cnt[4 * r] = int32_t(nself);
    run_cnt[4 * r + 1] = int32_t(nlive);
    run_cnt[4 * r + 2] = int32_t(nunit);
    run_cnt[4 * r + 3] = int32_t(nsing);
  }
}

// Beta-side record of list B2(k) (distance <= 2): column offset k2 and, packed, the position of
// the (particle, hole) pair inside an integral slice in both orientations, sign and flags
//   pk = (v2 + o2 n) | (o2 + v2 n) << 12 | sign << 24 | is_self << 25 | (k2 > k) << 26
// (first offset: bra = lower template index; second: orientation swapped)
struct __align__(8) B2Rec {
  uint32_t k2;
  uint32_t pk;
};
__global__ void __launch_bounds__(256)
k_beta_rec(int n, int32_t nstr, const int64_t* __restrict__ b2_ptr, const uint32_t* __restrict__ b2,
           const uint32_t* __restrict__ b2_meta, B2Rec* __restrict__ rec) {
  const int lane = threadIdx.x & 31;
  const int64_t k = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (k >= nstr) return;
  for (int64_t e = b2_ptr[k] + lane; e < b2_ptr[k + 1]; e += 32) {
    const uint32_t pk = b2[e], m = b2_meta[e];
    const uint32_t k2 = pk >> 2;
    const uint32_t o = m & 0xFFu, v = (m >> 8) & 0xFFu;
    B2Rec r;
    r.k2 = k2;
    r.pk = (v + o * n) | ((o + v * n) << 12) | (((m >> 16) & 1u) << 24) |
           ((pk & 3u) == 0 ? (1u << 25) : 0u) | (k2 > uint32_t(k) ? (1u << 26) : 0u);
    rec[e] = r;
  }
}

struct ProdArgs {
  IntsView I;
  const uint64_t* run_alpha;  // R
  const uint64_t* tmpl_beta;  // Nb
  const int64_t* cptr;        // compacted alpha-run adjacency
  const ARec* crec;
  const double* cval;
  const int64_t* sptr;        // single excitations of every run, in adjacency order
  const double* slead;
  const uint32_t* smeta;
  const int32_t* run_cnt;     // 4 per run: self, live singles, unit, singles
  const int64_t* b2_ptr;      // beta adjacency, distance <= 2
  const B2Rec* b2rec;
  const uint32_t* b2_meta;
  const double* b2_val;
  const int64_t* b4_ptr;      // distance <= 4
  const uint32_t* b4;
  const double* b4_val;
  const double* diag;         // <D|H|D> of every row of the block
  int64_t nb;
  int64_t row_begin;
  int64_t nrows;
  double thr;
  int smem_a;                 // doubles of per-warp scratch for alpha singles
  int smem_b;                 // ... and beta singles
  int smem_r;                 // opposite-spin records per row (smem_b rounded up to the group width)
  int nslice_max;             // integral slices staged per CTA (singles per run, upper bound)
  int32_t* row_cnt;           // structural count (count kernel) / surviving count (fill kernel)
  const int64_t* rowptr;      // slot offsets of the fill kernel
  int32_t* colind;
  double* nzval;
};

__global__ void k_prod_struct_count(const ProdArgs A) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= A.nrows) return;
  const int64_t i = A.row_begin + row;
  const int64_t r = i / A.nb, k = i % A.nb;
  const int64_t l2 = A.b2_ptr[k + 1] - A.b2_ptr[k], l4 = A.b4_ptr[k + 1] - A.b4_ptr[k];
  const int32_t* d = A.run_cnt + 4 * r;
  const int64_t c = int64_t(d[0]) * l4 + int64_t(d[1]) * l2 + int64_t(d[2]);
  A.row_cnt[row] = int32_t(c);
}
// diagonal elements, one thread per row (matrix_elements.hpp:203-230)
__global__ void k_row_diag(const ProdArgs A, double* __restrict__ diag) {
  const int64_t row = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (row >= A.nrows) return;
  const int64_t i = A.row_begin + row;
  diag[row] = me_diag(A.I, A.run_alpha[i / A.nb], A.tmpl_beta[i % A.nb]);
}

__device__ __forceinline__ double flip_sign_if(double v, unsigned neg) {
  return __hiloint2double(__double2hiint(v) ^ int((neg & 1u) << 31), __double2loint(v));
}

// One CTA per (alpha run, chunk of template strings); G lanes per row, 32 / G rows per warp.
// The CTA first stages, with bulk (TMA) copies signalled on an mbarrier, the integral slice
// Vt(v1 o1 | . .) of every single excitation of its run -- the only integrals the opposite-spin
// doubles of these rows read (SLICES; for large n they stay in global memory / L2). Every warp
// then walks the compacted adjacency of the run 32 entries at a time; inside a window, runs of
// unit entries are emitted with one lane per entry and every list entry (self: B4(k), live
// single: B2(k)) with one lane per list element, each lane group doing so for its own row. All
// rows of a run share the adjacency, so control flow is warp-uniform; output positions are
// ascending, so survivors are written in order with a ballot prefix. Narrow groups keep the
// lanes busy when the lists are short (CAS(12,12): |B2| = 37, unit runs of ~6).
// Single-excitation elements (leading sum + V_red terms of the other spin, added in ascending
// orbital order) are evaluated once per row with full lanes into shared memory, next to the
// row's B2 records.
#ifndef B2CI_PROD_PW
#define B2CI_PROD_PW 8       // measured: 8 warps x 4 CTAs/SM beats 16 x 2 by 2 % and halves the tail at 8 GPUs
#endif
constexpr int PW = B2CI_PROD_PW;  // warps per CTA of the product kernel
#ifndef B2CI_PROD_UH
#define B2CI_PROD_UH 1      // opposite-spin iterations whose shared-memory loads are grouped
#endif
#ifndef B2CI_PROD_U4
#define B2CI_PROD_U4 4      // B4-list iterations whose global loads are grouped
#endif
#ifndef B2CI_PROD_MINB
#define B2CI_PROD_MINB 4    // resident CTAs per SM the register allocation must allow
#endif
template <int G>
struct GroupOut {
  int32_t* ci;      // colind of this row (slot base), kept as an opaque 64-bit register pair
  double* nz;       // nzval of this row
  int rel;          // elements written so far
  unsigned ltmask;  // lanes of this group below this lane (warp-wide bit positions)
  unsigned gmask;   // lanes of this group
  int lig;          // lane index inside the group
  int ndrop;        // DENSE: elements this lane wrote that fail the threshold
};
// DENSE (lists whose integrals are dense, i.e. next to nothing fails |h| > thr): every element goes to
// its STRUCTURAL position -- the active lanes of a step are a prefix of the group, `nact` of them --
// so a step is two stores and an add: no ballot, no prefix count, no branch. Elements that fail the
// threshold are only counted; if the build finds any, one filtering pass packs the rows afterwards.
template <bool EVAL, int G, bool DENSE>
__device__ __forceinline__ void emit(GroupOut<G>& O, double thr, bool act, int nact, int32_t j, double v) {
  if (DENSE) {
    if (act) {
      const int pos = O.rel + O.lig;
      O.ci[pos] = j;
      O.nz[pos] = v;
      if (EVAL && !(fabs(v) > thr)) ++O.ndrop;
    }
    O.rel += nact;
    return;
  }
  const bool keep = act && (EVAL ? (fabs(v) > thr) : true);
  const unsigned m = __ballot_sync(0xffffffffu, keep);
  if (m == 0xffffffffu) {  // every lane of the warp survives (the common case): no prefix count
    const int pos = O.rel + O.lig;
    O.ci[pos] = j;
    O.nz[pos] = v;
    O.rel += G;
    return;
  }
  if (keep) {
    const int pos = O.rel + __popc(m & O.ltmask);
    O.ci[pos] = j;
    O.nz[pos] = v;
  }
  O.rel += __popc(m & O.gmask);
}
// opposite-spin record of one (row, B2 entry), resolved for both bra/ket orientations:
//   w = off(r < r2) | off(r > r2) << 12 | is_self << 30 | sign << 31; k2 = ~0 marks padding
struct __align__(8) OsRec {
  uint32_t k2;
  uint32_t w;
};

template <bool EVAL, int G, bool SLICES, bool DENSE = false>
__global__ void __launch_bounds__(PW * 32, B2CI_PROD_MINB)
k_rows_product(const ProdArgs A) {
  constexpr int RPW = 32 / G, RPC = PW * RPW;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbarrier, 16 B][slices][per row: singles alpha | singles beta][per row: B2 records]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* slices = reinterpret_cast<double*>(smem_raw + 16);
  double* rows_d = slices + (SLICES ? size_t(A.nslice_max) * A.I.n2p : 0);
  OsRec* rows_b = reinterpret_cast<OsRec*>(rows_d + size_t(RPC) * (A.smem_a + A.smem_b));
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int gid = lane / G, l = lane % G;
  const int64_t cpr = (A.nb + RPC - 1) / RPC;  // chunks per run
  const uint32_t r = uint32_t(A.row_begin / A.nb + blockIdx.x / cpr);
  const int64_t kk = (blockIdx.x % cpr) * RPC + w * RPW + gid;
  const int64_t i = int64_t(r) * A.nb + kk;
  const bool rowvalid = kk < A.nb && i >= A.row_begin && i < A.row_begin + A.nrows;
  const int64_t row = rowvalid ? i - A.row_begin : 0;
  const uint64_t ai = A.run_alpha[r];
  if (ai == 0) {  // alpha-empty determinants are skipped (uniform over the CTA)
    if (rowvalid && l == 0) A.row_cnt[row] = 0;
    return;
  }
  const int n = A.I.n;
  const size_t n2 = size_t(n) * n;
  const int64_t sp = A.sptr[r];
  const int ns = int(A.sptr[r + 1] - sp);
  if (SLICES) {
    if (threadIdx.x == 0) mbar_init(bar, 1);
    __syncthreads();
    if (w == 0) {
      const uint32_t bytes = uint32_t(A.I.n2p) * 8u;
      int mine = 0;
      for (int s = lane; s < ns; s += 32) mine += ((A.smeta[sp + s] >> 17) & 1u) ? 0 : 1;
      const int total = __reduce_add_sync(0xffffffffu, mine);
      if (lane == 0) mbar_arrive_expect_tx(bar, uint32_t(total) * bytes);
      __syncwarp();
      for (int s = lane; s < ns; s += 32) {
        const uint32_t m = A.smeta[sp + s];
        if ((m >> 17) & 1u) continue;  // no opposite-spin double through this single survives
        const size_t pq = ((m >> 8) & 0xFFu) + size_t(m & 0xFFu) * n;
        bulk_g2s(slices + size_t(s) * A.I.n2p, A.I.Vt + pq * A.I.n2p, bytes, bar);
      }
    }
  }
  const bool warp_active = __any_sync(0xffffffffu, rowvalid);
  const uint32_t k = rowvalid ? uint32_t(kk) : 0u;
  const uint32_t nb = uint32_t(A.nb);
  double* sa = rows_d + size_t(w * RPW + gid) * (A.smem_a + A.smem_b);
  double* sb = sa + A.smem_a;
  OsRec* sb2 = rows_b + size_t(w * RPW + gid) * A.smem_r;
  const uint64_t bi = A.tmpl_beta[k];
  const int64_t b2s = A.b2_ptr[k], b4s = A.b4_ptr[k];
  const int len2 = int(A.b2_ptr[k + 1] - b2s), len4 = int(A.b4_ptr[k + 1] - b4s);
  const int len2max = G == 32 ? len2 : __reduce_max_sync(0xffffffffu, rowvalid ? len2 : 0);
  const int len4max = G == 32 ? len4 : __reduce_max_sync(0xffffffffu, rowvalid ? len4 : 0);
  const int64_t out0 = A.rowptr[row];
  const double dgv = A.diag[row];  // fetched early: its latency hides behind the staging phase
  GroupOut<G> O;
  O.ci = A.colind + out0;
  O.nz = A.nzval + out0;
  // keep the two row pointers as plain registers: stores become one IMAD.WIDE + STG each
  asm volatile("" : "+l"(O.ci), "+l"(O.nz));
  __builtin_assume(__isGlobal(O.ci));
  __builtin_assume(__isGlobal(O.nz));
  O.rel = 0;
  O.ndrop = 0;
  const unsigned ltg = (1u << l) - 1u;   // lanes of the group below this lane (group-relative)
  const int gshift = gid * G;
  O.ltmask = ltg << gshift;
  O.lig = l;
  O.gmask = (G == 32 ? 0xffffffffu : ((1u << (G & 31)) - 1u)) << gshift;
  const unsigned lt = (1u << lane) - 1u;
  const int len2pad = (len2max + G - 1) / G * G;  // records per row incl. padding (<= smem_r)
  const int len2r = rowvalid ? len2 : 0;
  if (warp_active) {
    // ---- single-excitation elements and B2 records of this row
    for (int s = l; s < ns; s += G) {
      const uint32_t m = A.smeta[sp + s];
      double h = A.slead[sp + s];
      const double* Vr = A.I.Vr + ((m >> 8) & 0xFFu) * n + (m & 0xFFu) * n2;
      for (uint64_t q = bi; q; q &= q - 1) h += ldg(Vr + lsb64(q));
      sa[s] = flip_sign_if(h, m >> 16);
    }
    for (int t = l; t < (rowvalid ? len2 : 0); t += G) {
      const uint32_t m = A.b2_meta[b2s + t];
      double h = A.b2_val[b2s + t];
      const double* Vr = A.I.Vr + ((m >> 8) & 0xFFu) * n + (m & 0xFFu) * n2;
      for (uint64_t q = ai; q; q &= q - 1) h += ldg(Vr + lsb64(q));
      sb[t] = flip_sign_if(h, m >> 16);  // the self slot is never read
      const B2Rec br = A.b2rec[b2s + t];
      const uint32_t o0 = br.pk & 0xFFFu, o1 = (br.pk >> 12) & 0xFFFu;
      const bool up = (br.pk >> 26) & 1u;  // k2 > k
      OsRec orec;
      orec.k2 = br.k2;
      // row is the bra iff r < r2; the stored pair has bra = lower template index
      orec.w = (up ? o0 : o1) | ((up ? o1 : o0) << 12) | (((br.pk >> 25) & 1u) << 30) | (((br.pk >> 24) & 1u) << 31);
      sb2[t] = orec;
    }
    for (int t = (rowvalid ? len2 : 0) + l; t < len2pad; t += G) {
      OsRec orec;
      orec.k2 = 0xFFFFFFFFu; orec.w = 0u;
      sb2[t] = orec;
    }
    __syncwarp();
  }
  // Short lists (the usual full-CI shapes): every lane keeps the NR records it will ever touch
  // in registers, one packed word each --
  //   off(r < r2) | off(r > r2) << 8 | is_self << 16 | sign << 17 | k2 << 18, ~0 = padding
  // -- so the inner loop reads shared memory only for the integral itself.
  constexpr int NR = 3;
  const bool regrec = SLICES && len2pad <= NR * G && A.I.n2p <= 256 && A.nb < 16383;
  uint32_t rr[NR];
#pragma unroll
  for (int u = 0; u < NR; ++u) rr[u] = 0xFFFFFFFFu;
  if (warp_active && regrec) {
#pragma unroll
    for (int u = 0; u < NR; ++u) {
      if (u * G < len2pad) {
        const OsRec o = sb2[u * G + l];
        if (o.k2 != 0xFFFFFFFFu)
          rr[u] = (o.w & 0xFFu) | (((o.w >> 12) & 0xFFu) << 8) | (((o.w >> 30) & 1u) << 16) |
                  ((o.w >> 31) << 17) | (o.k2 << 18);
      }
    }
  }
  if (SLICES) mbar_wait(bar, 0);
  if (!warp_active) return;
  const int64_t E0 = A.cptr[r], E1 = A.cptr[r + 1];
  int sord0 = 0;  // singles before the current window
  // adjacency windows are fetched one window ahead of their use
  ARec rec_n;
  rec_n.r2t = 0; rec_n.meta = 0;
  double cv_n = 0.;
  if (E0 + lane < E1) { rec_n = A.crec[E0 + lane]; cv_n = A.cval[E0 + lane]; }
  for (int64_t eb = E0; eb < E1; eb += 32) {
    const int nv = int(min(int64_t(32), E1 - eb));
    const bool ev = lane < nv;
    const ARec rec = rec_n;
    const double cv = cv_n;
    rec_n.r2t = 0; rec_n.meta = 0;
    cv_n = 0.;
    if (eb + 32 + lane < E1) { rec_n = A.crec[eb + 32 + lane]; cv_n = A.cval[eb + 32 + lane]; }
    const bool sing_l = ev && ((rec.meta >> 18) & 1u);
    const unsigned lm = __ballot_sync(0xffffffffu, ev && (rec.r2t & 3u) != 2u);  // list entries
    const unsigned gm = __ballot_sync(0xffffffffu, sing_l);                      // singles
    const int sord_l = sord0 + __popc(gm & lt);
    int cur = 0;
    while (cur < nv) {
      const unsigned rem = lm & ~((1u << cur) - 1u);
      const int f = rem ? (__ffs(rem) - 1) : nv;
      if (f > cur) {
        // unit entries cur .. f-1: one element each, column (r2, k)
        const unsigned runmask = (f - cur >= 32 ? 0xffffffffu : ((1u << (f - cur)) - 1u)) << cur;
        const bool any_sing = (gm & runmask) != 0u;
        for (int u0 = cur; u0 < f; u0 += G) {
          const int src = u0 + l;
          const bool act = rowvalid && src < f;
          const uint32_t r2t = __shfl_sync(0xffffffffu, rec.r2t, src & 31);
          double v = __shfl_sync(0xffffffffu, cv, src & 31);
          if (any_sing) {
            const int so = __shfl_sync(0xffffffffu, sing_l ? sord_l : -1, src & 31);
            if (so >= 0) v = sa[so];
          }
          emit<EVAL, G, DENSE>(O, A.thr, act, rowvalid ? min(G, f - u0) : 0, int32_t((r2t >> 2) * nb + k), v);
        }
      }
      if (f >= nv) break;
      const uint32_t r2t = __shfl_sync(0xffffffffu, rec.r2t, f);
      const uint32_t r2 = r2t >> 2;
      if ((r2t & 3u) == 1u) {
        // live alpha single x B2(k): opposite-spin doubles + the same-beta single
        const uint32_t am = __shfl_sync(0xffffffffu, rec.meta, f);
        const int so = __shfl_sync(0xffffffffu, sord_l, f);
        const double* Va = SLICES ? slices + size_t(so) * A.I.n2p
                                  : A.I.Vt + (((am >> 8) & 0xFFu) + size_t(am & 0xFFu) * n) * A.I.n2p;
        const double vself = sa[so];
        // the row determinant is the bra iff r < r2 (orientation resolved when the records were staged)
        const uint32_t sh = r < r2 ? 0u : 12u;
        const uint32_t asign = ((am >> 16) & 1u) << 31;
        const uint32_t base = r2 * nb;
        if (regrec) {
          const uint32_t sh8 = r < r2 ? 0u : 8u;
#pragma unroll
          for (int u = 0; u < NR; ++u) {
            if (u * G < len2pad) {  // warp-uniform
              const uint32_t w = rr[u];
              const double vr = Va[(w >> sh8) & 0xFFu];
              double v = __hiloint2double(__double2hiint(vr) ^ int(((w << 14) & 0x80000000u) ^ asign),
                                          __double2loint(vr));
              if (w & (1u << 16)) v = vself;
              emit<EVAL, G, DENSE>(O, A.thr, w != 0xFFFFFFFFu, min(G, max(0, len2r - u * G)), int32_t(base + (w >> 18)), v);
            }
          }
        } else {
          for (int t0 = 0; t0 < len2pad; t0 += G) {
            const OsRec br = sb2[t0 + l];
            const bool act = br.k2 != 0xFFFFFFFFu;
            const uint32_t off = (br.w >> sh) & 0xFFFu;
            const double vr = SLICES ? Va[off] : ldg(Va + off);
            double v = __hiloint2double(__double2hiint(vr) ^ int((br.w & 0x80000000u) ^ asign), __double2loint(vr));
            if (br.w & (1u << 30)) v = vself;
            emit<EVAL, G, DENSE>(O, A.thr, act, min(G, max(0, len2r - t0)), int32_t(base + br.k2), v);
          }
        }
      } else {
        // same alpha string x B4(k): diagonal, beta singles, beta doubles
        const uint32_t base = r * nb;
        int t2run = 0;  // position in B2(k) of the next entry at distance <= 2
        // the list lives in global memory (L2): issue the loads of U iterations together, value
        // and descriptor side by side, so one latency is paid per U * G elements
        constexpr int U = B2CI_PROD_U4;
        for (int c0 = 0; c0 < len4max; c0 += U * G) {
          uint32_t bpk_u[U];
          double bv_u[U];
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int64_t e4 = b4s + min(c0 + u * G + l, len4 - 1);
            bpk_u[u] = __ldg(A.b4 + e4);
            bv_u[u] = ldg(A.b4_val + e4);
          }
#pragma unroll
          for (int u = 0; u < U; ++u) {
            const int t = c0 + u * G + l;
            if (c0 + u * G < len4max) {  // warp-uniform
              const bool act = rowvalid && t < len4;
              const uint32_t bpk = bpk_u[u];
              const int db = int(bpk & 3u);
              const unsigned g01 = __ballot_sync(0xffffffffu, act && db <= 1) & O.gmask;
              double v;
              if (db == 2) v = bv_u[u];
              else if (db == 1) v = act ? sb[t2run + __popc(g01 & O.ltmask)] : 0.;
              else v = dgv;
              t2run += __popc(g01);
              emit<EVAL, G, DENSE>(O, A.thr, act, rowvalid ? min(G, max(0, len4 - (c0 + u * G))) : 0, int32_t(base + (bpk >> 2)), v);
            }
          }
        }
      }
      cur = f + 1;
    }
    sord0 += __popc(gm);
  }
  if (DENSE && EVAL) {
    int nd = O.ndrop;
#pragma unroll
    for (int d = G / 2; d > 0; d >>= 1) nd += __shfl_xor_sync(0xffffffffu, nd, d);
    O.rel -= nd;
  }
  if (rowvalid && l == 0) A.row_cnt[row] = O.rel;
}

// ------------------------------------------------------------------ dense fill of uniform lists
// Full-CI lists are uniform: every beta string has the same number of neighbours at distance <= 2
// (len2) and <= 4 (len4), so every row of an alpha run has the SAME output layout -- the segment of
// adjacency entry e starts at the same offset in all of them. k_rows_dense uses that: per CTA (alpha
// run r, 32 rows) one warp turns the run's adjacency into two shared-memory tables (live singles,
// unit entries) with their segment offsets, and then each warp fills one row at a time in three
// branch-free sweeps with all 32 lanes busy:
//   singles sweep : flat index f over (live single s, B2 record t): one table read, one record
//                   read, one integral from the TMA-staged slice, sign flip, two stores
//   unit sweep    : one entry per lane (same-spin alpha doubles, dead singles)
//   self sweep    : the B4(k) list (diagonal, beta singles, beta doubles)
// Elements go to their structural positions; |h| <= thr is only counted (see emit<.., DENSE>), so
// the inner loop has no ballot, no prefix sum and no data-dependent branch. Used when the integrals
// are dense (< 2 % of V fails the threshold) and the slices fit; everything else keeps
// k_rows_product.
// element stores of the dense fill: address = base + pos (32-bit, unsigned) as ONE mad.wide + st
__device__ __forceinline__ void st_elem(int32_t* ci, double* nz, uint32_t pos, int32_t col, double v) {
  asm volatile(
      "{\n\t.reg .u64 a, b;\n\t"
      "mad.wide.u32 a, %2, 4, %0;\n\t"
      "mad.wide.u32 b, %2, 8, %1;\n\t"
      "st.global.u32 [a], %3;\n\t"
      "st.global.f64 [b], %4;\n\t}"
      ::"l"(ci), "l"(nz), "r"(pos), "r"(col), "d"(v));
}
struct DenseArgs {
  ProdArgs P;
  int len2, len4;  // uniform beta adjacency lengths
  int tab_max;     // table entries per run (capacity)
  int per_row;     // doubles of (sa | sb) per row in `sab`
  uint32_t* desc;  // [nruns_blk][desc_max]: per output position (outside the self segment) entry << 4 | record t << 16
  int tab_off;     // byte offset of the table in the CTA's shared memory
  int rec_stride;  // records per beta string: len2 + 2 (neutral record, sa record)
  int desc_max;    // row length rounded up to 4 (16-byte rows for the bulk copy)
  uint4* tab;      // [nruns_blk][tab_max]: live singles first, unit entries behind them
  int* tab_cnt;    // [nruns_blk][4]: live singles, units, offset of the self segment, row length
  uint2* rec;      // [nb][len2]: B2 records in the sweep's form (+ padding)
  double* sab;     // [nruns_blk * nb][per_row]: single-excitation elements of every row (alpha | beta)
  int64_t run0;    // first alpha run of the row block
};
#ifndef B2CI_DENSE_WARPS
#define B2CI_DENSE_WARPS 8
#endif
#ifndef B2CI_DENSE_ROWS
#define B2CI_DENSE_ROWS 16
#endif
#ifndef B2CI_DENSE_MINB
#define B2CI_DENSE_MINB 3
#endif
constexpr int DW = B2CI_DENSE_WARPS;     // warps per CTA
constexpr int DROWS = B2CI_DENSE_ROWS;   // rows per CTA

// ---- pre-passes of the dense fill (each O(rows x singles) or smaller, fully parallel) -------------------
// tables of one alpha run: where every adjacency entry's segment starts in a row (the same in all rows)
__global__ void __launch_bounds__(128)
k_dense_tables(const DenseArgs D, int64_t nruns_blk) {
  const ProdArgs& A = D.P;
  const int lane = threadIdx.x & 31;
  const int64_t rl = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (rl >= nruns_blk) return;
  const uint32_t r = uint32_t(D.run0 + rl);
  const uint32_t nb = uint32_t(A.nb);
  uint4* __restrict__ tab = D.tab + size_t(rl) * D.tab_max;
  int* __restrict__ cnt = D.tab_cnt + size_t(rl) * 4;
  const int64_t E0 = A.cptr[r], E1 = A.cptr[r + 1];
  const unsigned lt = (1u << lane) - 1u;
  int off0 = 0, nseg0 = 0, nunit0 = 0, sord0 = 0, nseg_total = 0;
  for (int64_t eb = E0; eb < E1; eb += 32) {  // number of live singles (units are stored behind them)
    const bool ev = eb + lane < E1;
    const uint32_t r2t = ev ? A.crec[eb + lane].r2t : 2u;
    nseg_total += __popc(__ballot_sync(0xffffffffu, ev && (r2t & 3u) == 1u));
  }
  int self_off = 0;
  for (int64_t eb = E0; eb < E1; eb += 32) {
    const bool ev = eb + lane < E1;
    ARec a; a.r2t = 2u; a.meta = 0u;
    double cv = 0.;
    if (ev) { a = A.crec[eb + lane]; cv = A.cval[eb + lane]; }
    const int kind = int(a.r2t & 3u);
    const bool is_sing = ev && ((a.meta >> 18) & 1u);
    const int len = !ev ? 0 : (kind == 0 ? D.len4 : (kind == 1 ? D.len2 : 1));
    int incl = len;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += v;
    }
    const int off = off0 + incl - len;
    const unsigned m1 = __ballot_sync(0xffffffffu, ev && kind == 1);
    const unsigned m2 = __ballot_sync(0xffffffffu, ev && kind == 2);
    const unsigned m0 = __ballot_sync(0xffffffffu, ev && kind == 0);
    const unsigned mg = __ballot_sync(0xffffffffu, is_sing);
    const int so = sord0 + __popc(mg & lt);
    const uint32_t r2 = a.r2t >> 2;
    uint32_t* __restrict__ desc = D.desc + size_t(rl) * D.desc_max;
    // every entry has the form {column base, -, byte offset of the value's home in shared memory, w}: pass 1
    // reads value = *(home + record offset) and flips its sign by (record ^ w) -- the same code for all kinds
    if (ev && kind == 1) {
      // live single: home = its integral slice; w = record shift (0: the row is the bra, r < r2; 15: the
      // ket) | slice index << 8 | sign of the alpha single << 31
      const int e = nseg0 + __popc(m1 & lt);
      tab[e] = make_uint4(r2 * nb, 0u, uint32_t(32 + size_t(so) * A.I.n2p * 8),
                          (r < r2 ? 0u : 15u) | (uint32_t(so) << 8) | (((a.meta >> 16) & 1u) << 31));
      for (int t = 0; t < D.len2; ++t) desc[off + t] = (uint32_t(e) << 4) | (uint32_t(t) << 16);
    } else if (ev && kind == 2) {
      // unit entry, two slots: {column base, -, home = the second slot, 0} {value}; it pairs with the row's
      // neutral record (t = len2: column k, offset 0, no sign). A dead single's value is the row's sa[so]:
      // slice index in w, paired with the record t = len2 + 1 whose same-beta flag selects sa.
      const int e = nseg_total + 2 * (nunit0 + __popc(m2 & lt));
      tab[e] = make_uint4(r2 * nb, 0u, uint32_t(D.tab_off + size_t(e + 1) * 16), is_sing ? uint32_t(so) << 8 : 0u);
      tab[e + 1] = make_uint4(uint32_t(__double2loint(cv)), uint32_t(__double2hiint(cv)), 0u, 0u);
      desc[off] = (uint32_t(e) << 4) | (uint32_t(D.len2 + (is_sing ? 1 : 0)) << 16);
    }
    if (m0) self_off = __shfl_sync(0xffffffffu, off, __ffs(m0) - 1);
    off0 += __shfl_sync(0xffffffffu, incl, 31);
    nseg0 += __popc(m1);
    nunit0 += __popc(m2);
    sord0 += __popc(mg);
  }
  if (lane == 0) { cnt[0] = nseg0; cnt[1] = nunit0; cnt[2] = self_off; cnt[3] = off0; }
}
// B2 records of every beta string in the form the singles sweep reads:
//   x = k2, y = 8 * offset when the row is the bra | 8 * offset when it is the ket << 15 | is_self << 30 | sign << 31
__global__ void k_dense_rec(const DenseArgs D) {
  const ProdArgs& A = D.P;
  const int64_t idx = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (idx >= A.nb * D.rec_stride) return;
  const int64_t k = idx / D.rec_stride;
  const int t = int(idx - k * D.rec_stride);
  if (t >= D.len2) {  // the neutral record of unit entries / the record that selects sa[so]
    D.rec[idx] = make_uint2(uint32_t(k), t == D.len2 ? 0u : (1u << 30));
    return;
  }
  const B2Rec br = A.b2rec[k * D.len2 + t];  // uniform lists: b2_ptr[k] == k * len2
  const uint32_t o0 = br.pk & 0xFFFu, o1 = (br.pk >> 12) & 0xFFFu;
  const bool up = (br.pk >> 26) & 1u;  // k2 > k
  D.rec[idx] = make_uint2(br.k2, ((up ? o0 : o1) << 3) | ((up ? o1 : o0) << 18) | (((br.pk >> 25) & 1u) << 30) |
                                     (((br.pk >> 24) & 1u) << 31));
}
// single-excitation elements of every row: leading sum + the other spin's V_red terms in ascending orbital
// order (the reference's order of additions, matrix_elements.hpp:176-186)
constexpr int SAB_ROWS = 64;  // rows of one alpha run per CTA
// occupied orbitals of a string, ascending, into p[]; returns their number (warp-uniform input: no divergence)
template <typename S, int NT = 0>
__device__ __forceinline__ int occupied_positions(S bits, int (&p)[16], S& rest) {
  int n = 0;
#pragma unroll
  for (int u = 0; u < (NT > 0 ? NT : 16); ++u) {
    p[u] = 0;
    if (bits) {
      p[u] = (sizeof(S) == 4 ? __ffs(int(bits)) : __ffsll((long long)bits)) - 1;
      bits &= bits - 1;
      ++n;
    }
  }
  rest = bits;  // strings with more than 16 electrons: the remainder goes through the generic loop
  return n;
}
// h + the NT terms Vr[p[0]], Vr[p[1]], ... added one after the other (the reference's order of additions)
template <int NT>
__device__ __forceinline__ double add_terms(double h, const double* __restrict__ Vr, const int (&p)[16]) {
#pragma unroll
  for (int u = 0; u < NT; ++u) h += Vr[p[u]];
  return h;
}
// PART 0: beta-single elements sb(k, t) of all rows of the CTA (positions: the run's alpha string);
// PART 1: alpha-single elements sa(k, s), one warp per row (positions: the row's beta string).
// NT = number of electrons of the string that supplies the positions (exact unrolling: the sums are the whole
// cost of the kernel); NT == 0: any count, generic loops.
template <typename S, bool SMEM, int PART, int NT>
__global__ void __launch_bounds__(256)
k_dense_sab(const DenseArgs D, int64_t nrows_tot, unsigned inv_len2) {
  // CTA = (alpha run, 64 consecutive beta strings). Every element is a leading value plus the other spin's
  // V_red terms in ascending orbital order (the reference's order of additions, matrix_elements.hpp:176-186).
  // The orbital positions are the same for all lanes of a step, so they are extracted once and each term is one
  // shared-memory load + one add per lane. The n^3 table of V_red is staged in shared memory per CTA (SMEM,
  // n <= 18): 8-byte gathers all over it run at two sectors per clock through L1.
  extern __shared__ double s_vr[];
  const ProdArgs& A = D.P;
  const int n = A.I.n;
  // shared-memory copy with the n-vectors V_red(., v, o) at an ODD stride: lanes read the same orbital of
  // different (v, o) vectors, and at the natural stride n = 12 those addresses fall on 4 of the 16 8-byte banks
  const int pad = SMEM ? (n | 1) : n;
  if (SMEM) {
    for (int i = threadIdx.x; i < n * n * n; i += blockDim.x) s_vr[(i / n) * pad + (i % n)] = A.I.Vr[i];
    __syncthreads();
  }
  const double* __restrict__ VR = SMEM ? s_vr : A.I.Vr;
  const unsigned per = unsigned(D.per_row);
  const uint32_t r = uint32_t(D.run0 + blockIdx.y);
  const unsigned k0 = blockIdx.x * SAB_ROWS;
  const unsigned nrows_cta = min(unsigned(SAB_ROWS), unsigned(A.nb) - k0);
  const int64_t row_base = int64_t(blockIdx.y) * A.nb + k0;
  if (row_base >= nrows_tot) return;
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const unsigned len2 = unsigned(D.len2);
  int pos[16];
  S rest;
  if (PART == 0) {
    const int na = occupied_positions<S, NT>(S(A.run_alpha[r]), pos, rest);
    for (unsigned f = threadIdx.x; f < nrows_cta * len2; f += blockDim.x) {
      const unsigned j = __umulhi(f, inv_len2);  // f / len2
      const unsigned t = f - j * len2;
      const int64_t e = int64_t(k0 + j) * len2 + t;
      const uint32_t m = A.b2_meta[e];
      double h = A.b2_val[e];
      const double* Vr = VR + (((m >> 8) & 0xFFu) + (m & 0xFFu) * n) * pad;
      if (NT > 0) {
        h = add_terms<NT>(h, Vr, pos);
      } else {
        for (int u = 0; u < 16; ++u)
          if (u < na) h += Vr[pos[u]];
        for (S qq = rest; qq; qq &= qq - 1) h += Vr[(sizeof(S) == 4 ? __ffs(int(qq)) : __ffsll((long long)qq)) - 1];
      }
      D.sab[(row_base + j) * per + A.smem_a + t] = flip_sign_if(h, m >> 16);  // (the self slot is never read)
    }
  } else {
    const int64_t sp = A.sptr[r];
    const unsigned ns = unsigned(A.sptr[r + 1] - sp);
    for (unsigned j = w; j < nrows_cta; j += blockDim.x / 32) {
      const int nbp = occupied_positions<S, NT>(S(A.tmpl_beta[k0 + j]), pos, rest);
      for (unsigned q = lane; q < ns; q += 32) {
        const uint32_t m = A.smeta[sp + q];
        double h = A.slead[sp + q];
        const double* Vr = VR + (((m >> 8) & 0xFFu) + (m & 0xFFu) * n) * pad;
        if (NT > 0) {
          h = add_terms<NT>(h, Vr, pos);
        } else {
          for (int u = 0; u < 16; ++u)
            if (u < nbp) h += Vr[pos[u]];
          for (S qq = rest; qq; qq &= qq - 1) h += Vr[(sizeof(S) == 4 ? __ffs(int(qq)) : __ffsll((long long)qq)) - 1];
        }
        D.sab[(row_base + j) * per + q] = flip_sign_if(h, m >> 16);
      }
    }
  }
}
// launch PART with the exact term count when it is 1 .. 12
template <typename S, bool SMEM, int PART>
void launch_dense_sab(int nt, dim3 gs, size_t smem, cudaStream_t st, const DenseArgs& DA, int64_t nrows_tot, unsigned inv) {
  switch (nt) {
#define B2_SAB_CASE(N) case N: k_dense_sab<S, SMEM, PART, N><<<gs, 256, smem, st>>>(DA, nrows_tot, inv); break;
    B2_SAB_CASE(1) B2_SAB_CASE(2) B2_SAB_CASE(3) B2_SAB_CASE(4) B2_SAB_CASE(5) B2_SAB_CASE(6)
    B2_SAB_CASE(7) B2_SAB_CASE(8) B2_SAB_CASE(9) B2_SAB_CASE(10) B2_SAB_CASE(11) B2_SAB_CASE(12)
#undef B2_SAB_CASE
    default: k_dense_sab<S, SMEM, PART, 0><<<gs, 256, smem, st>>>(DA, nrows_tot, inv); break;
  }
}

__device__ __forceinline__ double lds_f64(uint32_t saddr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(saddr));
  return v;
}
template <bool EVAL>
__global__ void __launch_bounds__(DW * 32, B2CI_DENSE_MINB)
k_rows_dense(const DenseArgs D) {
  const ProdArgs& A = D.P;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  // layout: [mbarrier 16 B][pad 16 B][slices][table: uint4 x tab_max][desc: u32 x desc_max]
  //         [rec: DROWS x len2 x 8 B, padded][sab: DROWS x per_row]
  uint64_t* bar = reinterpret_cast<uint64_t*>(smem_raw);
  double* slices = reinterpret_cast<double*>(smem_raw + 32);
  uint4* tab = reinterpret_cast<uint4*>(slices + size_t(A.nslice_max) * A.I.n2p);
  uint32_t* desc = reinterpret_cast<uint32_t*>(tab + D.tab_max);
  uint2* recs = reinterpret_cast<uint2*>(desc + D.desc_max);
  const int len2 = D.len2, len4 = D.len4;
  const int rstride = D.rec_stride;
  const size_t rec_bytes_max = (size_t(DROWS) * rstride * 8 + 15) & ~size_t(15);
  double* sab = reinterpret_cast<double*>(reinterpret_cast<unsigned char*>(recs) + rec_bytes_max);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  const int64_t cpr = (A.nb + DROWS - 1) / DROWS;  // chunks per run
  const int64_t rl = blockIdx.x / cpr;              // run of the block, local
  const uint32_t r = uint32_t(D.run0 + rl);
  const int64_t k0 = (blockIdx.x % cpr) * DROWS;
  const uint64_t ai = A.run_alpha[r];
  const uint32_t nb = uint32_t(A.nb);
  const int n = A.I.n;
  const int nrows_cta = int(min(int64_t(DROWS), int64_t(nb) - k0));
  if (ai == 0) {  // alpha-empty determinants are skipped (uniform over the CTA)
    for (int j = threadIdx.x; j < nrows_cta; j += blockDim.x) {
      const int64_t i = int64_t(r) * nb + k0 + j;
      if (i >= A.row_begin && i < A.row_begin + A.nrows) A.row_cnt[i - A.row_begin] = 0;
    }
    return;
  }
  if (threadIdx.x == 0) mbar_init(bar, 1);
  __syncthreads();
  if (w == 0) {
    // ---- everything the CTA reads comes in by bulk (TMA) copies: the integral slices of the run's live
    // singles, the run's tables and position descriptors, the B2 records and the single-excitation
    // elements of its rows
    const int64_t sp = A.sptr[r];
    const int ns = int(A.sptr[r + 1] - sp);
    const uint32_t bytes = uint32_t(A.I.n2p) * 8u;
    const uint32_t tab_bytes = uint32_t(D.tab_max) * 16u;
    const uint32_t desc_bytes = uint32_t(D.desc_max) * 4u;
    const uint32_t rec_bytes = uint32_t((size_t(nrows_cta) * rstride * 8 + 15) & ~size_t(15));
    const uint32_t sab_bytes = uint32_t(size_t(nrows_cta) * D.per_row * 8);
    int mine = 0;
    for (int q = lane; q < ns; q += 32) mine += ((A.smeta[sp + q] >> 17) & 1u) ? 0 : 1;
    const int total = __reduce_add_sync(0xffffffffu, mine);
    if (lane == 0) {
      mbar_arrive_expect_tx(bar, uint32_t(total) * bytes + tab_bytes + desc_bytes + rec_bytes + sab_bytes);
      bulk_g2s(tab, D.tab + size_t(rl) * D.tab_max, tab_bytes, bar);
      bulk_g2s(desc, D.desc + size_t(rl) * D.desc_max, desc_bytes, bar);
      bulk_g2s(recs, D.rec + size_t(k0) * rstride, rec_bytes, bar);
      bulk_g2s(sab, D.sab + (size_t(rl) * nb + k0) * D.per_row, sab_bytes, bar);
    }
    __syncwarp();
    for (int q = lane; q < ns; q += 32) {
      const uint32_t m = A.smeta[sp + q];
      if ((m >> 17) & 1u) continue;
      const size_t pq = ((m >> 8) & 0xFFu) + size_t(m & 0xFFu) * n;
      bulk_g2s(slices + size_t(q) * A.I.n2p, A.I.Vt + pq * A.I.n2p, bytes, bar);
    }
  }
  // row scalars of the CTA's rows (slot offset, diagonal element): fetched by one thread per row while the
  // bulk copies are in flight, so that no row starts with a dependent global load
  __shared__ int64_t s_out0[DROWS];
  __shared__ double s_diag[DROWS];
  if (w >= 1 && int(threadIdx.x) - 32 < nrows_cta) {
    const int j = int(threadIdx.x) - 32;
    const int64_t i = int64_t(r) * nb + k0 + j;
    const bool in = i >= A.row_begin && i < A.row_begin + A.nrows;
    s_out0[j] = in ? A.rowptr[i - A.row_begin] : -1;
    s_diag[j] = in ? A.diag[i - A.row_begin] : 0.;
  }
  const int* __restrict__ cnt = D.tab_cnt + size_t(rl) * 4;
  const int self_off = cnt[2], row_len = cnt[3];
  const unsigned lt = (1u << lane) - 1u;
  const uint32_t s_base = smem_u32(smem_raw), s_tab = smem_u32(tab);
  __syncthreads();
  mbar_wait(bar, 0);
  for (int j = w; j < nrows_cta; j += DW) {
    const int64_t out0 = s_out0[j];
    if (out0 < 0) continue;  // outside the row block (warp-uniform)
    const int64_t row = int64_t(r) * nb + k0 + j - A.row_begin;
    const uint32_t k = uint32_t(k0 + j);
    const int64_t b4s = int64_t(k) * len4;  // uniform lists: b4_ptr[k] == k * len4
    const double dgv = s_diag[j];
    const uint32_t s_sa = smem_u32(sab + size_t(j) * D.per_row);
    const double* sb = sab + size_t(j) * D.per_row + A.smem_a;
    const uint2* rec = recs + size_t(j) * rstride;
    int ndrop = 0;
    // the B4(k) list lives in global memory (L2): its first loads are in flight during the first pass
    constexpr int U4 = 4;
    uint32_t bpk_n[U4];
    double bv_n[U4];
    // ---- pass 2's windows are aligned like pass 1's: absolute position a = out0 + p, 32 per step
    const int64_t sa0 = out0 + self_off, sa1 = sa0 + len4;     // the self segment, absolute
    const int64_t sw0 = sa0 & ~int64_t(31);                    // its first window
#pragma unroll
    for (int u = 0; u < U4; ++u) {
      const int64_t t = sw0 + u * 32 + lane - sa0;
      const int64_t e4 = b4s + min(max(t, int64_t(0)), int64_t(len4 - 1));
      bpk_n[u] = __ldg(A.b4 + e4);
      bv_n[u] = ldg(A.b4_val + e4);
    }
    // ---- pass 1: every position outside the self segment, in order; each warp store is ONE contiguous,
    // 128-byte (colind) / 256-byte (nzval) aligned range -- what the store path wants (a store split in
    // two ranges costs more than twice as much, scripts/micro/store_pattern.cu). Branch-free: singles,
    // unit doubles and the sa[so] cases differ only in where the value is read from.
    {
      const int head = int(out0 & 31);
      int32_t* ci_row = A.colind + out0;
      double* nz_row = A.nzval + out0;
      const uint32_t s_rec = smem_u32(rec);
#pragma unroll 1
      for (int half = 0; half < 2; ++half) {
        const int lo = half == 0 ? 0 : self_off + len4, hi = half == 0 ? self_off : row_len;
#pragma unroll 4
        for (int pw = ((lo + head) & ~31) - head; pw < hi; pw += 32) {
          const int pp = pw + lane;
          if (pp >= lo && pp < hi) {
            const uint32_t d = desc[pp];
            const uint32_t e16 = d & 0xFFF0u;
            uint4 ent;
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(ent.x), "=r"(ent.y), "=r"(ent.z), "=r"(ent.w) : "r"(s_tab + e16));
            uint2 rc;
            asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(rc.x), "=r"(rc.y) : "r"(s_rec + ((d >> 13) & 0x7FFF8u)));
            // value = *(home + record offset), sign = record ^ entry; the same-beta element of a single (and a
            // dead single) reads the row's sa[so] instead
            uint32_t addr = s_base + ent.z + ((rc.y >> (ent.w & 31u)) & 0x7FF8u);
            uint32_t sg = (rc.y ^ ent.w) & 0x80000000u;
            if (rc.y & (1u << 30)) {
              addr = s_sa + ((ent.w >> 5) & 0x7FFF8u);
              sg = 0u;
            }
            const double vr = lds_f64(addr);
            const double v = __hiloint2double(__double2hiint(vr) ^ int(sg), __double2loint(vr));
            ci_row[pp] = int32_t(ent.x + rc.x);
            nz_row[pp] = v;
            if (EVAL && !(fabs(v) > A.thr)) ++ndrop;
          }
        }
      }
    }
    // ---- pass 2: the self segment, same alpha string x B4(k); U4 steps' loads are issued together
    {
      const uint32_t base = r * nb;
      int t2run = 0;
      for (int64_t aw = sw0; aw < sa1; aw += U4 * 32) {
        uint32_t bpk_u[U4];
        double bv_u[U4];
#pragma unroll
        for (int u = 0; u < U4; ++u) { bpk_u[u] = bpk_n[u]; bv_u[u] = bv_n[u]; }
        if (aw + U4 * 32 < sa1) {
#pragma unroll
          for (int u = 0; u < U4; ++u) {
            const int64_t t = aw + (U4 + u) * 32 + lane - sa0;
            const int64_t e4 = b4s + min(max(t, int64_t(0)), int64_t(len4 - 1));
            bpk_n[u] = __ldg(A.b4 + e4);
            bv_n[u] = ldg(A.b4_val + e4);
          }
        }
#pragma unroll
        for (int u = 0; u < U4; ++u) {
          const int64_t a = aw + u * 32 + lane;
          if (aw + u * 32 < sa1) {  // warp-uniform
            const bool act = a >= sa0 && a < sa1;
            const uint32_t bpk = bpk_u[u];
            const int db = int(bpk & 3u);
            const unsigned g01 = __ballot_sync(0xffffffffu, act && db <= 1);
            double v;
            if (db == 2) v = bv_u[u];
            else if (db == 1) v = act ? sb[t2run + __popc(g01 & lt)] : 0.;
            else v = dgv;
            t2run += __popc(g01);
            if (act) {
              A.colind[a] = int32_t(base + (bpk >> 2));
              A.nzval[a] = v;
              if (EVAL && !(fabs(v) > A.thr)) ++ndrop;
            }
          }
        }
      }
    }
    if (EVAL) ndrop = __reduce_add_sync(0xffffffffu, ndrop);
    if (lane == 0) A.row_cnt[row] = row_len - ndrop;
  }
}

// DENSE builds that did drop elements: pack every row, keeping |h| > thr in order
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_compact_rows_filter(int64_t nrows, const int64_t* __restrict__ slot_ptr, const int64_t* __restrict__ rowptr,
                      const int32_t* __restrict__ ci_in, const double* __restrict__ nz_in, double thr,
                      int32_t* __restrict__ ci_out, double* __restrict__ nz_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t src = slot_ptr[row], len = slot_ptr[row + 1] - src;
  int64_t dst = rowptr[row];
  for (int64_t t0 = 0; t0 < len; t0 += 32) {
    const int64_t t = t0 + lane;
    int32_t c = 0;
    double v = 0.;
    if (t < len) { c = ci_in[src + t]; v = nz_in[src + t]; }
    const bool keep = t < len && fabs(v) > thr;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t pos = dst + __popc(m & ((1u << lane) - 1u));
      ci_out[pos] = c;
      nz_out[pos] = v;
    }
    dst += __popc(m);
  }
}
// shape of a rectangular list: min / max of the beta adjacency lengths and the longest compacted run adjacency
__global__ void k_shape_minmax(const int64_t* __restrict__ b2_ptr, const int64_t* __restrict__ b4_ptr, int64_t nb,
                               const int64_t* __restrict__ cptr, const int32_t* __restrict__ run_cnt, int64_t nruns,
                               int* __restrict__ out /* 6, pre-set */) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i < nb) {
    const int l2 = int(b2_ptr[i + 1] - b2_ptr[i]), l4 = int(b4_ptr[i + 1] - b4_ptr[i]);
    atomicMin(out + 0, l2); atomicMax(out + 1, l2);
    atomicMin(out + 2, l4); atomicMax(out + 3, l4);
  }
  if (i < nruns) {
    atomicMax(out + 4, int(cptr[i + 1] - cptr[i]));
    // structural row length of the run's rows if the beta lists are uniform (lengths of string 0)
    const int64_t l2 = b2_ptr[1] - b2_ptr[0], l4 = b4_ptr[1] - b4_ptr[0];
    const int64_t len = run_cnt[4 * i] * l4 + run_cnt[4 * i + 1] * l2 + run_cnt[4 * i + 2];
    atomicMax(out + 5, int(min(len, int64_t(INT32_MAX))));
  }
}
__global__ void k_count_small(const double* __restrict__ v, int64_t n, double thr, unsigned int* __restrict__ cnt) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  const bool small = i < n && !(fabs(v[i]) > thr);
  const unsigned m = __ballot_sync(0xffffffffu, small);
  if ((threadIdx.x & 31) == 0 && m) atomicAdd(cnt, unsigned(__popc(m)));
}

// rows filled at their structural positions with the dropped elements marked (colind = -1): pack what is left
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_compact_rows_holes(int64_t nrows, const int64_t* __restrict__ slot_ptr, const int64_t* __restrict__ rowptr,
                     const int32_t* __restrict__ ci_in, const double* __restrict__ nz_in,
                     int32_t* __restrict__ ci_out, double* __restrict__ nz_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t src = slot_ptr[row], len = slot_ptr[row + 1] - src;
  int64_t dst = rowptr[row];
  if (rowptr[row + 1] - dst == len) {  // nothing dropped in this row
    for (int64_t t = lane; t < len; t += 32) {
      ci_out[dst + t] = ci_in[src + t];
      nz_out[dst + t] = nz_in[src + t];
    }
    return;
  }
  for (int64_t t0 = 0; t0 < len; t0 += 32) {
    const int64_t t = t0 + lane;
    const int32_t c = t < len ? ci_in[src + t] : -1;
    const bool keep = c >= 0;
    const unsigned m = __ballot_sync(0xffffffffu, keep);
    if (keep) {
      const int64_t pos = dst + __popc(m & ((1u << lane) - 1u));
      ci_out[pos] = c;
      nz_out[pos] = nz_in[src + t];
    }
    dst += __popc(m);
  }
}
// move the surviving prefix of every structural row slot to its final position
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_compact_rows(int64_t nrows, const int64_t* __restrict__ slot_ptr,
               const int64_t* __restrict__ rowptr, const int32_t* __restrict__ ci_in,
               const double* __restrict__ nz_in, int32_t* __restrict__ ci_out,
               double* __restrict__ nz_out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (row >= nrows) return;
  const int64_t src = slot_ptr[row], dst = rowptr[row], len = rowptr[row + 1] - dst;
  for (int64_t t = lane; t < len; t += 32) {
    ci_out[dst + t] = ci_in[src + t];
    nz_out[dst + t] = nz_in[src + t];
  }
}

__global__ void k_unpack_dets(const uint64_t* __restrict__ words, int wpd, int64_t n,
                              uint64_t* __restrict__ alpha, uint64_t* __restrict__ beta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (wpd == 1) {
    const uint64_t w = words[i];
    alpha[i] = w & 0xFFFFFFFFull;
    beta[i] = w >> 32;
  } else {
    alpha[i] = words[2 * i];
    beta[i] = words[2 * i + 1];
  }
}
__global__ void k_pack_dets(const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta,
                            int wpd, int64_t n, uint64_t* __restrict__ words) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (wpd == 1) words[i] = (alpha[i] & 0xFFFFFFFFull) | (beta[i] << 32);
  else { words[2 * i] = alpha[i]; words[2 * i + 1] = beta[i]; }
}

// generate_combs order (sd_operations.hpp:305-323): std::prev_permutation of a 0/1 vector
// whose first nset entries are set == combinations in DESCENDING order of the bit-reversed
// string. Thread t unranks combination t directly: walking positions 0..nbits-1, position p
// is set iff t < C(nbits-p-1, remaining-1) (the block of combinations that keep bit p).
__device__ __forceinline__ uint64_t unrank_comb(int nbits, int nset, int64_t t,
                                                const int64_t* __restrict__ binom /*65x65*/) {
  uint64_t s = 0;
  int rem = nset;
  for (int p = 0; p < nbits && rem > 0; ++p) {
    const int64_t with_p = binom[(nbits - p - 1) * 65 + (rem - 1)];
    if (t < with_p) { s |= uint64_t(1) << p; --rem; }
    else t -= with_p;
  }
  return s;
}
__global__ void k_generate_fci(int norb, int na, int nb, int64_t nalpha_str, int64_t nbeta_str,
                               const int64_t* __restrict__ binom, uint64_t* __restrict__ alpha,
                               uint64_t* __restrict__ beta) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= nalpha_str * nbeta_str) return;
  alpha[i] = unrank_comb(norb, na, i / nbeta_str, binom);
  beta[i] = unrank_comb(norb, nb, i % nbeta_str, binom);
}

// ------------------------------------------------------------------ patched (incremental) build
// adjacency between two string tables (bra runs x ket runs), entries as k_string_adjacency
template <bool FILL>
__global__ void __launch_bounds__(256)
k_string_adjacency2(const uint64_t* __restrict__ sa, int32_t na, const uint64_t* __restrict__ sb, int32_t nb,
                    int maxd, int skip_zero, int32_t* __restrict__ cnt, const int64_t* __restrict__ adj_ptr,
                    uint32_t* __restrict__ adj) {
  const int lane = threadIdx.x & 31;
  const int64_t r = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (r >= na) return;
  const uint64_t a = sa[r];
  int64_t out = FILL ? adj_ptr[r] : 0;
  int32_t c = 0;
  if (!(skip_zero && a == 0)) {
    for (int32_t r0 = 0; r0 < nb; r0 += 32) {
      const int32_t r2 = r0 + lane;
      bool ok = false;
      int d = 0;
      if (r2 < nb) {
        const uint64_t a2 = sb[r2];
        d = __popcll(a ^ a2);
        ok = !(skip_zero && a2 == 0) && d <= maxd;
      }
      const unsigned m = __ballot_sync(0xffffffffu, ok);
      if (FILL && ok) adj[out + __popc(m & ((1u << lane) - 1u))] = (uint32_t(r2) << 2) | uint32_t(d >> 1);
      out += __popc(m);
      c += __popc(m);
    }
  }
  if (!FILL && lane == 0) cnt[r] = c;
}
// beta group of the ket list that holds each bra determinant's beta string (-1: none)
__global__ void k_lookup_group(const uint64_t* __restrict__ bra_beta, int64_t n,
                               const uint64_t* __restrict__ key_sorted, const int64_t* __restrict__ grp_start,
                               int32_t ngroups, int32_t* __restrict__ grp) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t b = bra_beta[i];
  int32_t lo = 0, hi = ngroups;
  while (lo < hi) {
    const int32_t mid = (lo + hi) >> 1;
    if (key_sorted[grp_start[mid]] < b) lo = mid + 1; else hi = mid;
  }
  grp[i] = (lo < ngroups && key_sorted[grp_start[lo]] == b) ? lo : -1;
}
// position of every new determinant in the old list (both spin_comparator-sorted: alpha-major,
// then beta -- raw_bitset.hpp:119-141), -1 if it is new; the merge scan of
// build_patched_operator (incremental_h_build.hpp:237-262) as one binary search per determinant
__global__ void k_classify(const uint64_t* __restrict__ na, const uint64_t* __restrict__ nb, int64_t n_new,
                           const uint64_t* __restrict__ oa, const uint64_t* __restrict__ ob, int64_t n_old,
                           int32_t* __restrict__ new_to_old, int32_t* __restrict__ old_to_new,
                           int32_t* __restrict__ kept_flag, int32_t* __restrict__ added_flag) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  const uint64_t a = na[i], b = nb[i];
  int64_t lo = 0, hi = n_old;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    const uint64_t ma = oa[mid], mb = ob[mid];
    if (ma < a || (ma == a && mb < b)) lo = mid + 1; else hi = mid;
  }
  const bool found = lo < n_old && oa[lo] == a && ob[lo] == b;
  new_to_old[i] = found ? int32_t(lo) : -1;
  if (found) old_to_new[lo] = int32_t(i);
  kept_flag[i] = found ? 1 : 0;
  added_flag[i] = found ? 0 : 1;
}
// index lists and sub-lists of the kept / added determinants
__global__ void k_split_lists(const uint64_t* __restrict__ na, const uint64_t* __restrict__ nb, int64_t n_new,
                              const int32_t* __restrict__ kept_flag, const int32_t* __restrict__ kept_excl,
                              const int32_t* __restrict__ added_excl, int32_t* __restrict__ kept_new,
                              int32_t* __restrict__ added_new, uint64_t* __restrict__ ka, uint64_t* __restrict__ kb,
                              uint64_t* __restrict__ aa, uint64_t* __restrict__ ab) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  if (kept_flag[i]) {
    const int32_t k = kept_excl[i];
    kept_new[k] = int32_t(i); ka[k] = na[i]; kb[k] = nb[i];
  } else {
    const int32_t k = added_excl[i];
    added_new[k] = int32_t(i); aa[k] = na[i]; ab[k] = nb[i];
  }
}
// kept x kept block: the old rows of the kept determinants, dropped columns removed, columns
// renumbered (old_to_new is monotone on the kept determinants, so rows stay ascending)
template <bool FILL>
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_kept_rows(int64_t n_kept, const int32_t* __restrict__ kept_new, const int32_t* __restrict__ new_to_old,
            const int32_t* __restrict__ old_to_new, const int64_t* __restrict__ orp,
            const int32_t* __restrict__ oci, const double* __restrict__ onz, int32_t* __restrict__ cnt,
            const int64_t* __restrict__ kptr, int32_t* __restrict__ kci, double* __restrict__ knz) {
  const int lane = threadIdx.x & 31;
  const int64_t kr = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (kr >= n_kept) return;
  const int32_t io = new_to_old[kept_new[kr]];
  const int64_t e0 = orp[io], e1 = orp[io + 1];
  int64_t out = FILL ? kptr[kr] : 0;
  int32_t c = 0;
  for (int64_t e = e0; e < e1; e += 32) {
    const int64_t p = e + lane;
    int32_t cn = -1;
    if (p < e1) cn = old_to_new[oci[p]];
    const unsigned m = __ballot_sync(0xffffffffu, cn >= 0);
    if (FILL && cn >= 0) {
      const int64_t pos = out + __popc(m & ((1u << lane) - 1u));
      kci[pos] = cn;
      knz[pos] = onz[p];
    }
    out += __popc(m);
    c += __popc(m);
  }
  if (!FILL && lane == 0) cnt[kr] = c;
}
// row lengths of the patched matrix: added rows = their freshly built row, kept rows = kept x kept
// part + kept x added part
__global__ void k_patch_count(int64_t n_new, const int32_t* __restrict__ kept_flag,
                              const int32_t* __restrict__ kept_excl, const int32_t* __restrict__ added_excl,
                              const int64_t* __restrict__ kptr, const int64_t* __restrict__ dkptr,
                              const int64_t* __restrict__ daptr, int32_t* __restrict__ cnt) {
  const int64_t i = int64_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (i >= n_new) return;
  if (kept_flag[i]) {
    const int32_t k = kept_excl[i];
    cnt[i] = int32_t((kptr[k + 1] - kptr[k]) + (dkptr ? dkptr[k + 1] - dkptr[k] : 0));
  } else {
    const int32_t k = added_excl[i];
    cnt[i] = int32_t(daptr[k + 1] - daptr[k]);
  }
}
__device__ __forceinline__ int64_t lower_bound_i32(const int32_t* __restrict__ v, int64_t n, int32_t x) {
  int64_t lo = 0, hi = n;
  while (lo < hi) {
    const int64_t mid = (lo + hi) >> 1;
    if (v[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
// merge by column: the two parts of a kept row have disjoint, ascending column sets, so every
// element's final position is its own rank plus its lower bound in the other part
__global__ void __launch_bounds__(ROW_WARPS * 32)
k_patch_fill(int64_t n_new, const int32_t* __restrict__ kept_flag, const int32_t* __restrict__ kept_excl,
             const int32_t* __restrict__ added_excl, const int64_t* __restrict__ kptr,
             const int32_t* __restrict__ kci, const double* __restrict__ knz,
             const int64_t* __restrict__ dkptr, const int32_t* __restrict__ dkci,
             const double* __restrict__ dknz, const int64_t* __restrict__ daptr,
             const int32_t* __restrict__ daci, const double* __restrict__ danz,
             const int64_t* __restrict__ rowptr, int32_t* __restrict__ ci, double* __restrict__ nz) {
  const int lane = threadIdx.x & 31;
  const int64_t i = (int64_t(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
  if (i >= n_new) return;
  const int64_t dst = rowptr[i];
  if (!kept_flag[i]) {
    const int32_t k = added_excl[i];
    const int64_t src = daptr[k], len = daptr[k + 1] - src;
    for (int64_t t = lane; t < len; t += 32) { ci[dst + t] = daci[src + t]; nz[dst + t] = danz[src + t]; }
    return;
  }
  const int32_t k = kept_excl[i];
  const int64_t ks = kptr[k], kl = kptr[k + 1] - ks;
  const int64_t ds = dkptr ? dkptr[k] : 0, dl = dkptr ? dkptr[k + 1] - ds : 0;
  for (int64_t t = lane; t < kl; t += 32) {
    const int32_t c = kci[ks + t];
    const int64_t pos = dst + t + (dl ? lower_bound_i32(dkci + ds, dl, c) : 0);
    ci[pos] = c;
    nz[pos] = knz[ks + t];
  }
  for (int64_t t = lane; t < dl; t += 32) {
    const int32_t c = dkci[ds + t];
    const int64_t pos = dst + t + lower_bound_i32(kci + ks, kl, c);
    ci[pos] = c;
    nz[pos] = dknz[ds + t];
  }
}

}  // namespace

// ------------------------------------------------------------------------------------
void dets_from_words(b2ci_ctx* ctx, const uint64_t* words_host, int wpd, int64_t n,
                     b2ci_dets* d) {
  if (wpd != 1 && wpd != 2) throw Error("words_per_det must be 1 (wfn_t<64>) or 2 (wfn_t<128>)");
  d->n = n;
  DevBuf<uint64_t> a(n), b(n), w(size_t(n) * wpd);
  if (n) {
    B2_CUDA(cudaMemcpyAsync(w, words_host, size_t(n) * wpd * 8, cudaMemcpyHostToDevice, ctx->stream));
    k_unpack_dets<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(w, wpd, n, a, b);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaStreamSynchronize(ctx->stream));
  }
  d->alpha = a.take();
  d->beta = b.take();
}

void dets_to_words(b2ci_ctx* ctx, const b2ci_dets* d, uint64_t* words_host, int wpd) {
  if (wpd != 1 && wpd != 2) throw Error("words_per_det must be 1 or 2");
  if (!d->n) return;
  DevBuf<uint64_t> w(size_t(d->n) * wpd);
  k_pack_dets<<<unsigned((d->n + 255) / 256), 256, 0, ctx->stream>>>(d->alpha, d->beta, wpd, d->n, w);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  B2_CUDA(cudaMemcpyAsync(words_host, w, size_t(d->n) * wpd * 8, cudaMemcpyDeviceToHost, ctx->stream));
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
}

void dets_generate_fci(b2ci_ctx* ctx, int norb, int na, int nb, b2ci_dets* d) {
  if (norb < 1 || norb > 64 || na < 0 || nb < 0 || na > norb || nb > norb)
    throw Error("generate_hilbert_space: invalid (norb, nalpha, nbeta)");
  std::vector<int64_t> binom(65 * 65, 0);
  for (int n = 0; n <= 64; ++n) {
    binom[n * 65 + 0] = 1;
    for (int k = 1; k <= n; ++k) {
      const __int128 v = (__int128)binom[(n - 1) * 65 + (k - 1)] + (k <= n - 1 ? binom[(n - 1) * 65 + k] : 0);
      binom[n * 65 + k] = v > (__int128)INT64_MAX ? INT64_MAX : (int64_t)v;
    }
  }
  const int64_t nas = binom[norb * 65 + na], nbs = binom[norb * 65 + nb];
  if (nas == INT64_MAX || nbs == INT64_MAX || nas > INT64_MAX / (nbs ? nbs : 1))
    throw Error("generate_hilbert_space: dimension overflows int64");
  const int64_t n = nas * nbs;
  DevBuf<int64_t> dbinom(binom.size());
  DevBuf<uint64_t> a(n), b(n);
  B2_CUDA(cudaMemcpyAsync(dbinom, binom.data(), binom.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  k_generate_fci<<<unsigned((n + 255) / 256), 256, 0, ctx->stream>>>(norb, na, nb, nas, nbs, dbinom, a, b);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  B2_CUDA(cudaStreamSynchronize(ctx->stream));
  d->n = n;
  d->alpha = a.take();
  d->beta = b.take();
}

namespace {
// Row scan of a general (or rectangular-block) list, shared by hbuild_csr and the patched build.
// Count pass = the scan alone (structural connections, no matrix element); the fill pass
// evaluates every element ONCE, writes the survivors compacted inside the row's structural
// slot and records how many there were; rows are packed afterwards only if something was
// dropped (threshold_parallel, csr_matrix.hpp:317-370).
// The scan (XOR + popcount over the beta strings of the adjacent runs) is what a general build
// costs, so it is done once: a sampled estimate pass (every 64th row) sizes a chunk store, the
// count pass keeps the connections it finds there, and the fill pass evaluates them from the
// store. If the store turns out too small the fill pass scans again (same result).
struct RowScanOut {
  DevBuf<int64_t> rowptr;
  DevBuf<int32_t> colind;
  DevBuf<double> nzval;
  int64_t nnz = 0;
  size_t ci_cap = 0, nz_cap = 0;  // non-zero: blocks came from the context's slot cache
};
// B2CI_HBUILD_TRACE: device time of the scan's kernels, one line per phase
struct PhaseTrace {
  cudaStream_t st;
  bool on;
  cudaEvent_t a, b;
  explicit PhaseTrace(cudaStream_t s) : st(s), on(getenv("B2CI_HBUILD_TRACE") != nullptr) {
    if (on) { cudaEventCreate(&a); cudaEventCreate(&b); }
  }
  ~PhaseTrace() { if (on) { cudaEventDestroy(a); cudaEventDestroy(b); } }
  void begin() { if (on) cudaEventRecord(a, st); }
  void end(const char* what, double units) {
    if (!on) return;
    cudaEventRecord(b, st);
    cudaEventSynchronize(b);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    fprintf(stderr, "[hbuild scan] %-22s %9.3f ms  (%.4g)\n", what, ms, units);
  }
};
template <bool BLK>
void run_row_scan(b2ci_ctx* ctx, RowArgs A, int64_t nrows, int64_t nket, double thr, bool use_slot_cache, RowScanOut& out) {
  cudaStream_t st = ctx->stream;
  PhaseTrace PT(st);
  const unsigned grid = unsigned((nrows + ROW_WARPS - 1) / ROW_WARPS);
  int64_t nslots = 0;
  DevBuf<int64_t> slot_ptr(nrows + 1);
  DevBuf<int32_t> row_cnt(nrows);
  // store of the warp-per-row scan (rows of short units)
  DevBuf<int32_t> hit_cols, hit_next, hit_head;
  DevBuf<unsigned int> cursors(2);  // [0] warp-per-row chunks, [1] tile chunks
  B2_CUDA(cudaMemsetAsync(cursors, 0, 2 * sizeof(unsigned int), st));  // (read back below whether or not a store is used)
  unsigned int hit_capacity = 0, hit_used = 0;
  // tiled scan (units of >= tile_min rows of one alpha run)
  DevBuf<int32_t> unit_row0, unit_len, unit_head, unit_nent, unit_nent_c, row_cnt_c, scan_rows, tile_next;
  DevBuf<uint2> tile_ent;
  unsigned int tile_capacity = 0, tile_used = 0;
  int32_t ntile = 0;
  int64_t nscan = nrows;
  // norb <= 32: beta strings as 32-bit words
  DevBuf<uint32_t> beta32, alpha32;
  const bool w32 = ctx->norb <= 32 && !getenv("B2CI_HBUILD_WIDE_STRINGS");  // (test hook: 64-bit strings)
  auto launch_scan_count = [&](const RowArgs& R, int64_t nr) {
    const unsigned g = unsigned((nr + ROW_WARPS - 1) / ROW_WARPS);
    if (w32) k_rows<false, false, BLK, uint32_t><<<g, ROW_WARPS * 32, 0, st>>>(R);
    else k_rows<false, false, BLK, uint64_t><<<g, ROW_WARPS * 32, 0, st>>>(R);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  };
  TileArgs TA;
  {
    ScopedTimer t(ctx, "h_build.count", true);
    if (w32) {
      beta32.alloc(nket);
      k_narrow_u32<<<unsigned((nket + 255) / 256), 256, 0, st>>>(A.beta, nket, beta32);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      A.beta_s = beta32;
      alpha32.alloc(nket);
      k_narrow_u32<<<unsigned((nket + 255) / 256), 256, 0, st>>>(A.alpha, nket, alpha32);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      A.alpha_s = alpha32;
    }
    // (test hooks: B2CI_HBUILD_HITLIST_MIN = smallest list that keeps the connections of its count pass,
    // B2CI_HBUILD_HITLIST_CAP / B2CI_HBUILD_TILE_CAP = capacities in chunks, to exercise the overflow fallback,
    // B2CI_HBUILD_NO_TILE = warp-per-row scan for every row, B2CI_HBUILD_TILE_MIN = smallest tiled unit)
    const char* env_min = getenv("B2CI_HBUILD_HITLIST_MIN");
    const int64_t min_rows = env_min ? atoll(env_min) : 4096;
    const bool want_hits = nrows >= min_rows && !getenv("B2CI_HBUILD_NO_HITLIST");
    if (want_hits) {
      // estimate: structural row lengths of every 64th (256th) row
      const int64_t stride = nrows >= (int64_t(1) << 21) ? 256 : 64, ns = (nrows + stride - 1) / stride;
      DevBuf<int32_t> scnt(ns);
      DevBuf<int64_t> sptr(ns + 1);
      RowArgs S = A;
      S.row_stride = stride;
      S.nrows = ns;
      S.row_cnt = scnt;
      PT.begin();
      launch_scan_count(S, ns);
      PT.end("estimate (rows)", double(ns));
      exclusive_scan_i32_to_i64(ctx, scnt, sptr, ns);
      int64_t* pin = pinned_words(ctx);
      B2_CUDA(cudaMemcpyAsync(pin, sptr.p + ns, 8, cudaMemcpyDeviceToHost, st));
      // units of the tiled scan: <= 32 consecutive rows of one alpha run, aligned to the run's start
      const bool want_tile = !getenv("B2CI_HBUILD_NO_TILE");
      // smallest tiled unit: measured on the 1e6-determinant N2 / Cr2 ASCI lists with the per-step emit of round 2
      // (tiled + row scan): 3 -> 34.7 / 57.9 ms, 5 -> 35.0 / 58.0, 8 -> 35.4 / 59.3, 12 -> 36.8 / 59.8
      int32_t tile_min = 4;
      if (const char* env = getenv("B2CI_HBUILD_TILE_MIN")) tile_min = std::max(1, atoi(env));
      DevBuf<int32_t> uflag, uexcl, tflag, texcl, sflag, sexcl;
      if (want_tile) {
        uflag.alloc(nrows); uexcl.alloc(nrows + 1); tflag.alloc(nrows); texcl.alloc(nrows + 1);
        sflag.alloc(nrows); sexcl.alloc(nrows + 1);
        const unsigned gr = unsigned((nrows + 255) / 256);
        k_unit_flags<<<gr, 256, 0, st>>>(nrows, A.row_begin, BLK ? A.bra_run : A.run_of,
                                         BLK ? A.bra_run_start : A.run_start, uflag);
        k_unit_classify<<<gr, 256, 0, st>>>(nrows, uflag, uexcl, tile_min, tflag, sflag);
        ctx->launches += 2;
        B2_CHECK_LAUNCH();
        exclusive_scan_i32(ctx, tflag, texcl, nrows);
        exclusive_scan_i32(ctx, sflag, sexcl, nrows);
        B2_CUDA(cudaMemcpyAsync(pin + 1, texcl.p + nrows, 4, cudaMemcpyDeviceToHost, st));
        B2_CUDA(cudaMemcpyAsync(pin + 2, sexcl.p + nrows, 4, cudaMemcpyDeviceToHost, st));
      }
      B2_CUDA(cudaStreamSynchronize(st));
      const int64_t sample = pin[0];
      if (want_tile) {
        ntile = *reinterpret_cast<const int32_t*>(pin + 1);
        nscan = *reinterpret_cast<const int32_t*>(pin + 2);
      }
      const double est = double(sample) * double(nrows) / double(ns) * 1.25 + 4096.0;
      const double frac_scan = double(nscan) / double(nrows);
      // warp-per-row store: 132 B per chunk of 32 connections + one partial chunk per row;
      // tile store: 8 B per entry, at most one entry per connection, + one partial chunk per unit
      const double chunks = est * std::min(1.0, 2.0 * frac_scan + 0.01) / 32.0 + double(nscan) + 1024.0;
      const double tchunks = ntile ? est / double(TILE_CH) + double(ntile) + 1024.0 : 0.0;
      size_t free_b = 0, total_b = 0;
      B2_CUDA(cudaMemGetInfo(&free_b, &total_b));
      // the stores must leave room for the matrix itself (12 B per entry)
      if (chunks < 2.0e9 && tchunks < 2.0e9 &&
          chunks * 132.0 + tchunks * (TILE_CH * 8.0 + 4.0) + est * 12.0 < 0.8 * double(free_b)) {
        B2_CUDA(cudaMemsetAsync(cursors, 0, 2 * sizeof(unsigned int), st));
        if (nscan) {
          hit_capacity = unsigned(chunks);
          if (const char* env_cap = getenv("B2CI_HBUILD_HITLIST_CAP")) hit_capacity = unsigned(std::max<long long>(1, atoll(env_cap)));
          hit_cols.alloc(size_t(hit_capacity) * 32);
          hit_next.alloc(hit_capacity);
          hit_head.alloc(nrows);
          A.hit_cols = hit_cols;
          A.hit_next = hit_next;
          A.hit_head = hit_head;
          A.hit_cursor = cursors.p;
          A.hit_capacity = hit_capacity;
        }
        if (ntile) {
          tile_capacity = unsigned(tchunks);
          if (const char* env_cap = getenv("B2CI_HBUILD_TILE_CAP")) tile_capacity = unsigned(std::max<long long>(1, atoll(env_cap)));
          tile_ent.alloc(size_t(tile_capacity) * TILE_CH);
          tile_next.alloc(tile_capacity);
          unit_row0.alloc(ntile); unit_len.alloc(ntile); unit_head.alloc(ntile); unit_nent.alloc(ntile);
          unit_nent_c.alloc(ntile); row_cnt_c.alloc(nrows);
        }
        if (want_tile) {
          scan_rows.alloc(nscan > 0 ? nscan : 1);
          if (!ntile) { unit_row0.alloc(1); unit_len.alloc(1); unit_head.alloc(1); }
          k_unit_lists<<<unsigned((nrows + 255) / 256), 256, 0, st>>>(nrows, uflag, tflag, texcl, sflag, sexcl, unit_row0,
                                                                      unit_len, unit_head, scan_rows);
          ctx->launches++;
          B2_CHECK_LAUNCH();
        }
      } else {
        ntile = 0;
        nscan = nrows;
      }
    }
    A.row_cnt = row_cnt;
    if (ntile) {
      TA.A = A;
      TA.beta_s = w32 ? static_cast<const void*>(beta32.p) : static_cast<const void*>(A.beta);
      TA.unit_row0 = unit_row0; TA.unit_len = unit_len; TA.nunits = ntile;
      TA.tile_ent = tile_ent; TA.tile_next = tile_next; TA.unit_head = unit_head; TA.unit_nent = unit_nent;
      TA.unit_nent_c = unit_nent_c; TA.row_cnt_c = row_cnt_c;
      TA.cursor = cursors.p + 1;
      TA.capacity = tile_capacity;
      const unsigned gt = unsigned((ntile + TILE_WARPS - 1) / TILE_WARPS);
      PT.begin();
      if (w32) k_rows_tile<uint32_t, BLK><<<gt, TILE_WARPS * 32, 0, st>>>(TA);
      else k_rows_tile<uint64_t, BLK><<<gt, TILE_WARPS * 32, 0, st>>>(TA);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      PT.end("tiled scan (units)", double(ntile));
    }
    if (nscan) {
      RowArgs R = A;
      if (ntile) { R.row_list = scan_rows; R.nrows = nscan; }
      PT.begin();
      launch_scan_count(R, R.nrows);
      PT.end("row scan (rows)", double(R.nrows));
    }
    exclusive_scan_i32_to_i64(ctx, row_cnt, slot_ptr, nrows);
    int64_t* pin = pinned_words(ctx);
    B2_CUDA(cudaMemcpyAsync(pin, slot_ptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(pin + 1, cursors, 2 * sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    nslots = pin[0];
    hit_used = reinterpret_cast<const unsigned int*>(pin + 1)[0];
    tile_used = reinterpret_cast<const unsigned int*>(pin + 1)[1];
  }
  const bool from_hits = (hit_capacity != 0 || tile_capacity != 0) && hit_used <= hit_capacity && tile_used <= tile_capacity &&
                         (nscan == 0 || hit_capacity != 0) && (ntile == 0 || tile_capacity != 0);
  ctx->timers["h_build.hit_lists"] = from_hits ? 1. : 0.;
  // B2CI_HBUILD_FLAT_FILL=1: evaluate the connections in arrival order (the former fill; kept for comparison)
  const bool binned = getenv("B2CI_HBUILD_FLAT_FILL") == nullptr || atoi(getenv("B2CI_HBUILD_FLAT_FILL")) == 0;
  ctx->timers["h_build.tile_units"] = double(ntile);
  ctx->timers["h_build.scan_rows"] = double(nscan);
  ctx->timers["h_build.tile_chunks"] = double(tile_used);
  if (!from_hits) { hit_cols.release(); hit_next.release(); hit_head.release(); tile_ent.release(); tile_next.release(); }
  DevBuf<int32_t> colind, kept(nrows);
  DevBuf<double> nzval;
  size_t ci_cap = 0, nz_cap = 0;
  if (use_slot_cache) {
    colind.p = static_cast<int32_t*>(big_alloc(ctx, 0, size_t(nslots > 0 ? nslots : 1) * sizeof(int32_t), &ci_cap));
    colind.n = ci_cap / sizeof(int32_t);
  } else {
    colind.alloc(nslots > 0 ? nslots : 1);
  }
  A.struct_cnt = row_cnt;
  if (from_hits) {
    // connections -> the rows' slot ranges of the column array (ket indices, ascending); the stores go
    // back before the value array is allocated
    ScopedTimer t(ctx, "h_build.count", true);
    PT.begin();
    if (ntile) {
      TA.A.struct_cnt = row_cnt;
      k_tile_gather<<<unsigned((ntile + TILE_WARPS - 1) / TILE_WARPS), TILE_WARPS * 32, 0, st>>>(TA, slot_ptr, colind);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    if (nscan) {
      RowArgs R = A;
      if (ntile) { R.row_list = scan_rows; R.nrows = nscan; }
      k_hits_gather<<<unsigned((R.nrows + ROW_WARPS - 1) / ROW_WARPS), ROW_WARPS * 32, 0, st>>>(R, slot_ptr, colind);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    PT.end("gather (connections)", double(nslots));
    hit_cols.release(); hit_next.release(); hit_head.release(); tile_ent.release(); tile_next.release();
  }
  if (use_slot_cache) {
    nzval.p = static_cast<double*>(big_alloc(ctx, 1, size_t(nslots > 0 ? nslots : 1) * sizeof(double), &nz_cap));
    nzval.n = nz_cap / sizeof(double);
  } else {
    nzval.alloc(nslots > 0 ? nslots : 1);
  }
  DevBuf<double> row_diag;
  {
    ScopedTimer t(ctx, "h_build.fill", true);
    A.row_cnt = kept;
    A.rowptr = slot_ptr;
    A.colind = colind;
    A.nzval = nzval;
    A.row_list = nullptr;
    A.nrows = nrows;
    PT.begin();
    if (from_hits && binned) {
      row_diag.alloc(nrows > 0 ? nrows : 1);
      k_row_diag<BLK><<<unsigned((nrows + 127) / 128), 128, 0, st>>>(A, row_diag);
      A.diag = row_diag;
      ctx->launches++;
      const unsigned gb = unsigned((nrows + BIN_WARPS - 1) / BIN_WARPS);
      if (w32) {
        if (thr > 0.0) k_rows_hits_binned<true, BLK, uint32_t><<<gb, BIN_WARPS * 32, 0, st>>>(A);
        else k_rows_hits_binned<false, BLK, uint32_t><<<gb, BIN_WARPS * 32, 0, st>>>(A);
      } else {
        if (thr > 0.0) k_rows_hits_binned<true, BLK, uint64_t><<<gb, BIN_WARPS * 32, 0, st>>>(A);
        else k_rows_hits_binned<false, BLK, uint64_t><<<gb, BIN_WARPS * 32, 0, st>>>(A);
      }
    } else if (from_hits) {
      if (thr > 0.0) k_rows_hits_flat<true, BLK><<<grid, ROW_WARPS * 32, 0, st>>>(A);
      else k_rows_hits_flat<false, BLK><<<grid, ROW_WARPS * 32, 0, st>>>(A);
    } else {
      A.hit_cols = nullptr;
      if (thr > 0.0) k_rows<true, true, BLK><<<grid, ROW_WARPS * 32, 0, st>>>(A);
      else k_rows<true, false, BLK><<<grid, ROW_WARPS * 32, 0, st>>>(A);
    }
    ctx->launches++;
    B2_CHECK_LAUNCH();
    PT.end(from_hits ? "fill from connections" : "fill by rescan", double(nslots));
  }
  int64_t nnz = nslots;
  out.rowptr.alloc(nrows + 1);
  if (thr > 0.0) {
    ScopedTimer t(ctx, "h_build.thresh", true);
    exclusive_scan_i32_to_i64(ctx, kept, out.rowptr, nrows);
    int64_t* pin = pinned_words(ctx);
    B2_CUDA(cudaMemcpyAsync(pin, out.rowptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    nnz = pin[0];
    if (nnz != nslots) {
      DevBuf<int32_t> ci_f(nnz > 0 ? nnz : 1);
      DevBuf<double> nz_f(nnz > 0 ? nnz : 1);
      const unsigned gc = unsigned((nrows * 32 + ROW_WARPS * 32 - 1) / (ROW_WARPS * 32));
      if (from_hits && binned)
        k_compact_rows_holes<<<gc, ROW_WARPS * 32, 0, st>>>(nrows, slot_ptr, out.rowptr, colind, nzval, ci_f, nz_f);
      else
        k_compact_rows<<<gc, ROW_WARPS * 32, 0, st>>>(nrows, slot_ptr, out.rowptr, colind, nzval, ci_f, nz_f);
      ctx->launches++;
      B2_CHECK_LAUNCH();
      B2_CUDA(cudaStreamSynchronize(st));
      if (use_slot_cache) {
        big_release(ctx, 0, colind.take(), ci_cap);
        big_release(ctx, 1, nzval.take(), nz_cap);
        ci_cap = nz_cap = 0;
      }
      colind = std::move(ci_f);
      nzval = std::move(nz_f);
    }
  } else {
    out.rowptr = std::move(slot_ptr);
  }
  B2_CUDA(cudaStreamSynchronize(st));
  out.nnz = nnz;
  out.colind = std::move(colind);
  out.nzval = std::move(nzval);
  out.ci_cap = ci_cap;
  out.nz_cap = nz_cap;
}
}  // namespace

// ------------------------------------------------------------------ balanced row cuts
// degree[s] = number of determinants of the list within four spin-orbital substitutions of sample s (itself
// included): brute force, SAMPLES_PER_CTA bras against a slice of the kets per CTA, so that a ket is loaded
// once for eight tests.
namespace {
constexpr int DEG_SPC = 8;     // samples per CTA
constexpr int DEG_SLICES = 8;  // ket slices (grid.y)
__global__ void __launch_bounds__(256)
k_sample_degree(const uint64_t* __restrict__ alpha, const uint64_t* __restrict__ beta, int64_t n,
                const int64_t* __restrict__ sample_row, int nsamples, unsigned long long* __restrict__ degree) {
  __shared__ unsigned int tot[DEG_SPC];
  const int s0 = blockIdx.x * DEG_SPC;
  uint64_t sa[DEG_SPC], sb[DEG_SPC];
  unsigned int cnt[DEG_SPC];
#pragma unroll
  for (int u = 0; u < DEG_SPC; ++u) {
    const int s = min(s0 + u, nsamples - 1);
    const int64_t r = sample_row[s];
    sa[u] = alpha[r];
    sb[u] = beta[r];
    cnt[u] = 0u;
  }
  if (threadIdx.x < DEG_SPC) tot[threadIdx.x] = 0u;
  __syncthreads();
  const int64_t per = (n + gridDim.y - 1) / gridDim.y;
  const int64_t j0 = int64_t(blockIdx.y) * per, j1 = min(n, j0 + per);
  for (int64_t j = j0 + threadIdx.x; j < j1; j += blockDim.x) {
    const uint64_t a = alpha[j], b = beta[j];
#pragma unroll
    for (int u = 0; u < DEG_SPC; ++u) cnt[u] += (__popcll(sa[u] ^ a) + __popcll(sb[u] ^ b) <= 4) ? 1u : 0u;
  }
#pragma unroll
  for (int u = 0; u < DEG_SPC; ++u) {
    unsigned int v = cnt[u];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(0xffffffffu, v, d);
    if ((threadIdx.x & 31) == 0) atomicAdd(&tot[u], v);
  }
  __syncthreads();
  if (threadIdx.x < DEG_SPC && s0 + threadIdx.x < nsamples)
    atomicAdd(&degree[s0 + threadIdx.x], (unsigned long long)tot[threadIdx.x]);
}
}  // namespace

void dets_balanced_partition(b2ci_ctx* ctx, const b2ci_dets* dets, int nparts, int64_t nsamples, int64_t* offsets) {
  const int64_t n = dets->n;
  if (nparts < 1 || !offsets) throw Error("b2ci_dets_balanced_partition: bad arguments");
  offsets[0] = 0;
  offsets[nparts] = n;
  const int64_t base = n / nparts, rem = n % nparts;
  auto even = [&]() {
    for (int r = 0; r < nparts; ++r) offsets[r + 1] = offsets[r] + base + (r < rem ? 1 : 0);
  };
  nsamples = std::min<int64_t>(std::max<int64_t>(nsamples, 1), 4096);
  int64_t min_per_part = 1024;  // (test hook: B2CI_BALANCE_MIN = smallest mean block that is balanced)
  if (const char* env = getenv("B2CI_BALANCE_MIN")) min_per_part = std::max<int64_t>(1, atoll(env));
  nsamples = std::min(nsamples, std::max<int64_t>(1, n / 4));
  if (nparts == 1 || n < int64_t(nparts) * min_per_part) { even(); return; }
  cudaStream_t st = ctx->stream;
  // segment s = rows [s n / ns, (s + 1) n / ns), sampled at its middle
  std::vector<int64_t> seg(size_t(nsamples) + 1), rows(static_cast<size_t>(nsamples));
  for (int64_t s = 0; s <= nsamples; ++s) seg[size_t(s)] = int64_t((__int128)s * n / nsamples);
  for (int64_t s = 0; s < nsamples; ++s) rows[size_t(s)] = (seg[size_t(s)] + seg[size_t(s) + 1]) / 2;
  DevBuf<int64_t> d_rows(nsamples);
  DevBuf<unsigned long long> d_deg(nsamples);
  B2_CUDA(cudaMemcpyAsync(d_rows, rows.data(), size_t(nsamples) * 8, cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemsetAsync(d_deg, 0, size_t(nsamples) * 8, st));
  const dim3 grid(unsigned((nsamples + DEG_SPC - 1) / DEG_SPC), DEG_SLICES);
  k_sample_degree<<<grid, 256, 0, st>>>(dets->alpha, dets->beta, n, d_rows, int(nsamples), d_deg);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  std::vector<unsigned long long> deg(static_cast<size_t>(nsamples));
  B2_CUDA(cudaMemcpyAsync(deg.data(), d_deg, size_t(nsamples) * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  // cost of a row = its connections + a quarter of the mean (the part of the scan that does not depend on hits);
  // integer arithmetic throughout: every rank must arrive at the same cuts
  unsigned __int128 sum = 0;
  for (int64_t s = 0; s < nsamples; ++s) sum += (unsigned __int128)deg[size_t(s)] * (unsigned __int128)(seg[size_t(s) + 1] - seg[size_t(s)]);
  const unsigned long long floor_cost = (unsigned long long)(sum / (unsigned __int128)n / 4) + 1;
  std::vector<unsigned __int128> cum(size_t(nsamples) + 1, 0);
  for (int64_t s = 0; s < nsamples; ++s)
    cum[size_t(s) + 1] = cum[size_t(s)] + (unsigned __int128)(deg[size_t(s)] + floor_cost) * (unsigned __int128)(seg[size_t(s) + 1] - seg[size_t(s)]);
  const unsigned __int128 total = cum[size_t(nsamples)];
  int64_t s = 0;
  for (int r = 1; r < nparts; ++r) {
    const unsigned __int128 target = total * (unsigned)r / (unsigned)nparts;
    while (s + 1 < nsamples && cum[size_t(s) + 1] <= target) ++s;
    const unsigned __int128 per_row = deg[size_t(s)] + floor_cost;
    int64_t cut = seg[size_t(s)] + int64_t((target - cum[size_t(s)]) / per_row);
    cut = std::min(cut, seg[size_t(s) + 1]);
    cut = std::max(cut, offsets[r - 1]);
    offsets[r] = std::min(cut, n);
  }
  ctx->timers["h_build.partition_max_over_mean_rows"] = 0.;
  int64_t mx = 0;
  for (int r = 0; r < nparts; ++r) mx = std::max(mx, offsets[r + 1] - offsets[r]);
  ctx->timers["h_build.partition_max_over_mean_rows"] = double(mx) * nparts / double(n);
}

void hbuild_csr(b2ci_ctx* ctx, const b2ci_dets* dets, int64_t row_begin, int64_t row_end,
                double thr, b2ci_csr* out) {
  if (!ctx->ints_dev) throw Error("b2ci_hbuild_csr: integrals not uploaded");
  const int64_t n = dets->n;
  if (row_begin < 0 || row_end < row_begin || row_end > n) throw Error("b2ci_hbuild_csr: bad row range");
  if (n >= (int64_t(1) << 30)) throw Error("b2ci_hbuild_csr: more than 2^30 determinants per list");
  if (!(thr >= 0.0)) throw Error("b2ci_hbuild_csr: h_thresh must be >= 0");
  const int64_t nrows = row_end - row_begin;
  cudaStream_t st = ctx->stream;
  ctx->timers["h_build.setup"] = ctx->timers["h_build.count"] = ctx->timers["h_build.fill"] = 0.;
  ctx->timers["h_build.thresh"] = 0.;
  ctx->timers["h_build.fill_kernel"] = 0.;

  DevBuf<int32_t> run_of(n > 0 ? n : 1);
  DevBuf<int64_t> run_start, adj_ptr;
  DevBuf<uint64_t> run_alpha;
  DevBuf<uint32_t> adj;
  DevBuf<int32_t> run_deg;  // per run: neighbours at alpha distance 0 / 2 / 4
  int32_t nruns = 0;
  out->nrows = nrows;
  out->ncols = n;
  out->row_begin = row_begin;
  DevBuf<int64_t> rowptr(nrows + 1);
  if (n == 0 || nrows == 0) {
    B2_CUDA(cudaMemsetAsync(rowptr, 0, (nrows + 1) * 8, st));
    B2_CUDA(cudaStreamSynchronize(st));
    out->nnz = 0;
    out->rowptr = rowptr.take();
    return;
  }
  DeferredTimers DT(ctx);  // phase events are resolved at the (few) host synchronisations
  // host wall-clock trace of the orchestration (B2CI_HBUILD_TRACE=1): where the time between
  // kernels goes (allocations, synchronisations, launch gaps)
  const bool trace = getenv("B2CI_HBUILD_TRACE") != nullptr;
  auto t_host0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[hbuild] %-28s +%8.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_host0).count());
    t_host0 = t;
  };
  bool rect = false;
  {
    DeferredScope t(DT, "h_build.setup");
    DevBuf<int32_t> flag(n), excl(n + 1);
    DevBuf<int> bad(1);
    const unsigned gb = unsigned((n + 255) / 256);
    k_run_flags<<<gb, 256, 0, st>>>(dets->alpha, n, flag);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32(ctx, flag, excl, n);
    // run tables are sized for the worst case (every determinant its own run): the run count
    // and the shape test come back in one synchronisation
    run_start.alloc(n + 1);
    run_alpha.alloc(n);
    k_run_scatter<<<gb, 256, 0, st>>>(dets->alpha, n, flag, excl, run_of, run_start, run_alpha);
    B2_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), st));
    k_check_rect<<<gb, 256, 0, st>>>(dets->alpha, dets->beta, n, excl.p + n, bad);
    ctx->launches += 2;
    B2_CHECK_LAUNCH();
    int64_t* pin = pinned_words(ctx);
    B2_CUDA(cudaMemcpyAsync(pin, excl.p + n, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaMemcpyAsync(pin + 1, bad, sizeof(int), cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    nruns = *reinterpret_cast<const int32_t*>(pin);
    const int hbad = *reinterpret_cast<const int*>(pin + 1);
    // the product enumeration implements the sorted_double_loop rules; lists built under the
    // pair-based generators' rules take the row scan
    rect = hbad == 0 && nruns > 0 && !getenv("B2CI_HBUILD_FORCE_SCAN") && ctx->generator == 0;
    mark("runs + shape test (sync A)");
  }
  // run adjacency (count, scan, fill): distance <= 2 for the scan path
  auto build_adjacency = [&](int maxd) {
    DeferredScope t(DT, "h_build.setup");
    DevBuf<int32_t> acnt(nruns);
    adj_ptr.alloc(nruns + 1);
    const unsigned ga = unsigned((int64_t(nruns) * 32 + 255) / 256);
    const int skipz = ctx->generator == 0 ? 1 : 0;
    // the runs of the first and the last row of the block bound the runs whose adjacency is needed
    int32_t r_lo = 0, r_hi = nruns;
    if (row_begin > 0 || row_end < n) {
      int32_t h[2] = {0, 0};
      B2_CUDA(cudaMemcpyAsync(&h[0], run_of.p + row_begin, 4, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(&h[1], run_of.p + (row_end - 1), 4, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      r_lo = h[0];
      r_hi = h[1] + 1;
    }
    k_string_adjacency<false><<<ga, 256, 0, st>>>(run_alpha, nruns, maxd, skipz, acnt, nullptr, nullptr, nullptr, r_lo, r_hi);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, acnt, adj_ptr, nruns);
    int64_t nadj = 0;
    B2_CUDA(cudaMemcpyAsync(&nadj, adj_ptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    adj.alloc(nadj > 0 ? nadj : 1);
    k_string_adjacency<true><<<ga, 256, 0, st>>>(run_alpha, nruns, maxd, skipz, nullptr, nullptr, adj_ptr, adj, r_lo, r_hi);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  };

  // ---- rectangular (FCI-shaped) lists: product enumeration, no beta scan
  const int64_t nb = rect ? n / nruns : 0;
  ctx->timers["h_build.rectangular"] = rect ? 1. : 0.;
  if (rect) {
    DevBuf<int64_t> b2_ptr(nb + 1), b4_ptr(nb + 1);
    DevBuf<uint32_t> b2, b4;
    const unsigned ga = unsigned((int64_t(nruns) * 32 + 255) / 256);
    const unsigned gb = unsigned((nb * 32 + 255) / 256);
    int64_t nadj_h = 0, nb2_h = 0, nb4_h = 0;
    {
      // the three adjacency counts (alpha runs at distance <= 4, beta strings at <= 2 and <= 4)
      // go out together and their totals come back in one synchronisation
      DeferredScope t(DT, "h_build.setup");
      DevBuf<int32_t> acnt(nruns), bc2(nb), bc4(nb);
      adj_ptr.alloc(nruns + 1);
      k_string_adjacency<false><<<ga, 256, 0, st>>>(run_alpha, nruns, 4, 1, acnt, nullptr, nullptr, nullptr);
      k_string_adjacency<false><<<gb, 256, 0, st>>>(dets->beta, int32_t(nb), 2, 0, bc2, nullptr, nullptr, nullptr);
      k_string_adjacency<false><<<gb, 256, 0, st>>>(dets->beta, int32_t(nb), 4, 0, bc4, nullptr, nullptr, nullptr);
      ctx->launches += 3;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, acnt, adj_ptr, nruns);
      exclusive_scan_i32_to_i64(ctx, bc2, b2_ptr, nb);
      exclusive_scan_i32_to_i64(ctx, bc4, b4_ptr, nb);
      int64_t* pin = pinned_words(ctx);
      B2_CUDA(cudaMemcpyAsync(pin, adj_ptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 1, b2_ptr.p + nb, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 2, b4_ptr.p + nb, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaStreamSynchronize(st));
      nadj_h = pin[0]; nb2_h = pin[1]; nb4_h = pin[2];
      mark("adjacency counts (sync B)");
      adj.alloc(nadj_h > 0 ? nadj_h : 1);
      b2.alloc(nb2_h > 0 ? nb2_h : 1);
      b4.alloc(nb4_h > 0 ? nb4_h : 1);
      k_string_adjacency<true><<<ga, 256, 0, st>>>(run_alpha, nruns, 4, 1, nullptr, nullptr, adj_ptr, adj);
      k_string_adjacency<true><<<gb, 256, 0, st>>>(dets->beta, int32_t(nb), 2, 0, nullptr, nullptr, b2_ptr, b2);
      k_string_adjacency<true><<<gb, 256, 0, st>>>(dets->beta, int32_t(nb), 4, 0, nullptr, nullptr, b4_ptr, b4);
      ctx->launches += 3;
      B2_CHECK_LAUNCH();
    }
    // per-pair metadata (values of same-spin doubles, hole/particle/sign/leading sum of singles)
    DevBuf<double> b4_val(nb4_h > 0 ? nb4_h : 1), b2_val(nb2_h > 0 ? nb2_h : 1);
    DevBuf<uint32_t> b2_meta(nb2_h > 0 ? nb2_h : 1);
    DevBuf<B2Rec> b2rec(nb2_h > 0 ? nb2_h : 1);
    DevBuf<int32_t> run_cnt(size_t(nruns) * 4);
    DevBuf<int64_t> cptr(nruns + 1), sptr(nruns + 1);
    DevBuf<ARec> crec;
    DevBuf<double> cval, slead, diag(nrows);
    DevBuf<uint32_t> smeta;
    DevBuf<uint32_t> a_meta(nadj_h > 0 ? nadj_h : 1), b4_meta(nb4_h > 0 ? nb4_h : 1);
    DevBuf<double> a_val(nadj_h > 0 ? nadj_h : 1);
    DevBuf<int32_t> ecnt(nruns), scnt(nruns);
    DevBuf<unsigned int> small_cnt(1);
    DevBuf<int> shape(6);
    {
      DeferredScope t(DT, "h_build.setup");
      DevBuf<unsigned char> dead_ov(size_t(ctx->norb) * ctx->norb);
      const int nn = ctx->norb * ctx->norb;
      // how sparse are the integrals under this threshold? decides between the compacting and the dense fill
      B2_CUDA(cudaMemsetAsync(small_cnt, 0, sizeof(unsigned int), st));
      const int64_t n4 = int64_t(nn) * nn;
      k_count_small<<<unsigned((n4 + 255) / 256), 256, 0, st>>>(ctx->ints.V, n4, thr, small_cnt);
      ctx->launches++;
      k_dead_ov<<<(nn + 127) / 128, 128, 0, st>>>(ctx->ints, thr, dead_ov);
      k_pair_meta<<<ga, 256, 0, st>>>(ctx->ints, run_alpha, nruns, adj_ptr, adj, thr, dead_ov, a_meta, a_val);
      k_pair_meta<<<gb, 256, 0, st>>>(ctx->ints, dets->beta, int32_t(nb), b2_ptr, b2, thr, nullptr, b2_meta, b2_val);
      k_pair_meta<<<gb, 256, 0, st>>>(ctx->ints, dets->beta, int32_t(nb), b4_ptr, b4, thr, nullptr, b4_meta, b4_val);
      k_beta_rec<<<gb, 256, 0, st>>>(ctx->norb, int32_t(nb), b2_ptr, b2, b2_meta, b2rec);
      k_adj_compact<false><<<ga, 256, 0, st>>>(nruns, adj_ptr, adj, a_meta, a_val, ecnt, scnt, run_cnt, nullptr,
                                              nullptr, nullptr, nullptr, nullptr, nullptr);
      ctx->launches += 6;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, ecnt, cptr, nruns);
      exclusive_scan_i32_to_i64(ctx, scnt, sptr, nruns);
      const int init[6] = {INT32_MAX, 0, INT32_MAX, 0, 0, 0};
      B2_CUDA(cudaMemcpyAsync(shape, init, sizeof(init), cudaMemcpyHostToDevice, st));
      const int64_t nmx = std::max<int64_t>(nb, nruns);
      k_shape_minmax<<<unsigned((nmx + 255) / 256), 256, 0, st>>>(b2_ptr, b4_ptr, nb, cptr, run_cnt, nruns, shape);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    ProdArgs P;
    P.I = ctx->ints;
    P.run_alpha = run_alpha;
    P.tmpl_beta = dets->beta;
    P.cptr = cptr;
    P.sptr = sptr;
    P.run_cnt = run_cnt;
    P.b2_ptr = b2_ptr; P.b2rec = b2rec; P.b2_meta = b2_meta; P.b2_val = b2_val;
    P.b4_ptr = b4_ptr; P.b4 = b4; P.b4_val = b4_val;
    P.diag = diag;
    P.nb = nb;
    P.row_begin = row_begin;
    P.nrows = nrows;
    P.thr = thr;
    // per-warp scratch for the single-excitation elements of a row: at most
    // nocc * nvirt <= floor(n/2) * ceil(n/2) singles per spin, + 1 for the self slot of B2(k)
    P.smem_a = (ctx->norb / 2) * (ctx->norb - ctx->norb / 2) + 1;
    P.smem_b = P.smem_a;
    P.rowptr = nullptr; P.colind = nullptr; P.nzval = nullptr;
    P.crec = nullptr; P.cval = nullptr; P.slead = nullptr; P.smeta = nullptr;
    DevBuf<int64_t> slot_ptr(nrows + 1);
    DevBuf<int32_t> struct_cnt(nrows);
    int64_t nslots = 0, ncadj = 0, nsing = 0;
    unsigned int nsmall = 0;
    int shp[6] = {0, 0, 0, 0, 0, 0};
    uint64_t first_alpha = 0, first_beta = 0;  // electrons per spin of a full-CI list = popcount of any string
    int32_t rc[4] = {0, 0, 0, 0};
    int64_t bp[2] = {0, 0}, bq[2] = {0, 0};
    {
      // structural row lengths and diagonal elements need only the run counts: they are queued
      // behind the metadata kernels and everything the host must know comes back together
      DeferredScope t(DT, "h_build.count");
      P.row_cnt = struct_cnt;
      k_prod_struct_count<<<unsigned((nrows + 255) / 256), 256, 0, st>>>(P);
      k_row_diag<<<unsigned((nrows + 127) / 128), 128, 0, st>>>(P, diag);
      ctx->launches += 2;
      B2_CHECK_LAUNCH();
      exclusive_scan_i32_to_i64(ctx, struct_cnt, slot_ptr, nrows);
      int64_t* pin = pinned_words(ctx);
      B2_CUDA(cudaMemcpyAsync(pin, slot_ptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 1, cptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 2, sptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 4, run_cnt.p, 16, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 6, b2_ptr.p, 16, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 8, b4_ptr.p, 16, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 10, small_cnt.p, sizeof(unsigned int), cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 11, shape.p, 6 * sizeof(int), cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 15, run_alpha.p + row_begin / nb, 8, cudaMemcpyDeviceToHost, st));
      B2_CUDA(cudaMemcpyAsync(pin + 16, dets->beta, 8, cudaMemcpyDeviceToHost, st));
      mark("metadata + count queued");
      B2_CUDA(cudaStreamSynchronize(st));
      nslots = pin[0]; ncadj = pin[1]; nsing = pin[2];
      nsmall = *reinterpret_cast<const unsigned int*>(pin + 10);
      memcpy(shp, pin + 11, 6 * sizeof(int));
      first_alpha = uint64_t(pin[15]);
      first_beta = uint64_t(pin[16]);
      memcpy(rc, pin + 4, 16); memcpy(bp, pin + 6, 16); memcpy(bq, pin + 8, 16);
      mark("metadata + count (sync C)");
    }
    crec.alloc(ncadj > 0 ? ncadj : 1);
    cval.alloc(ncadj > 0 ? ncadj : 1);
    slead.alloc(nsing > 0 ? nsing : 1);
    smeta.alloc(nsing > 0 ? nsing : 1);
    {
      DeferredScope t(DT, "h_build.setup");
      k_adj_compact<true><<<ga, 256, 0, st>>>(nruns, adj_ptr, adj, a_meta, a_val, nullptr, nullptr, nullptr, cptr,
                                             crec, cval, sptr, slead, smeta);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    P.crec = crec; P.cval = cval; P.slead = slead; P.smeta = smeta;
    // structural slots: recycled through the context (a freed matrix or the slots of a previous
    // thresholded build); a pool allocation of this size costs ~0.3 ms of host time
    DevBuf<int32_t> ci_s, kept(nrows);
    DevBuf<double> nz_s;
    size_t ci_cap = 0, nz_cap = 0;
    ci_s.p = static_cast<int32_t*>(big_alloc(ctx, 0, size_t(nslots > 0 ? nslots : 1) * sizeof(int32_t), &ci_cap));
    ci_s.n = ci_cap / sizeof(int32_t);
    nz_s.p = static_cast<double*>(big_alloc(ctx, 1, size_t(nslots > 0 ? nslots : 1) * sizeof(double), &nz_cap));
    nz_s.n = nz_cap / sizeof(double);
    mark("slot allocation");
    bool dense = false;
    DevBuf<uint4> dn_tab;
    DevBuf<uint32_t> dn_desc;
    DevBuf<int> dn_cnt;
    DevBuf<uint2> dn_rec;
    DevBuf<double> dn_sab;
    {
      DeferredScope t(DT, "h_build.fill");
      P.row_cnt = kept;
      P.rowptr = slot_ptr;
      P.colind = ci_s;
      P.nzval = nz_s;
      // lanes per row: the group width that wastes the fewest lane slots on this list shape
      int G = 32;
      {
        const double l2 = double(bp[1] - bp[0]), l4 = double(bq[1] - bq[0]);
        const double nunit_runs = std::min<double>(rc[2], rc[0] + rc[1] + 1);
        const double unit_len = nunit_runs > 0 ? rc[2] / nunit_runs : 0.;
        double best = 0.;
        for (int g : {32, 16, 8}) {
          const double slots = g * (rc[1] * std::ceil(l2 / g) + rc[0] * std::ceil(l4 / g) +
                                    nunit_runs * std::ceil(std::max(1., unit_len) / g));
          // measured cost per lane slot on B200 (Cr2 CAS(12,12)): 8-lane groups write 64-byte
          // pieces and pay ~1.6x per slot, 16 and 32 lanes are level
          const double cost = slots * (g == 8 ? 1.58 : (g == 16 ? 1.0 : 0.96));
          if (best == 0. || cost < best) { best = cost; G = g; }
        }
        if (const char* env = getenv("B2CI_HBUILD_GROUP")) {
          const int g = atoi(env);
          if (g == 8 || g == 16 || g == 32) G = g;
        }
      }
      ctx->timers["h_build.group_width"] = G;
      const int rpc = PW * (32 / G);
      const int64_t cpr = (nb + rpc - 1) / rpc;
      const int64_t nruns_blk = (row_end - 1) / nb - row_begin / nb + 1;
      const unsigned grid = unsigned(nruns_blk * cpr);
      // shared memory: mbarrier + integral slices + per-row singles and B2 records
      P.nslice_max = P.smem_a - 1;
      P.smem_r = (P.smem_b + G - 1) / G * G;
      const size_t rows_bytes = size_t(rpc) * ((P.smem_a + P.smem_b) * sizeof(double) + P.smem_r * sizeof(OsRec));
      const size_t slice_bytes = size_t(P.nslice_max) * ctx->ints.n2p * sizeof(double);
      // stage the slices when two CTAs per SM still fit (227 KB per SM)
      bool slices = 16 + slice_bytes + rows_bytes <= 110 * 1024;
      if (const char* env = getenv("B2CI_HBUILD_SLICES")) slices = atoi(env) != 0 && 16 + slice_bytes + rows_bytes <= 220 * 1024;
      ctx->timers["h_build.smem_slices"] = slices ? 1. : 0.;
      const size_t smem = 16 + (slices ? slice_bytes : 0) + rows_bytes;
      if (smem > 220 * 1024) throw Error("b2ci_hbuild_csr: product path needs " + std::to_string(smem) + " bytes of shared memory per CTA");
      auto launch = [&](auto kern) {
        if (smem > 48 * 1024)
          B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        // the largest carve-out: resident CTAs are then limited by registers, not by the split
        B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
        kern<<<grid, PW * 32, smem, st>>>(P);
      };
      auto launch_g = [&](auto ev, auto sl, auto dn) {
        constexpr bool EV = decltype(ev)::value, SL = decltype(sl)::value, DN = decltype(dn)::value;
        if (G == 8) launch(k_rows_product<EV, 8, SL, DN>);
        else if (G == 16) launch(k_rows_product<EV, 16, SL, DN>);
        else launch(k_rows_product<EV, 32, SL, DN>);
      };
      // dense fill (structural positions, drops only counted): when fewer than 2 % of the two-electron
      // integrals fail the threshold. Sparse models (Hubbard: 99 % zeros) keep the compacting fill, which
      // writes only what survives.
      const int64_t n4 = int64_t(ctx->norb) * ctx->norb * ctx->norb * ctx->norb;
      dense = slices && double(nsmall) < 0.02 * double(n4);
      // B2CI_HBUILD_DENSE: 0 = compacting fill, 1 = dense emit in the product kernel, 2 (default where it
      // applies) = the specialised kernel for uniform lists
      int dense_mode = dense ? 2 : 0;
      if (const char* env = getenv("B2CI_HBUILD_DENSE")) dense_mode = slices ? atoi(env) : 0;
      dense = dense_mode != 0;
      const bool uniform = shp[0] == shp[1] && shp[2] == shp[3] && shp[0] > 0 && ctx->ints.n2p <= 4096;
      const int dense_per_row = (P.smem_a + P.smem_b + 1) & ~1;  // doubles of (sa | sb) per row, 16-byte multiple
      // structural row length of the list (uniform): self x len4 + live singles x len2 + units of the longest run
      const int dense_desc_max = (shp[5] + 3) & ~3;  // longest structural row, 16-byte multiple
      const size_t dense_smem = 32 + slice_bytes + size_t(2 * shp[4] + 2) * 16 + size_t(dense_desc_max) * 4 +
                                ((size_t(DROWS) * (shp[0] + 2) * 8 + 15) & ~size_t(15)) + size_t(DROWS) * dense_per_row * sizeof(double);
      if (dense_mode == 2 && (!uniform || dense_smem > 110 * 1024 || 2 * shp[4] + 2 >= 4095 || shp[0] + 2 >= 4095 ||
                              nb * dense_per_row >= (int64_t(1) << 24)))
        dense_mode = 1;
      ctx->timers["h_build.dense_fill"] = double(dense_mode);
      if (dense_mode == 2) {
        DenseArgs DA;
        DA.P = P;
        DA.len2 = shp[0];
        DA.len4 = shp[2];
        DA.per_row = dense_per_row;
        DA.run0 = row_begin / nb;
        const int64_t nrows_tot = nruns_blk * nb;
        DA.desc_max = dense_desc_max;
        DA.tab_off = int(32 + slice_bytes);
        DA.rec_stride = shp[0] + 2;
        DA.tab_max = 2 * shp[4] + 2;
        dn_desc.alloc(size_t(nruns_blk) * DA.desc_max);
        DA.desc = dn_desc;
        dn_tab.alloc(size_t(nruns_blk) * DA.tab_max);
        dn_cnt.alloc(size_t(nruns_blk) * 4);
        dn_rec.alloc(size_t(nb) * DA.rec_stride + 2 * DROWS);  // padding: the last chunk's copy is rounded up to 16 B
        dn_sab.alloc(size_t(nrows_tot) * DA.per_row);
        DA.tab = dn_tab; DA.tab_cnt = dn_cnt; DA.rec = dn_rec; DA.sab = dn_sab;
        k_dense_tables<<<unsigned((nruns_blk * 32 + 127) / 128), 128, 0, st>>>(DA, nruns_blk);
        k_dense_rec<<<unsigned((nb * DA.rec_stride + 255) / 256), 256, 0, st>>>(DA);
        {
          const unsigned inv_per = unsigned((uint64_t(1) << 32) / unsigned(DA.len2)) + 1u;  // f / len2 by umulhi
          const dim3 gs(unsigned((nb + SAB_ROWS - 1) / SAB_ROWS), unsigned(nruns_blk));
          const size_t vr_bytes = size_t(ctx->norb) * ctx->norb * (ctx->norb | 1) * 8;
          // electrons per spin (all strings of a full-CI list have the same count): exact unrolling of the sums
          const int na_e = __builtin_popcountll(first_alpha), nb_e = __builtin_popcountll(first_beta);
          if (vr_bytes <= 46 * 1024) {
            if (ctx->norb <= 32) {
              launch_dense_sab<uint32_t, true, 0>(na_e, gs, vr_bytes, st, DA, nrows_tot, inv_per);
              launch_dense_sab<uint32_t, true, 1>(nb_e, gs, vr_bytes, st, DA, nrows_tot, inv_per);
            } else {
              launch_dense_sab<uint64_t, true, 0>(0, gs, vr_bytes, st, DA, nrows_tot, inv_per);
              launch_dense_sab<uint64_t, true, 1>(0, gs, vr_bytes, st, DA, nrows_tot, inv_per);
            }
          } else {
            if (ctx->norb <= 32) {
              launch_dense_sab<uint32_t, false, 0>(0, gs, 0, st, DA, nrows_tot, inv_per);
              launch_dense_sab<uint32_t, false, 1>(0, gs, 0, st, DA, nrows_tot, inv_per);
            } else {
              launch_dense_sab<uint64_t, false, 0>(0, gs, 0, st, DA, nrows_tot, inv_per);
              launch_dense_sab<uint64_t, false, 1>(0, gs, 0, st, DA, nrows_tot, inv_per);
            }
          }
          ctx->launches++;
        }
        ctx->launches += 3;
        B2_CHECK_LAUNCH();
        const int64_t dcpr = (nb + DROWS - 1) / DROWS;
        const unsigned dgrid = unsigned(nruns_blk * dcpr);
        auto dl = [&](auto kern) {
          if (dense_smem > 48 * 1024)
            B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(dense_smem)));
          B2_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
          kern<<<dgrid, DW * 32, dense_smem, st>>>(DA);
        };
        {
          DeferredScope tk(DT, "h_build.fill_kernel");  // the dominant kernel alone (bench.py's roofline line)
          if (thr > 0.0) dl(k_rows_dense<true>);
          else dl(k_rows_dense<false>);
        }
      } else if (thr > 0.0) {
        if (dense) launch_g(std::true_type{}, std::true_type{}, std::true_type{});
        else if (slices) launch_g(std::true_type{}, std::true_type{}, std::false_type{});
        else launch_g(std::true_type{}, std::false_type{}, std::false_type{});
      } else {
        if (dense) launch_g(std::false_type{}, std::true_type{}, std::true_type{});
        else if (slices) launch_g(std::false_type{}, std::true_type{}, std::false_type{});
        else launch_g(std::false_type{}, std::false_type{}, std::false_type{});
      }
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    mark("fill queued");
    int64_t nnz = nslots;
    if (thr > 0.0) {
      // threshold_parallel equivalent: rows were written compacted inside their structural
      // slots; when something was dropped, pack the rows (csr_matrix.hpp:317-370)
      size_t tt = DT.start("h_build.thresh");
      exclusive_scan_i32_to_i64(ctx, kept, rowptr, nrows);
      int64_t* pin = pinned_words(ctx);
      B2_CUDA(cudaMemcpyAsync(pin, rowptr.p + nrows, 8, cudaMemcpyDeviceToHost, st));
      DT.stop(tt);
      B2_CUDA(cudaStreamSynchronize(st));
      nnz = pin[0];
      mark("fill + nnz (sync E)");
      if (nnz != nslots) {
        DeferredScope t(DT, "h_build.thresh");
        DevBuf<int32_t> ci_f(nnz > 0 ? nnz : 1);
        DevBuf<double> nz_f(nnz > 0 ? nnz : 1);
        const unsigned gc = unsigned((nrows * 32 + ROW_WARPS * 32 - 1) / (ROW_WARPS * 32));
        if (dense) k_compact_rows_filter<<<gc, ROW_WARPS * 32, 0, st>>>(nrows, slot_ptr, rowptr, ci_s, nz_s, thr, ci_f, nz_f);
        else k_compact_rows<<<gc, ROW_WARPS * 32, 0, st>>>(nrows, slot_ptr, rowptr, ci_s, nz_s, ci_f, nz_f);
        ctx->launches++;
        B2_CHECK_LAUNCH();
        B2_CUDA(cudaStreamSynchronize(st));
        // the structural arrays stay with the context for the next build
        big_release(ctx, 0, ci_s.take(), ci_cap);
        big_release(ctx, 1, nz_s.take(), nz_cap);
        ci_cap = nz_cap = 0;  // the packed arrays are ordinary pool allocations
        ci_s = std::move(ci_f);
        nz_s = std::move(nz_f);
      }
    } else {
      rowptr = std::move(slot_ptr);
      B2_CUDA(cudaStreamSynchronize(st));
    }
    DT.resolve();
    mark("compaction + timers");
    out->nnz = nnz;
    out->rowptr = rowptr.take();
    out->colind = ci_s.take();
    out->nzval = nz_s.take();
    out->colind_cap = ci_cap;
    out->nzval_cap = nz_cap;
    return;
  }
  build_adjacency(2);
  if (trace) {
    // shape of a general list: how the rows are distributed over alpha runs of which length
    std::vector<int64_t> rs(size_t(nruns) + 1);
    B2_CUDA(cudaMemcpyAsync(rs.data(), run_start.p, (size_t(nruns) + 1) * 8, cudaMemcpyDeviceToHost, st));
    int64_t nadj_h = 0;
    B2_CUDA(cudaMemcpyAsync(&nadj_h, adj_ptr.p + nruns, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    const int64_t edges[] = {1, 2, 4, 8, 16, 32, 64, 128, 256, 1024, 4096, int64_t(1) << 40};
    int64_t rows_in[12] = {0}, runs_in[12] = {0};
    for (int32_t r = 0; r < nruns; ++r) {
      const int64_t len = rs[size_t(r) + 1] - rs[r];
      int b = 0;
      while (len > edges[b]) ++b;
      rows_in[b] += len;
      runs_in[b] += 1;
    }
    fprintf(stderr, "[hbuild] general list: n = %lld, alpha runs = %d, run adjacency entries = %lld\n", (long long)n, nruns,
            (long long)nadj_h);
    for (int b = 0; b < 12; ++b)
      if (runs_in[b])
        fprintf(stderr, "[hbuild]   runs of length <= %-6lld : %8lld runs, %9lld rows (%.1f %%)\n",
                (long long)std::min<int64_t>(edges[b], n), (long long)runs_in[b], (long long)rows_in[b], 100.0 * rows_in[b] / n);
  }

  // ---- general lists: determinants grouped by beta string (stable radix sort of the indices)
  DevBuf<int32_t> bgrp_of(n);
  DevBuf<int64_t> bgrp_start;
  DevBuf<uint32_t> bgrp_mem(n), idx_alt(n);
  {
    ScopedTimer t(ctx, "h_build.setup", true);
    DevBuf<uint64_t> key(n), key_alt(n);
    B2_CUDA(cudaMemcpyAsync(key, dets->beta, size_t(n) * 8, cudaMemcpyDeviceToDevice, st));
    iota_u32(ctx, bgrp_mem, n);
    std::vector<int> shifts;
    for (int d = 0; d < (ctx->norb + 7) / 8; ++d) shifts.push_back(8 * d);
    radix_sort_pairs(ctx, key, key_alt, bgrp_mem, idx_alt, n, shifts);
    DevBuf<int32_t> flag(n), excl(n + 1);
    const unsigned gb = unsigned((n + 255) / 256);
    k_run_flags<<<gb, 256, 0, st>>>(key, n, flag);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32(ctx, flag, excl, n);
    int32_t ngroups = 0;
    B2_CUDA(cudaMemcpyAsync(&ngroups, excl.p + n, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    bgrp_start.alloc(size_t(ngroups) + 1);
    k_group_scatter<<<gb, 256, 0, st>>>(key, bgrp_mem, n, flag, excl, bgrp_of, bgrp_start, ngroups);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }
  RowArgs A;
  A.I = ctx->ints;
  A.alpha = dets->alpha;
  A.beta = dets->beta;
  A.run_of = run_of;
  A.run_start = run_start;
  A.adj_ptr = adj_ptr;
  A.adj = adj;
  A.bgrp_of = bgrp_of;
  A.bgrp_start = bgrp_start;
  A.bgrp_mem = bgrp_mem;
  A.row_begin = row_begin;
  A.nrows = nrows;
  A.thr = thr;
  A.row_cnt = nullptr;
  A.rowptr = nullptr;
  A.colind = nullptr;
  A.nzval = nullptr;
  A.pair_rule = ctx->generator != 0;
  // count pass (keeps the connections it finds) + fill pass (evaluates them): run_row_scan
  ctx->timers["h_build.count"] = ctx->timers["h_build.fill"] = ctx->timers["h_build.thresh"] = 0.;
  RowScanOut R;
  run_row_scan<false>(ctx, A, nrows, n, thr, true, R);
  out->colind_cap = R.ci_cap;
  out->nzval_cap = R.nz_cap;
  out->nnz = R.nnz;
  out->rowptr = R.rowptr.take();
  out->colind = R.colind.take();
  out->nzval = R.nzval.take();
}

// ------------------------------------------------------------------------------------
// Rectangular block over two alpha-grouped lists (make_csr_hamiltonian_block with bra != ket,
// csr_hamiltonian.hpp:41-72), rows = bra determinants, columns = ket determinants renumbered by
// gmap (their index in a common list, which also decides the bra/ket roles of every element).
namespace {
struct DetView {
  const uint64_t* alpha;
  const uint64_t* beta;
  int64_t n;
  const int32_t* gmap;  // may be NULL: the list is the common list
};
struct BlockOut {
  DevBuf<int64_t> rowptr;
  DevBuf<int32_t> colind;
  DevBuf<double> nzval;
  int64_t nnz = 0;
};
int32_t run_tables(b2ci_ctx* ctx, const DetView& L, DevBuf<int32_t>& run_of, DevBuf<int64_t>& run_start,
                   DevBuf<uint64_t>& run_alpha) {
  cudaStream_t st = ctx->stream;
  DevBuf<int32_t> flag(L.n), excl(L.n + 1);
  const unsigned gb = unsigned((L.n + 255) / 256);
  k_run_flags<<<gb, 256, 0, st>>>(L.alpha, L.n, flag);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  exclusive_scan_i32(ctx, flag, excl, L.n);
  run_of.alloc(L.n);
  run_start.alloc(L.n + 1);
  run_alpha.alloc(L.n);
  k_run_scatter<<<gb, 256, 0, st>>>(L.alpha, L.n, flag, excl, run_of, run_start, run_alpha);
  ctx->launches++;
  B2_CHECK_LAUNCH();
  int32_t nruns = 0;
  B2_CUDA(cudaMemcpyAsync(&nruns, excl.p + L.n, 4, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  return nruns;
}
void build_block_general(b2ci_ctx* ctx, const DetView& bra, const DetView& ket, double thr, BlockOut& out) {
  cudaStream_t st = ctx->stream;
  const int64_t nrows = bra.n, nk = ket.n;
  out.rowptr.alloc(nrows + 1);
  out.nnz = 0;
  if (nrows == 0 || nk == 0) {
    B2_CUDA(cudaMemsetAsync(out.rowptr, 0, size_t(nrows + 1) * 8, st));
    out.colind.alloc(1);
    out.nzval.alloc(1);
    return;
  }
  DevBuf<int32_t> bra_run, ket_run_of;
  DevBuf<int64_t> bra_run_start, ket_run_start, adj_ptr;
  DevBuf<uint64_t> bra_run_alpha, ket_run_alpha;
  const int32_t nrb = run_tables(ctx, bra, bra_run, bra_run_start, bra_run_alpha);
  const int32_t nrk = run_tables(ctx, ket, ket_run_of, ket_run_start, ket_run_alpha);
  // bra-run x ket-run adjacency, alpha distance <= 2
  DevBuf<uint32_t> adj;
  const int skipz = ctx->generator == 0 ? 1 : 0;
  {
    DevBuf<int32_t> acnt(nrb);
    adj_ptr.alloc(size_t(nrb) + 1);
    const unsigned ga = unsigned((int64_t(nrb) * 32 + 255) / 256);
    k_string_adjacency2<false><<<ga, 256, 0, st>>>(bra_run_alpha, nrb, ket_run_alpha, nrk, 2, skipz, acnt, nullptr, nullptr);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, acnt, adj_ptr, nrb);
    int64_t nadj = 0;
    B2_CUDA(cudaMemcpyAsync(&nadj, adj_ptr.p + nrb, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    adj.alloc(nadj > 0 ? nadj : 1);
    k_string_adjacency2<true><<<ga, 256, 0, st>>>(bra_run_alpha, nrb, ket_run_alpha, nrk, 2, skipz, nullptr, adj_ptr, adj);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }
  // ket determinants grouped by beta string; the group of every bra determinant's beta string
  DevBuf<int32_t> bgrp_of(nk), bra_grp(nrows);
  DevBuf<int64_t> bgrp_start;
  DevBuf<uint32_t> bgrp_mem(nk), idx_alt(nk);
  DevBuf<uint64_t> key(nk), key_alt(nk);
  {
    B2_CUDA(cudaMemcpyAsync(key, ket.beta, size_t(nk) * 8, cudaMemcpyDeviceToDevice, st));
    iota_u32(ctx, bgrp_mem, nk);
    std::vector<int> shifts;
    for (int d = 0; d < (ctx->norb + 7) / 8; ++d) shifts.push_back(8 * d);
    radix_sort_pairs(ctx, key, key_alt, bgrp_mem, idx_alt, nk, shifts);
    DevBuf<int32_t> flag(nk), excl(nk + 1);
    const unsigned gb = unsigned((nk + 255) / 256);
    k_run_flags<<<gb, 256, 0, st>>>(key, nk, flag);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32(ctx, flag, excl, nk);
    int32_t ngroups = 0;
    B2_CUDA(cudaMemcpyAsync(&ngroups, excl.p + nk, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    bgrp_start.alloc(size_t(ngroups) + 1);
    k_group_scatter<<<gb, 256, 0, st>>>(key, bgrp_mem, nk, flag, excl, bgrp_of, bgrp_start, ngroups);
    k_lookup_group<<<unsigned((nrows + 255) / 256), 256, 0, st>>>(bra.beta, nrows, key, bgrp_start, ngroups, bra_grp);
    ctx->launches += 2;
    B2_CHECK_LAUNCH();
  }
  RowArgs A;
  A.I = ctx->ints;
  A.alpha = ket.alpha;
  A.beta = ket.beta;
  A.run_of = ket_run_of;
  A.run_start = ket_run_start;
  A.adj_ptr = adj_ptr;
  A.adj = adj;
  A.bgrp_of = bgrp_of;
  A.bgrp_start = bgrp_start;
  A.bgrp_mem = bgrp_mem;
  A.row_begin = 0;
  A.nrows = nrows;
  A.thr = thr;
  A.row_cnt = nullptr; A.rowptr = nullptr; A.colind = nullptr; A.nzval = nullptr;
  A.bra_alpha = bra.alpha;
  A.bra_beta = bra.beta;
  A.bra_run = bra_run;
  A.bra_run_start = bra_run_start;
  A.bra_grp = bra_grp;
  A.rowmap = bra.gmap;
  A.colmap = ket.gmap;
  A.pair_rule = ctx->generator != 0;
  // the block's own phase times are folded into the patched build's timers by its caller
  const double t_count = ctx->timers["h_build.count"], t_fill = ctx->timers["h_build.fill"],
               t_thresh = ctx->timers["h_build.thresh"];
  RowScanOut R;
  run_row_scan<true>(ctx, A, nrows, nk, thr, false, R);
  ctx->timers["h_build.count"] = t_count;
  ctx->timers["h_build.fill"] = t_fill;
  ctx->timers["h_build.thresh"] = t_thresh;
  out.nnz = R.nnz;
  out.rowptr = std::move(R.rowptr);
  out.colind = std::move(R.colind);
  out.nzval = std::move(R.nzval);
}
}  // namespace

// Patched (incremental) build: the CSR of `nd` from the CSR of an overlapping list `od`
// (CachedHamiltonianState / build_patched_operator, solvers/incremental_h_build.hpp:192-356).
// The reference keeps the old matrix and applies three blocks (old-old with an index remap,
// added x added, kept x added and its transpose) inside the Davidson operator; on the device the
// blocks are merged into ONE ordinary CSR of the new list -- bit-identical to a full build,
// because every element is the same function of its determinant pair and the old-to-new map is
// monotone -- so sigma stays a single streaming SpMV and the result can seed the next patch.
//   kept x kept  : old rows, dropped columns removed, columns renumbered
//   kept x added : block build, bra = kept determinants, ket = added determinants
//   added x all  : block build, bra = added determinants, ket = the whole new list
// Returns false (nothing built) when n_kept / n_new < min_overlap (:266-283).
bool hbuild_csr_patched(b2ci_ctx* ctx, const b2ci_dets* od, const b2ci_csr* oH, const b2ci_dets* nd, double thr,
                        double min_overlap, b2ci_csr* out, int64_t* n_kept_out) {
  if (!ctx->ints_dev) throw Error("b2ci_hbuild_csr_patched: integrals not uploaded");
  if (ctx->nranks > 1) throw Error("b2ci_hbuild_csr_patched: not available with a communicator (row-sharded builds are full builds)");
  if (!od || !oH || !nd) throw Error("b2ci_hbuild_csr_patched: null argument");
  const int64_t n_old = od->n, n_new = nd->n;
  if (oH->row_begin != 0 || oH->nrows != n_old || oH->ncols != n_old)
    throw Error("b2ci_hbuild_csr_patched: the cached matrix must be the full square matrix of the old list");
  if (n_new >= (int64_t(1) << 30)) throw Error("b2ci_hbuild_csr_patched: more than 2^30 determinants per list");
  if (!(thr >= 0.0)) throw Error("b2ci_hbuild_csr_patched: h_thresh must be >= 0");
  cudaStream_t st = ctx->stream;
  auto& T = ctx->timers;
  T["h_build.setup"] = T["h_build.count"] = T["h_build.fill"] = T["h_build.thresh"] = 0.;
  T["h_build.patch_kept"] = T["h_build.patch_added"] = 0.;
  if (n_kept_out) *n_kept_out = 0;
  if (n_new == 0 || n_old == 0) return false;
  const bool trace = getenv("B2CI_HBUILD_TRACE") != nullptr;
  auto t_host0 = std::chrono::steady_clock::now();
  auto mark = [&](const char* what) {
    if (!trace) return;
    cudaStreamSynchronize(st);
    auto t = std::chrono::steady_clock::now();
    fprintf(stderr, "[hbuild patched] %-24s +%9.1f us\n", what, std::chrono::duration<double, std::micro>(t - t_host0).count());
    t_host0 = t;
  };
  // ---- classify (both lists spin-sorted)
  DevBuf<int32_t> new_to_old(n_new), old_to_new(n_old), kflag(n_new), aflag(n_new), kexcl(n_new + 1), aexcl(n_new + 1);
  int32_t n_kept = 0;
  {
    ScopedTimer t(ctx, "h_build.setup", true);
    B2_CUDA(cudaMemsetAsync(old_to_new, 0xFF, size_t(n_old) * 4, st));
    k_classify<<<unsigned((n_new + 255) / 256), 256, 0, st>>>(nd->alpha, nd->beta, n_new, od->alpha, od->beta, n_old,
                                                              new_to_old, old_to_new, kflag, aflag);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32(ctx, kflag, kexcl, n_new);
    exclusive_scan_i32(ctx, aflag, aexcl, n_new);
    B2_CUDA(cudaMemcpyAsync(&n_kept, kexcl.p + n_new, 4, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
  }
  if (n_kept_out) *n_kept_out = n_kept;
  const int64_t n_added = n_new - n_kept;
  T["h_build.patch_kept"] = double(n_kept);
  T["h_build.patch_added"] = double(n_added);
  mark("classify");
  if (double(n_kept) / double(n_new) < min_overlap) return false;
  DevBuf<int32_t> kept_new(n_kept > 0 ? n_kept : 1), added_new(n_added > 0 ? n_added : 1);
  DevBuf<uint64_t> ka(n_kept > 0 ? n_kept : 1), kb(n_kept > 0 ? n_kept : 1), aa(n_added > 0 ? n_added : 1),
      ab(n_added > 0 ? n_added : 1);
  DevBuf<int64_t> kptr(size_t(n_kept) + 1);
  DevBuf<int32_t> kci;
  DevBuf<double> knz;
  {
    ScopedTimer t(ctx, "h_build.setup", true);
    k_split_lists<<<unsigned((n_new + 255) / 256), 256, 0, st>>>(nd->alpha, nd->beta, n_new, kflag, kexcl, aexcl,
                                                                 kept_new, added_new, ka, kb, aa, ab);
    ctx->launches++;
    B2_CHECK_LAUNCH();
  }
  // ---- kept x kept from the cached matrix
  {
    ScopedTimer t(ctx, "h_build.count", true);
    DevBuf<int32_t> kcnt(n_kept > 0 ? n_kept : 1);
    const unsigned gk = unsigned((int64_t(n_kept) * 32 + ROW_WARPS * 32 - 1) / (ROW_WARPS * 32));
    int64_t nkk = 0;
    if (n_kept) {
      k_kept_rows<false><<<gk, ROW_WARPS * 32, 0, st>>>(n_kept, kept_new, new_to_old, old_to_new, oH->rowptr, oH->colind,
                                                       oH->nzval, kcnt, nullptr, nullptr, nullptr);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
    exclusive_scan_i32_to_i64(ctx, kcnt, kptr, n_kept);
    B2_CUDA(cudaMemcpyAsync(&nkk, kptr.p + n_kept, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    kci.alloc(nkk > 0 ? nkk : 1);
    knz.alloc(nkk > 0 ? nkk : 1);
    if (n_kept) {
      k_kept_rows<true><<<gk, ROW_WARPS * 32, 0, st>>>(n_kept, kept_new, new_to_old, old_to_new, oH->rowptr, oH->colind,
                                                      oH->nzval, nullptr, kptr, kci, knz);
      ctx->launches++;
      B2_CHECK_LAUNCH();
    }
  }
  mark("kept x kept");
  // ---- the two freshly evaluated blocks
  BlockOut DK, DA;
  if (n_added) {
    ScopedTimer t(ctx, "h_build.fill", true);
    const DetView kept_v{ka, kb, n_kept, kept_new}, added_v{aa, ab, n_added, added_new}, all_v{nd->alpha, nd->beta, n_new, nullptr};
    if (n_kept) build_block_general(ctx, kept_v, added_v, thr, DK);
    mark("kept x added");
    build_block_general(ctx, added_v, all_v, thr, DA);
    mark("added x all");
  }
  // ---- merge
  out->nrows = n_new;
  out->ncols = n_new;
  out->row_begin = 0;
  DevBuf<int64_t> rowptr(n_new + 1);
  DevBuf<int32_t> ci;
  DevBuf<double> nz;
  int64_t nnz = 0;
  {
    ScopedTimer t(ctx, "h_build.thresh", true);
    DevBuf<int32_t> cnt(n_new);
    const int64_t* dkp = (n_added && n_kept) ? DK.rowptr.p : nullptr;
    DevBuf<int64_t> zero_ptr;
    const int64_t* dap = DA.rowptr.p;
    if (!n_added) {  // nothing was added: every row is a kept row
      zero_ptr.alloc(1);
      B2_CUDA(cudaMemsetAsync(zero_ptr, 0, 8, st));
      dap = zero_ptr;
    }
    k_patch_count<<<unsigned((n_new + 255) / 256), 256, 0, st>>>(n_new, kflag, kexcl, aexcl, kptr, dkp, dap, cnt);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    exclusive_scan_i32_to_i64(ctx, cnt, rowptr, n_new);
    B2_CUDA(cudaMemcpyAsync(&nnz, rowptr.p + n_new, 8, cudaMemcpyDeviceToHost, st));
    B2_CUDA(cudaStreamSynchronize(st));
    mark("merge count");
    ci.alloc(nnz > 0 ? nnz : 1);
    nz.alloc(nnz > 0 ? nnz : 1);
    mark("merge alloc");
    k_patch_fill<<<unsigned((n_new * 32 + ROW_WARPS * 32 - 1) / (ROW_WARPS * 32)), ROW_WARPS * 32, 0, st>>>(
        n_new, kflag, kexcl, aexcl, kptr, kci, knz, dkp, DK.colind.p, DK.nzval.p, dap, DA.colind.p, DA.nzval.p, rowptr,
        ci, nz);
    ctx->launches++;
    B2_CHECK_LAUNCH();
    B2_CUDA(cudaStreamSynchronize(st));
    mark("merge fill");
  }
  out->nnz = nnz;
  out->rowptr = rowptr.take();
  out->colind = ci.take();
  out->nzval = nz.take();
  out->colind_cap = 0;
  out->nzval_cap = 0;
  return true;
}

}  // namespace b2ci
