// Integral tables on the device and their reduced intermediates.
// Replaces HamiltonianGeneratorBase<double>::generate_integral_intermediates_
// (external/macis/src/macis/hamiltonian_generator/base.ipp:37-77):
//   G_red(k,i,j) = V(k,k,i,j) - V(k,j,i,k)     V_red(k,i,j) = V(k,k,i,j)
//   G2_red(i,j)  = 0.5 * (V(i,i,j,j) - V(i,j,j,i))   V2_red(i,j) = V(i,i,j,j)
#include "common.cuh"

namespace b2ci {
namespace {
__global__ void k_intermediates(int n, const double* __restrict__ V, double* __restrict__ G,
                                double* __restrict__ Vr, double* __restrict__ G2,
                                double* __restrict__ V2) {
  const size_t n2 = size_t(n) * n, n3 = n2 * n;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t < n3) {
    const size_t k = t % n, i = (t / n) % n, j = t / n2;
    const double vkkij = V[k + k * n + i * n2 + j * n3];
    G[t] = vkkij - V[k + j * n + i * n2 + k * n3];
    Vr[t] = vkkij;
  }
  if (t < n2) {
    const size_t i = t % n, j = t / n;
    const double viijj = V[i + i * n + j * n2 + j * n3];
    G2[t] = 0.5 * (viijj - V[i + j * n + j * n2 + i * n3]);
    V2[t] = viijj;
  }
}
__global__ void k_transpose_pairs(int n, int n2p, const double* __restrict__ V, double* __restrict__ Vt) {
  const size_t n2 = size_t(n) * n;
  const size_t t = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (t >= n2 * n2p) return;
  const size_t rs = t % n2p, pq = t / n2p;
  Vt[t] = rs < n2 ? V[pq + rs * n2] : 0.0;
}
// one quarter of an index transformation: out(.., p, ..) = sum_i C(i, p) in(.., i, ..) on the index of
// stride `stride` of a column-major tensor with `total` elements and extent n in every index
// (two_index_transform / four_index_transform, src/macis/transform.cxx:22-96, as n-long dot products)
__global__ void k_transform_mode(int n, size_t stride, size_t total, const double* __restrict__ Cm,
                                 const double* __restrict__ in, double* __restrict__ out) {
  const size_t o = size_t(blockIdx.x) * blockDim.x + threadIdx.x;
  if (o >= total) return;
  const size_t lo = o % stride, p = (o / stride) % n, hi = o / (stride * n);
  const double* src = in + lo + hi * stride * n;
  double acc = 0.;
  for (int i = 0; i < n; ++i) acc += Cm[i + p * size_t(n)] * src[size_t(i) * stride];
  out[o] = acc;
}
}  // namespace

// Orbital rotation of the resident integrals: T <- C^T T C, V(pqrs) <- sum C(ip) C(jq) C(kr) C(ls) V(ijkl),
// then the reduced intermediates again -- the natural-orbital step of asci_grow (asci/grow.hpp:163-215:
// two_index_transform, four_index_transform, generate_integral_intermediates). C: HOST, n x n column-major
// (columns = new orbitals). T_out / V_out (HOST, may be NULL) receive the rotated integrals.
void integrals_rotate(b2ci_ctx* ctx, const double* C, double* T_out, double* V_out) {
  if (!ctx->ints_dev) throw Error("b2ci_integrals_rotate: integrals not uploaded");
  if (!C) throw Error("b2ci_integrals_rotate: null rotation");
  const int norb = ctx->norb;
  const size_t n = norb, n2 = n * n, n4 = n2 * n2;
  cudaStream_t st = ctx->stream;
  DevBuf<double> dC(n2), a(n4), b(n4);
  B2_CUDA(cudaMemcpyAsync(dC, C, n2 * 8, cudaMemcpyHostToDevice, st));
  // T: TMP(i,q) = T(i,j) C(j,q); T'(p,q) = C(i,p) TMP(i,q)
  k_transform_mode<<<unsigned((n2 + 255) / 256), 256, 0, st>>>(norb, n, n2, dC, ctx->ints.T, a);
  k_transform_mode<<<unsigned((n2 + 255) / 256), 256, 0, st>>>(norb, 1, n2, dC, a, b);
  std::vector<double> T(n2), V(n4);
  B2_CUDA(cudaMemcpyAsync(T.data(), b, n2 * 8, cudaMemcpyDeviceToHost, st));
  // V: the four quarters in the reference's order (first index first)
  const unsigned g4 = unsigned((n4 + 255) / 256);
  k_transform_mode<<<g4, 256, 0, st>>>(norb, 1, n4, dC, ctx->ints.V, a);
  k_transform_mode<<<g4, 256, 0, st>>>(norb, n, n4, dC, a, b);
  k_transform_mode<<<g4, 256, 0, st>>>(norb, n2, n4, dC, b, a);
  k_transform_mode<<<g4, 256, 0, st>>>(norb, n2 * n, n4, dC, a, b);
  ctx->launches += 6;
  B2_CHECK_LAUNCH();
  B2_CUDA(cudaMemcpyAsync(V.data(), b, n4 * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
  void integrals_upload(b2ci_ctx*, int, const double*, const double*);
  integrals_upload(ctx, norb, T.data(), V.data());
  if (T_out) memcpy(T_out, T.data(), n2 * 8);
  if (V_out) memcpy(V_out, V.data(), n4 * 8);
}

void integrals_upload(b2ci_ctx* ctx, int norb, const double* T, const double* V) {
  if (norb < 1 || norb > 64) throw Error("b2ci_integrals_upload: norb must be in [1, 64]");
  if (!T || !V) throw Error("b2ci_integrals_upload: null integrals");
  const size_t n = norb, n2 = n * n, n3 = n2 * n, n4 = n2 * n2;
  const size_t total = ints_total_doubles(norb);
  dev_free(ctx->ints_dev);
  ctx->ints_dev = static_cast<double*>(dev_alloc(total * 8));
  double* base = ctx->ints_dev;
  cudaStream_t st = ctx->stream;
  B2_CUDA(cudaMemcpyAsync(base, T, n2 * 8, cudaMemcpyHostToDevice, st));
  B2_CUDA(cudaMemcpyAsync(base + 3 * n2 + 2 * n3, V, n4 * 8, cudaMemcpyHostToDevice, st));
  ctx->norb = norb;
  ctx->ints = make_view(norb, base);
  k_intermediates<<<unsigned((n3 + 255) / 256), 256, 0, st>>>(
      norb, ctx->ints.V, const_cast<double*>(ctx->ints.G), const_cast<double*>(ctx->ints.Vr),
      const_cast<double*>(ctx->ints.G2), const_cast<double*>(ctx->ints.V2));
  k_transpose_pairs<<<unsigned((n2 * ctx->ints.n2p + 255) / 256), 256, 0, st>>>(
      norb, ctx->ints.n2p, ctx->ints.V, const_cast<double*>(ctx->ints.Vt));
  ctx->launches += 2;
  B2_CHECK_LAUNCH();
  ctx->ints_host.resize(total);
  B2_CUDA(cudaMemcpyAsync(ctx->ints_host.data(), base, total * 8, cudaMemcpyDeviceToHost, st));
  B2_CUDA(cudaStreamSynchronize(st));
}

}  // namespace b2ci
