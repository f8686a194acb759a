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
}  // namespace

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
