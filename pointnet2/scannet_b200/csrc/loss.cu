// Weighted cross-entropy of the semseg training step for sm_100a
// (/root/reference/lib/loss.py:10-18: F.cross_entropy(reduction="none", ignore_index) -> mean(loss * weights)),
// on the classifier's [P, C] logit rows (P = B * N points, C = num_classes).
//
// torch runs log_softmax, nll_loss, a multiply and a mean forward (4 kernels, the [P, C] log-probabilities
// written and re-read) and softmax_backward + nll_backward behind them.  Here the forward pass reads the
// logits once and keeps one float per row (the row's log-sum-exp); the adjoint re-reads the logits and
// writes d logits in one pass:
//
//   loss   = 1/P * sum_i w_i * (lse_i - x[i, t_i])                     (rows with t_i == ignore_index: 0)
//   dx[i,c] = g/P * w_i * (exp(x[i,c] - lse_i) - [c == t_i])            (g = upstream gradient, on the device)
//
// The mean is order-deterministic: warp shuffles, one fp32 partial per CTA, and the CTA that finishes
// last adds the partials in index order in fp64.
#include "pn2_common.cuh"

namespace pn2 {

constexpr int kCeThreads = 256;

__global__ void __launch_bounds__(kCeThreads)
weighted_ce_fwd_kernel(int P, int C4, long long ignore_index, const float *__restrict__ logits,
                       const long long *__restrict__ targets, const float *__restrict__ weights,
                       float *__restrict__ lse_out, float *__restrict__ partials,
                       unsigned *__restrict__ counter, float *__restrict__ loss) {
  pdl_enter();
  __shared__ float s_warp[kCeThreads / 32];
  __shared__ bool s_last;
  const int i = blockIdx.x * kCeThreads + threadIdx.x;
  float term = 0.f;
  if (i < P) {
    const float4 *row = reinterpret_cast<const float4 *>(logits) + static_cast<size_t>(i) * C4;
    float mx = -INFINITY;
    for (int c = 0; c < C4; ++c) {
      const float4 v = __ldg(row + c);
      mx = fmaxf(fmaxf(fmaxf(mx, v.x), fmaxf(v.y, v.z)), v.w);
    }
    float s = 0.f;
    for (int c = 0; c < C4; ++c) {  // second pass over the row hits L1
      const float4 v = __ldg(row + c);
      s += expf(v.x - mx) + expf(v.y - mx) + expf(v.z - mx) + expf(v.w - mx);
    }
    const float lse = mx + logf(s);
    lse_out[i] = lse;
    const long long t = targets[i];
    if (t != ignore_index && t >= 0 && t < 4ll * C4) {
      const float w = weights != nullptr ? __ldg(weights + i) : 1.f;
      term = w * (lse - __ldg(logits + static_cast<size_t>(i) * C4 * 4 + t));
    }
  }
  for (int o = 16; o > 0; o >>= 1) term += __shfl_xor_sync(0xffffffffu, term, o);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = term;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < kCeThreads / 32; ++w) t += s_warp[w];
    partials[blockIdx.x] = t;
    __threadfence();
    s_last = atomicAdd(counter, 1u) == gridDim.x - 1;
  }
  __syncthreads();
  if (!s_last || threadIdx.x >= 32) return;
  __threadfence();
  double acc = 0.0;  // lane l adds partials l, l+32, ...; fixed shuffle tree
  for (unsigned k = threadIdx.x; k < gridDim.x; k += 32) acc += static_cast<double>(__ldcg(partials + k));
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (threadIdx.x == 0) {
    *loss = static_cast<float>(acc / static_cast<double>(P));
    *counter = 0u;  // re-armed for the next launch (kernels of one stream run one after another)
  }
}

__global__ void __launch_bounds__(kCeThreads)
weighted_ce_bwd_kernel(size_t total4, int C4, int P, long long ignore_index, const float *__restrict__ logits,
                       const long long *__restrict__ targets, const float *__restrict__ weights,
                       const float *__restrict__ lse, const float *__restrict__ grad_loss,
                       float *__restrict__ dlogits) {
  pdl_enter();
  const size_t e = static_cast<size_t>(blockIdx.x) * kCeThreads + threadIdx.x;
  if (e >= total4) return;
  const unsigned i = static_cast<unsigned>(e / C4);
  const int c = static_cast<int>(e - static_cast<size_t>(i) * C4) * 4;
  const long long t = targets[i];
  float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
  if (t != ignore_index && t >= 0 && t < 4ll * C4) {
    const float g = __ldg(grad_loss) / static_cast<float>(P) * (weights != nullptr ? __ldg(weights + i) : 1.f);
    const float4 v = __ldg(reinterpret_cast<const float4 *>(logits) + e);
    const float l = __ldg(lse + i);
    const int tc = static_cast<int>(t) - c;
    o.x = g * (expf(v.x - l) - (tc == 0 ? 1.f : 0.f));
    o.y = g * (expf(v.y - l) - (tc == 1 ? 1.f : 0.f));
    o.z = g * (expf(v.z - l) - (tc == 2 ? 1.f : 0.f));
    o.w = g * (expf(v.w - l) - (tc == 3 ? 1.f : 0.f));
  }
  reinterpret_cast<float4 *>(dlogits)[e] = o;
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_weighted_ce_partials(int p) { return (p + kCeThreads - 1) / kCeThreads; }

extern "C" int pn2_weighted_ce_fwd(int p, int c, long long ignore_index, const float *logits,
                                   const long long *targets, const float *weights, float *lse,
                                   float *partials, unsigned *counter, float *loss, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(logits, targets, lse, partials, counter, loss);
  if (c % 4 != 0 || c > 4096 || !aligned16(logits)) return PN2_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(targets) & 7u) != 0) return PN2_ERR_ALIGNMENT;
  pn2::launch(weighted_ce_fwd_kernel, pn2_weighted_ce_partials(p), kCeThreads, 0, static_cast<cudaStream_t>(stream), p, c / 4, ignore_index, logits, targets, weights, lse, partials, counter, loss);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_weighted_ce_bwd(int p, int c, long long ignore_index, const float *logits,
                                   const long long *targets, const float *weights, const float *lse,
                                   const float *grad_loss, float *dlogits, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(logits, targets, lse, grad_loss, dlogits);
  if (c % 4 != 0 || c > 4096 || !aligned16(logits) || !aligned16(dlogits)) return PN2_ERR_UNSUPPORTED;
  if ((reinterpret_cast<uintptr_t>(targets) & 7u) != 0) return PN2_ERR_ALIGNMENT;
  const size_t total4 = static_cast<size_t>(p) * (c / 4);
  const unsigned blocks = static_cast<unsigned>((total4 + kCeThreads - 1) / kCeThreads);
  pn2::launch(weighted_ce_bwd_kernel, blocks, kCeThreads, 0, static_cast<cudaStream_t>(stream), total4, c / 4, p, ignore_index, logits, targets, weights, lse, grad_loss, dlogits);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
