// Three-point feature interpolation (forward + adjoint) for sm_100a.
//
// Replace three_interpolate(_grad)_kernel of the reference
// (/root/reference/pointnet2/_ext_src/src/interpolate_gpu.cu:72-154): one CTA
// per cloud; idx and weight re-read for every channel; the adjoint issues three
// global atomicAdds per element into a zero-filled output.
//
// Forward: a thread owns FOUR consecutive target points j -- 12 indices and 12
// weights loaded once with 128-bit loads and reused for every channel of the
// CTA's channel tile -- and writes one 128-bit streaming store per channel.
//   out = ((p[i1]*w1) + (p[i2]*w2)) + (p[i3]*w3), each operation rounded (the
//   reference's expression, interpolate_gpu.cu:98-99, without FMA contraction).
// Adjoint: a CTA owns (cloud, channel tile) and accumulates g*w into a
// shared-memory image of its output rows, then writes each row once.
#include <stdlib.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kItpThreads = 256;

__device__ __forceinline__ float interp3(float p1, float p2, float p3, float w1,
                                         float w2, float w3) {
  return __fadd_rn(__fadd_rn(__fmul_rn(p1, w1), __fmul_rn(p2, w2)), __fmul_rn(p3, w3));
}

template <bool kVec>
__global__ void __launch_bounds__(kItpThreads)
interp_fwd_kernel(int c, int m, int n, int ch_per_cta, const float *__restrict__ points,
                  const int *__restrict__ idx, const float *__restrict__ weight,
                  float *__restrict__ out) {
  const int cloud = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta;
  const int c1 = min(c0 + ch_per_cta, c);
  const int *ix = idx + static_cast<size_t>(cloud) * n * 3;
  const float *wt = weight + static_cast<size_t>(cloud) * n * 3;
  const float *src = points + (static_cast<size_t>(cloud) * c + c0) * m;
  float *dst = out + (static_cast<size_t>(cloud) * c + c0) * n;
  if (kVec) {
    const int j = (blockIdx.x * kItpThreads + threadIdx.x) * 4;
    if (j >= n) return;
    const int4 ia = ld_nc_i4(ix + 3 * j), ib = ld_nc_i4(ix + 3 * j + 4), ic = ld_nc_i4(ix + 3 * j + 8);
    const float4 wa = __ldg(reinterpret_cast<const float4 *>(wt + 3 * j));
    const float4 wb = __ldg(reinterpret_cast<const float4 *>(wt + 3 * j + 4));
    const float4 wc = __ldg(reinterpret_cast<const float4 *>(wt + 3 * j + 8));
#pragma unroll 2
    for (int l = c0; l < c1; ++l, src += m, dst += n) {
      float4 v;
      v.x = interp3(__ldg(src + ia.x), __ldg(src + ia.y), __ldg(src + ia.z), wa.x, wa.y, wa.z);
      v.y = interp3(__ldg(src + ia.w), __ldg(src + ib.x), __ldg(src + ib.y), wa.w, wb.x, wb.y);
      v.z = interp3(__ldg(src + ib.z), __ldg(src + ib.w), __ldg(src + ic.x), wb.z, wb.w, wc.x);
      v.w = interp3(__ldg(src + ic.y), __ldg(src + ic.z), __ldg(src + ic.w), wc.y, wc.z, wc.w);
      st_cs_f4(dst + j, v);
    }
  } else {
    const int j = blockIdx.x * kItpThreads + threadIdx.x;
    if (j >= n) return;
    const int i1 = __ldg(ix + 3 * j), i2 = __ldg(ix + 3 * j + 1), i3 = __ldg(ix + 3 * j + 2);
    const float w1 = __ldg(wt + 3 * j), w2 = __ldg(wt + 3 * j + 1), w3 = __ldg(wt + 3 * j + 2);
    for (int l = c0; l < c1; ++l, src += m, dst += n)
      dst[j] = interp3(__ldg(src + i1), __ldg(src + i2), __ldg(src + i3), w1, w2, w3);
  }
}

// ---- Shared-memory gather (the default where the shape suits it; PN2_INTERP_SMEM=0 selects the kernel
// above).  The kernel above spends its time in L1 gathers (12 scattered loads per 4 outputs and channel).
// Here a CTA owns (cloud, channel tile), stages its [channels][m] rows of known features in shared memory
// once -- they are contiguous, each element leaves HBM exactly once -- and gathers from there.  Same
// thread mapping, same arithmetic (interp3), same stores; bit-identical output
// (tests/test_gather_variants_gpu.py).  Measured (B200, B = 32, profiles/r02_gather_smem_vs_l1.txt):
// FP0 128 x 1024 -> 8192: 112.5 -> 72.9 us (21 -> 33 % of the HBM peak).
constexpr int kIsThreads = 512;
constexpr int kIsSmemBytes = 56 * 1024;  // four CTAs per SM

template <bool kVec>
__global__ void __launch_bounds__(kIsThreads)
interp_fwd_smem_kernel(int c, int m, int n, int ch_per_cta, const float *__restrict__ points,
                       const int *__restrict__ idx, const float *__restrict__ weight,
                       float *__restrict__ out) {
  extern __shared__ __align__(16) float s_src[];  // [channels of this CTA][m]
  const int cloud = blockIdx.y, tid = threadIdx.x;
  const int c0 = blockIdx.x * ch_per_cta;
  const int nc = min(ch_per_cta, c - c0);
  const float *src = points + (static_cast<size_t>(cloud) * c + c0) * m;
  const int total = nc * m;
  if ((m & 3) == 0 && (reinterpret_cast<uintptr_t>(points) & 15u) == 0) {
    for (int i = tid * 4; i < total; i += kIsThreads * 4)
      *reinterpret_cast<float4 *>(s_src + i) = __ldg(reinterpret_cast<const float4 *>(src + i));
  } else {
    for (int i = tid; i < total; i += kIsThreads) s_src[i] = __ldg(src + i);
  }
  __syncthreads();
  const int *ix = idx + static_cast<size_t>(cloud) * n * 3;
  const float *wt = weight + static_cast<size_t>(cloud) * n * 3;
  float *dst = out + (static_cast<size_t>(cloud) * c + c0) * n;
  if (kVec) {
    for (int j = tid * 4; j < n; j += kIsThreads * 4) {
      const int4 ia = ld_nc_i4(ix + 3 * j), ib = ld_nc_i4(ix + 3 * j + 4), ic = ld_nc_i4(ix + 3 * j + 8);
      const float4 wa = __ldg(reinterpret_cast<const float4 *>(wt + 3 * j));
      const float4 wb = __ldg(reinterpret_cast<const float4 *>(wt + 3 * j + 4));
      const float4 wc = __ldg(reinterpret_cast<const float4 *>(wt + 3 * j + 8));
#pragma unroll 2
      for (int l = 0; l < nc; ++l) {
        const float *row = s_src + l * m;
        float4 v;
        v.x = interp3(row[ia.x], row[ia.y], row[ia.z], wa.x, wa.y, wa.z);
        v.y = interp3(row[ia.w], row[ib.x], row[ib.y], wa.w, wb.x, wb.y);
        v.z = interp3(row[ib.z], row[ib.w], row[ic.x], wb.z, wb.w, wc.x);
        v.w = interp3(row[ic.y], row[ic.z], row[ic.w], wc.y, wc.z, wc.w);
        st_cs_f4(dst + static_cast<size_t>(l) * n + j, v);
      }
    }
  } else {
    for (int j = tid; j < n; j += kIsThreads) {
      const int i1 = __ldg(ix + 3 * j), i2 = __ldg(ix + 3 * j + 1), i3 = __ldg(ix + 3 * j + 2);
      const float w1 = __ldg(wt + 3 * j), w2 = __ldg(wt + 3 * j + 1), w3 = __ldg(wt + 3 * j + 2);
      for (int l = 0; l < nc; ++l) {
        const float *row = s_src + l * m;
        dst[static_cast<size_t>(l) * n + j] = interp3(row[i1], row[i2], row[i3], w1, w2, w3);
      }
    }
  }
}

static bool interp_smem_wanted(int m, int n) {
  const char *env = getenv("PN2_INTERP_SMEM");
  return !(env != nullptr && env[0] == '0') && m > 0 && static_cast<size_t>(m) * sizeof(float) <= kIsSmemBytes &&
         3LL * n >= 2LL * m;
}

constexpr int kItpGradThreads = 512;
constexpr int kItpGradSmemFloats = 48 * 1024;

__global__ void __launch_bounds__(kItpGradThreads, 1)
interp_grad_smem_kernel(int c, int n, int m, int ch_per_cta,
                        const float *__restrict__ grad_out, const int *__restrict__ idx,
                        const float *__restrict__ weight, float *__restrict__ grad_points) {
  extern __shared__ float s_acc[];  // [channels of this CTA][m]
  const int cloud = blockIdx.y;
  const int c0 = blockIdx.x * ch_per_cta;
  const int nc = min(ch_per_cta, c - c0);
  const int tid = threadIdx.x;
  const int *ix = idx + static_cast<size_t>(cloud) * n * 3;
  const float *wt = weight + static_cast<size_t>(cloud) * n * 3;
  const float *g = grad_out + (static_cast<size_t>(cloud) * c + c0) * n;

  for (int i = tid; i < nc * m; i += kItpGradThreads) s_acc[i] = 0.f;
  __syncthreads();
  for (int j = tid; j < n; j += kItpGradThreads) {
    const int i1 = __ldg(ix + 3 * j), i2 = __ldg(ix + 3 * j + 1), i3 = __ldg(ix + 3 * j + 2);
    const float w1 = __ldg(wt + 3 * j), w2 = __ldg(wt + 3 * j + 1), w3 = __ldg(wt + 3 * j + 2);
    for (int l = 0; l < nc; ++l) {
      const float v = __ldg(g + static_cast<size_t>(l) * n + j);
      float *row = s_acc + l * m;
      atomicAdd(row + i1, __fmul_rn(v, w1));
      atomicAdd(row + i2, __fmul_rn(v, w2));
      atomicAdd(row + i3, __fmul_rn(v, w3));
    }
  }
  __syncthreads();
  float *dst = grad_points + (static_cast<size_t>(cloud) * c + c0) * m;
  for (int i = tid; i < nc * m; i += kItpGradThreads) dst[i] = s_acc[i];
}

__global__ void __launch_bounds__(kItpThreads)
interp_grad_global_kernel(int c, int n, int m, const float *__restrict__ grad_out,
                          const int *__restrict__ idx, const float *__restrict__ weight,
                          float *__restrict__ grad_points) {
  const int cloud = blockIdx.z, l = blockIdx.y;
  const int j = blockIdx.x * kItpThreads + threadIdx.x;
  if (j >= n) return;
  const int *ix = idx + (static_cast<size_t>(cloud) * n + j) * 3;
  const float *wt = weight + (static_cast<size_t>(cloud) * n + j) * 3;
  const float v = __ldg(grad_out + (static_cast<size_t>(cloud) * c + l) * n + j);
  float *row = grad_points + (static_cast<size_t>(cloud) * c + l) * m;
  atomicAdd(row + __ldg(ix + 0), __fmul_rn(v, __ldg(wt + 0)));
  atomicAdd(row + __ldg(ix + 1), __fmul_rn(v, __ldg(wt + 1)));
  atomicAdd(row + __ldg(ix + 2), __fmul_rn(v, __ldg(wt + 2)));
}

}  // namespace pn2

extern "C" int pn2_three_interpolate(int b, int c, int m, int n, const float *points,
                                     const int *idx, const float *weight, float *out,
                                     void *stream) {
  using namespace pn2;
  if (b < 0 || c < 0 || m < 0 || n < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || n == 0) return PN2_OK;
  PN2_CHECK_PTRS(points, idx, weight, out);
  if (b > 65535) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const bool vec = (n % 4 == 0) && aligned16(idx) && aligned16(weight) && aligned16(out);
  if (interp_smem_wanted(m, n)) {
    const int ch_max = kIsSmemBytes / (static_cast<int>(sizeof(float)) * m);
    int tiles = (c + ch_max - 1) / ch_max;
    const int want = (2 * kNumSMs + b - 1) / b;  // two waves of CTAs when the channel count allows it
    if (tiles < want) tiles = want < c ? want : c;
    const int per_cta = (c + tiles - 1) / tiles;
    tiles = (c + per_cta - 1) / per_cta;
    const size_t smem = sizeof(float) * static_cast<size_t>(per_cta) * m;
    auto kern = vec ? interp_fwd_smem_kernel<true> : interp_fwd_smem_kernel<false>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kIsSmemBytes);
    if (e != cudaSuccess) return static_cast<int>(e);
    kern<<<dim3(tiles, b), kIsThreads, smem, s>>>(c, m, n, per_cta, points, idx, weight, out);
    count_launch();
    PN2_RETURN_LAUNCH_STATUS();
  }
  const int per_thread = vec ? 4 : 1;
  const int xblocks = (n + kItpThreads * per_thread - 1) / (kItpThreads * per_thread);
  int ctiles = (4 * kNumSMs + b * xblocks - 1) / (b * xblocks);
  if (ctiles > (c + 3) / 4) ctiles = (c + 3) / 4;
  if (ctiles < 1) ctiles = 1;
  if (ctiles > 65535) ctiles = 65535;
  const int ch_per_cta = (c + ctiles - 1) / ctiles;
  ctiles = (c + ch_per_cta - 1) / ch_per_cta;
  dim3 grid(xblocks, ctiles, b);
  if (vec)
    interp_fwd_kernel<true><<<grid, kItpThreads, 0, s>>>(c, m, n, ch_per_cta, points, idx, weight, out);
  else
    interp_fwd_kernel<false><<<grid, kItpThreads, 0, s>>>(c, m, n, ch_per_cta, points, idx, weight, out);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_three_interpolate_grad(int b, int c, int n, int m,
                                          const float *grad_out, const int *idx,
                                          const float *weight, float *grad_points,
                                          void *stream) {
  using namespace pn2;
  if (b < 0 || c < 0 || m < 0 || n < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || m == 0) return PN2_OK;
  PN2_CHECK_PTRS(grad_points);
  if (n > 0) PN2_CHECK_PTRS(grad_out, idx, weight);
  if (b > 65535 || c > 65535) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (m <= kItpGradSmemFloats) {
    const int ch_max = kItpGradSmemFloats / m;
    int ctiles = (c + ch_max - 1) / ch_max;
    const int want = (2 * kNumSMs + b - 1) / b;
    if (ctiles < want) ctiles = want < c ? want : c;
    const int ch_per_cta = (c + ctiles - 1) / ctiles;
    ctiles = (c + ch_per_cta - 1) / ch_per_cta;
    const size_t smem = sizeof(float) * static_cast<size_t>(ch_per_cta) * m;
    cudaError_t e = cudaFuncSetAttribute(interp_grad_smem_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    interp_grad_smem_kernel<<<dim3(ctiles, b), kItpGradThreads, smem, s>>>(
        c, n, m, ch_per_cta, grad_out, idx, weight, grad_points);
    count_launch();
    PN2_RETURN_LAUNCH_STATUS();
  }
  cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * static_cast<size_t>(b) * c * m, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (n == 0) return PN2_OK;
  interp_grad_global_kernel<<<dim3((n + kItpThreads - 1) / kItpThreads, c, b), kItpThreads, 0, s>>>(
      c, n, m, grad_out, idx, weight, grad_points);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
