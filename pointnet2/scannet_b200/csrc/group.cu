// Gather / grouping (forward) and their scatter-add adjoints for sm_100a.
//
// Replace gather_points(_grad)_kernel (sampling_gpu.cu:8-57) and
// group_points(_grad)_kernel (group_points_gpu.cu:8-75) of the reference.
//
// Forward (reference: thread <-> (channel, centre) with an inner loop over
// nsample, so stores are strided by nsample*4 bytes across a warp and idx is
// re-read once per channel):  here a thread owns FOUR consecutive output
// positions e = (centre, sample) -- one 128-bit index load, reused for every
// channel of the CTA's channel tile -- and writes one 128-bit streaming store
// per channel, so a warp stores 512 contiguous bytes.
//
// Backward (reference: one global atomicAdd per element into a zero-filled
// output):  here a CTA owns (cloud, channel tile), accumulates into a
// shared-memory image of its output rows with shared-memory atomics, and
// writes every row exactly once, coalesced (the output needs no zero fill).
// Ball-query rows repeat their first hit in every unused slot, so most of a
// warp's adds hit ONE address: those lanes are summed with a shuffle butterfly
// first (warp-aggregated scatter-add) and only the row's head lane issues the
// atomic.  Outputs too large for shared memory fall back to memset + global
// reductions.
#include <stdlib.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kGrpThreads = 256;

// ------------------------------------------------------------------ forward
// out[b, c, e] = points[b, c, idx[b, e]],  e in [0, E)
template <bool kVec>
__global__ void __launch_bounds__(kGrpThreads)
group_fwd_kernel(int c, int n, int E, int ch_per_cta, const float *__restrict__ points,
                 const int *__restrict__ idx, float *__restrict__ out) {
  const int cloud = blockIdx.z;
  const int c0 = blockIdx.y * ch_per_cta;
  const int c1 = min(c0 + ch_per_cta, c);
  const int *ix = idx + static_cast<size_t>(cloud) * E;
  const float *src = points + (static_cast<size_t>(cloud) * c + c0) * n;
  float *dst = out + (static_cast<size_t>(cloud) * c + c0) * E;
  if (kVec) {
    const int e = (blockIdx.x * kGrpThreads + threadIdx.x) * 4;
    if (e >= E) return;
    const int4 i = ld_nc_i4(ix + e);
#pragma unroll 4
    for (int l = c0; l < c1; ++l, src += n, dst += E) {
      float4 v;
      v.x = __ldg(src + i.x), v.y = __ldg(src + i.y);
      v.z = __ldg(src + i.z), v.w = __ldg(src + i.w);
      st_cs_f4(dst + e, v);
    }
  } else {
    const int e = blockIdx.x * kGrpThreads + threadIdx.x;
    if (e >= E) return;
    const int i = __ldg(ix + e);
    for (int l = c0; l < c1; ++l, src += n, dst += E) dst[e] = __ldg(src + i);
  }
}

// ---- Shared-memory gather (the default where the shape suits it; PN2_GROUP_SMEM=0 selects the kernel above).
// The kernel above is bound by L1 gathers (a warp's 32 indices touch up to 32 sectors per load
// instruction).  Here a CTA owns (cloud, channel tile), copies its [channels][n] rows -- contiguous in
// memory, each element read from HBM exactly once -- into shared memory with 128-bit loads, and gathers
// from there: a 32-way scattered LDS costs a few bank-conflict replays instead of 32 sector lookups.
// Same thread mapping and stores as above; bit-identical output (tests/test_gather_variants_gpu.py).
// Measured (B200, B = 32, profiles/r02_gather_smem_vs_l1.txt): SA2 features 64 x 1024 -> 256 x 32:
// 31.9 -> 21.5 us (37 -> 54 % of the HBM peak); multiview SA1 131 x 8192 -> 1024 x 32: 476 -> 155 us
// (22 -> 68 %).
constexpr int kGsThreads = 512;
constexpr int kGsSmemBytes = 56 * 1024;  // four CTAs per SM

template <bool kVec>
__global__ void __launch_bounds__(kGsThreads)
group_fwd_smem_kernel(int c, int n, int E, int ch_per_cta, const float *__restrict__ points,
                      const int *__restrict__ idx, float *__restrict__ out) {
  extern __shared__ __align__(16) float s_src[];  // [channels of this CTA][n]
  const int cloud = blockIdx.y, tid = threadIdx.x;
  const int c0 = blockIdx.x * ch_per_cta;
  const int nc = min(ch_per_cta, c - c0);
  const float *src = points + (static_cast<size_t>(cloud) * c + c0) * n;
  const int total = nc * n;
  if ((n & 3) == 0 && (reinterpret_cast<uintptr_t>(points) & 15u) == 0) {
    for (int i = tid * 4; i < total; i += kGsThreads * 4)
      *reinterpret_cast<float4 *>(s_src + i) = __ldg(reinterpret_cast<const float4 *>(src + i));
  } else {
    for (int i = tid; i < total; i += kGsThreads) s_src[i] = __ldg(src + i);
  }
  __syncthreads();
  const int *ix = idx + static_cast<size_t>(cloud) * E;
  float *dst = out + (static_cast<size_t>(cloud) * c + c0) * E;
  if (kVec) {
    for (int e = tid * 4; e < E; e += kGsThreads * 4) {
      const int4 i = ld_nc_i4(ix + e);
#pragma unroll 4
      for (int l = 0; l < nc; ++l) {
        const float *row = s_src + l * n;
        st_cs_f4(dst + static_cast<size_t>(l) * E + e, make_float4(row[i.x], row[i.y], row[i.z], row[i.w]));
      }
    }
  } else {
    for (int e = tid; e < E; e += kGsThreads) {
      const int i = __ldg(ix + e);
      for (int l = 0; l < nc; ++l) dst[static_cast<size_t>(l) * E + e] = s_src[l * n + i];
    }
  }
}

// true unless switched off, when the shape suits it (rows fit, enough reuse)
static bool group_smem_wanted(int n, int E) {
  const char *env = getenv("PN2_GROUP_SMEM");
  return !(env != nullptr && env[0] == '0') && n > 0 && static_cast<size_t>(n) * sizeof(float) <= kGsSmemBytes &&
         static_cast<long long>(E) >= 2LL * n;
}

static int launch_group_fwd_smem(int b, int c, int n, int E, bool vec, const float *points, const int *idx,
                                 float *out, cudaStream_t s) {
  const int ch_max = kGsSmemBytes / (static_cast<int>(sizeof(float)) * n);
  int ctiles = (c + ch_max - 1) / ch_max;
  const int want = (2 * kNumSMs + b - 1) / b;  // two waves of CTAs when the channel count allows it
  if (ctiles < want) ctiles = want < c ? want : c;
  const int ch_per_cta = (c + ctiles - 1) / ctiles;
  ctiles = (c + ch_per_cta - 1) / ch_per_cta;
  const size_t smem = sizeof(float) * static_cast<size_t>(ch_per_cta) * n;
  auto kern = vec ? group_fwd_smem_kernel<true> : group_fwd_smem_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, kGsSmemBytes);
  if (e != cudaSuccess) return static_cast<int>(e);
  kern<<<dim3(ctiles, b), kGsThreads, smem, s>>>(c, n, E, ch_per_cta, points, idx, out);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

static int launch_group_fwd(int b, int c, int n, int E, const float *points,
                            const int *idx, float *out, cudaStream_t s) {
  if (b == 0 || c == 0 || E == 0) return PN2_OK;
  if (b > 65535) return PN2_ERR_UNSUPPORTED;
  const bool vec = (E % 4 == 0) && aligned16(idx) && aligned16(out);
  if (group_smem_wanted(n, E)) return launch_group_fwd_smem(b, c, n, E, vec, points, idx, out, s);
  const int per_thread = vec ? 4 : 1;
  const int xblocks = (E + kGrpThreads * per_thread - 1) / (kGrpThreads * per_thread);
  // enough CTAs for ~4 waves of 148 SMs, but keep >= 4 channels per idx load
  int ctiles = (4 * kNumSMs + b * xblocks - 1) / (b * xblocks);
  if (ctiles > (c + 3) / 4) ctiles = (c + 3) / 4;
  if (ctiles < 1) ctiles = 1;
  if (ctiles > 65535) ctiles = 65535;
  const int ch_per_cta = (c + ctiles - 1) / ctiles;
  ctiles = (c + ch_per_cta - 1) / ch_per_cta;
  dim3 grid(xblocks, ctiles, b);
  if (vec)
    group_fwd_kernel<true><<<grid, kGrpThreads, 0, s>>>(c, n, E, ch_per_cta, points, idx, out);
  else
    group_fwd_kernel<false><<<grid, kGrpThreads, 0, s>>>(c, n, E, ch_per_cta, points, idx, out);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

// ----------------------------------------------------------------- backward
// grad_points[b, c, idx[b, e]] += grad_out[b, c, e]
constexpr int kGradThreads = 512;
constexpr int kGradSmemFloats = 48 * 1024;  // 192 KB accumulator image per CTA

__global__ void __launch_bounds__(kGradThreads, 1)
group_grad_smem_kernel(int c, int n, int E, int seg, int ch_per_cta,
                       const float *__restrict__ grad_out, const int *__restrict__ idx,
                       float *__restrict__ grad_points) {
  extern __shared__ float s_acc[];  // [channels of this CTA][n]
  const int cloud = blockIdx.y;
  const int c0 = blockIdx.x * ch_per_cta;
  const int nc = min(ch_per_cta, c - c0);
  const int tid = threadIdx.x, lane = tid & 31;
  const int *ix = idx + static_cast<size_t>(cloud) * E;
  const float *g = grad_out + (static_cast<size_t>(cloud) * c + c0) * E;

  for (int i = tid; i < nc * n; i += kGradThreads) s_acc[i] = 0.f;
  __syncthreads();

  const int head = lane & ~(seg - 1);
  const int Eround = (E + 31) & ~31;
  for (int e = tid; e < Eround; e += kGradThreads) {
    const bool in = e < E;
    const int i = in ? __ldg(ix + e) : -1;
    // lanes of this row segment that address the same element as its head lane
    const int ihead = __shfl_sync(0xffffffffu, i, head);
    const bool dup = in && i == ihead;
    const bool aggregate = __ballot_sync(0xffffffffu, dup && lane != head) != 0;
    if (aggregate) {
      for (int l = 0; l < nc; ++l) {
        const float v = in ? __ldg(g + static_cast<size_t>(l) * E + e) : 0.f;
        float sum = dup ? v : 0.f;
        for (int o = 1; o < seg; o <<= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (in && (lane == head || !dup)) atomicAdd(&s_acc[l * n + i], lane == head ? sum : v);
      }
    } else if (in) {
      for (int l = 0; l < nc; ++l)
        atomicAdd(&s_acc[l * n + i], __ldg(g + static_cast<size_t>(l) * E + e));
    }
  }
  __syncthreads();
  float *dst = grad_points + (static_cast<size_t>(cloud) * c + c0) * n;
  for (int i = tid; i < nc * n; i += kGradThreads) dst[i] = s_acc[i];
}

__global__ void __launch_bounds__(kGrpThreads)
group_grad_global_kernel(int c, int n, int E, const float *__restrict__ grad_out,
                         const int *__restrict__ idx, float *__restrict__ grad_points) {
  const int cloud = blockIdx.z, l = blockIdx.y;
  const int e = blockIdx.x * kGrpThreads + threadIdx.x;
  if (e >= E) return;
  const int i = __ldg(idx + static_cast<size_t>(cloud) * E + e);
  atomicAdd(grad_points + (static_cast<size_t>(cloud) * c + l) * n + i,
            __ldg(grad_out + (static_cast<size_t>(cloud) * c + l) * E + e));
}

static int launch_group_grad(int b, int c, int n, int E, int nsample,
                             const float *grad_out, const int *idx,
                             float *grad_points, cudaStream_t s) {
  if (b == 0 || c == 0 || n == 0) return PN2_OK;
  if (b > 65535 || c > 65535) return PN2_ERR_UNSUPPORTED;
  if (n <= kGradSmemFloats) {
    const int ch_max = kGradSmemFloats / n;
    // at least one CTA per SM when the channel count allows it
    int ctiles = (c + ch_max - 1) / ch_max;
    const int want = (kNumSMs + b - 1) / b;
    if (ctiles < want) ctiles = want < c ? want : c;
    const int ch_per_cta = (c + ctiles - 1) / ctiles;
    ctiles = (c + ch_per_cta - 1) / ch_per_cta;
    // row segment for the duplicate aggregation: one ball-query row per
    // segment when nsample is a power of two <= 32, else the whole warp
    int seg = 32;
    if (nsample >= 1 && nsample <= 32 && (nsample & (nsample - 1)) == 0) seg = nsample;
    const size_t smem = sizeof(float) * static_cast<size_t>(ch_per_cta) * n;
    cudaError_t e = cudaFuncSetAttribute(group_grad_smem_kernel,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return static_cast<int>(e);
    group_grad_smem_kernel<<<dim3(ctiles, b), kGradThreads, smem, s>>>(
        c, n, E, seg, ch_per_cta, grad_out, idx, grad_points);
    count_launch();
    PN2_RETURN_LAUNCH_STATUS();
  }
  cudaError_t e = cudaMemsetAsync(grad_points, 0, sizeof(float) * static_cast<size_t>(b) * c * n, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  if (E == 0) return PN2_OK;
  group_grad_global_kernel<<<dim3((E + kGrpThreads - 1) / kGrpThreads, c, b), kGrpThreads, 0, s>>>(
      c, n, E, grad_out, idx, grad_points);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

}  // namespace pn2

// gather == grouping with nsample = 1 (idx (b, npoints))
extern "C" int pn2_gather_points(int b, int c, int n, int npoints,
                                 const float *points, const int *idx, float *out,
                                 void *stream) {
  if (b < 0 || c < 0 || n < 0 || npoints < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || npoints == 0) return PN2_OK;
  PN2_CHECK_PTRS(points, idx, out);
  return pn2::launch_group_fwd(b, c, n, npoints, points, idx, out,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int pn2_gather_points_grad(int b, int c, int n, int npoints,
                                      const float *grad_out, const int *idx,
                                      float *grad_points, void *stream) {
  if (b < 0 || c < 0 || n < 0 || npoints < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || n == 0) return PN2_OK;
  PN2_CHECK_PTRS(grad_points);
  if (npoints > 0) PN2_CHECK_PTRS(grad_out, idx);
  return pn2::launch_group_grad(b, c, n, npoints, 1, grad_out, idx, grad_points,
                                static_cast<cudaStream_t>(stream));
}

extern "C" int pn2_group_points(int b, int c, int n, int npoints, int nsample,
                                const float *points, const int *idx, float *out,
                                void *stream) {
  if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || npoints == 0 || nsample == 0) return PN2_OK;
  PN2_CHECK_PTRS(points, idx, out);
  if (static_cast<long long>(npoints) * nsample > 0x7fffffffLL) return PN2_ERR_UNSUPPORTED;
  return pn2::launch_group_fwd(b, c, n, npoints * nsample, points, idx, out,
                               static_cast<cudaStream_t>(stream));
}

extern "C" int pn2_group_points_grad(int b, int c, int n, int npoints, int nsample,
                                     const float *grad_out, const int *idx,
                                     float *grad_points, void *stream) {
  if (b < 0 || c < 0 || n < 0 || npoints < 0 || nsample < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || c == 0 || n == 0) return PN2_OK;
  PN2_CHECK_PTRS(grad_points);
  const long long E = static_cast<long long>(npoints) * nsample;
  if (E > 0x7fffffffLL) return PN2_ERR_UNSUPPORTED;
  if (E > 0) PN2_CHECK_PTRS(grad_out, idx);
  return pn2::launch_group_grad(b, c, n, static_cast<int>(E), nsample, grad_out, idx,
                                grad_points, static_cast<cudaStream_t>(stream));
}
