// Three nearest neighbours for sm_100a.
//
// Replaces three_nn_kernel / _wrapper of the reference
// (/root/reference/pointnet2/_ext_src/src/interpolate_gpu.cu:9-68): one CTA per
// cloud, every thread streaming all m known points from global memory with
// fp64 compares (running bests are `double`, :27).
//
// Here the grid covers (query tile, cloud); the known points are staged in
// shared memory tile by tile, re-laid out as pairs [x0 x1 y0 y1 z0 z1]; every lane
// owns one query and reads the same pair (broadcast LDS.64), evaluating both
// distances with packed fp32 instructions.
// The compares are fp32: a float candidate compared with a double that holds
// either 1e40 or an earlier float candidate orders exactly like the same
// compare in fp32 against +inf -- and (float)1e40, which the reference stores
// when m < 3, is +inf as well.
#include <math.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kNnThreads = 256;
constexpr int kNnTilePts = 4096;  // at most 48 KB of known xyz per tile

__global__ void __launch_bounds__(kNnThreads)
three_nn_kernel(int n, int m, const float *__restrict__ unknown,
                const float *__restrict__ known, float *__restrict__ dist2,
                int *__restrict__ idx) {
  extern __shared__ __align__(16) float s_known[];  // [ceil(min(m, kNnTilePts) / 2)][6]

  const int tid = threadIdx.x;
  const int cloud = blockIdx.y;
  const int j = blockIdx.x * kNnThreads + tid;
  const float *kn = known + static_cast<size_t>(cloud) * m * 3;
  const float *un = unknown + static_cast<size_t>(cloud) * n * 3;

  float ux = 0.f, uy = 0.f, uz = 0.f;
  if (j < n) ux = __ldg(un + 3 * j), uy = __ldg(un + 3 * j + 1), uz = __ldg(un + 3 * j + 2);

  float b1 = INFINITY, b2 = INFINITY, b3 = INFINITY;
  int i1 = 0, i2 = 0, i3 = 0;

  // known points are staged as PAIRS in structure-of-arrays form, [x0 x1 y0 y1 z0 z1] per pair,
  // so that three LDS.64 deliver the register pairs the packed distance evaluation consumes
  for (int t0 = 0; t0 < m; t0 += kNnTilePts) {
    const int tn = min(kNnTilePts, m - t0);
    const float *src = kn + static_cast<size_t>(t0) * 3;
    if (t0 > 0) __syncthreads();  // previous tile fully consumed
    for (int i = tid; i < tn * 3; i += kNnThreads) {
      const int k = i / 3, c = i - 3 * k;
      s_known[6 * (k >> 1) + 2 * c + (k & 1)] = __ldg(src + i);
    }
    if (tn & 1) {  // odd tail: duplicate the last point into the unused half (re-insertion of an
      if (tid < 3) s_known[6 * (tn >> 1) + 2 * tid + 1] = __ldg(src + 3 * (tn - 1) + tid);  // equal
    }                                                                 // (d, later index) is a no-op)
    __syncthreads();
    const float2 u2x = make_float2(ux, ux), u2y = make_float2(uy, uy), u2z = make_float2(uz, uz);
    const int pairs = (tn + 1) >> 1;
#pragma unroll 4
    for (int h = 0; h < pairs; ++h) {
      const float2 kx = *reinterpret_cast<const float2 *>(s_known + 6 * h);
      const float2 ky = *reinterpret_cast<const float2 *>(s_known + 6 * h + 2);
      const float2 kz = *reinterpret_cast<const float2 *>(s_known + 6 * h + 4);
      const float2 d = sqdist2_rn(u2x, u2y, u2z, kx, ky, kz);
      if (fminf(d.x, d.y) < b3) {  // rare after the first few dozen points
#pragma unroll
        for (int e = 0; e < 2; ++e) {  // ascending index order, as the reference scans
          const float dd = e ? d.y : d.x;
          const int kk = min(t0 + 2 * h + e, t0 + tn - 1);
          if (dd < b3 && !(e == 1 && 2 * h + 1 >= tn)) {  // interpolate_gpu.cu:34-49, strict '<'
            if (dd < b1) {
              b3 = b2, i3 = i2;
              b2 = b1, i2 = i1;
              b1 = dd, i1 = kk;
            } else if (dd < b2) {
              b3 = b2, i3 = i2;
              b2 = dd, i2 = kk;
            } else {
              b3 = dd, i3 = kk;
            }
          }
        }
      }
    }
  }
  if (j < n) {
    float *dj = dist2 + (static_cast<size_t>(cloud) * n + j) * 3;
    int *ij = idx + (static_cast<size_t>(cloud) * n + j) * 3;
    dj[0] = b1, dj[1] = b2, dj[2] = b3;
    ij[0] = i1, ij[1] = i2, ij[2] = i3;
  }
}

}  // namespace pn2

extern "C" int pn2_three_nn(int b, int n, int m, const float *unknown,
                            const float *known, float *dist2, int *idx,
                            void *stream) {
  using namespace pn2;
  if (b < 0 || n < 0 || m < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || n == 0) return PN2_OK;
  PN2_CHECK_PTRS(unknown, dist2, idx);
  if (m > 0) PN2_CHECK_PTRS(known);
  if (b > 65535) return PN2_ERR_UNSUPPORTED;
  const size_t smem = sizeof(float) * 6 * ((static_cast<size_t>(m < kNnTilePts ? m : kNnTilePts) + 1) / 2);
  cudaError_t e = cudaFuncSetAttribute(three_nn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(sizeof(float) * 3 * kNnTilePts));
  if (e != cudaSuccess) return static_cast<int>(e);
  three_nn_kernel<<<dim3((n + kNnThreads - 1) / kNnThreads, b), kNnThreads, smem,
                    static_cast<cudaStream_t>(stream)>>>(n, m, unknown, known, dist2, idx);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
