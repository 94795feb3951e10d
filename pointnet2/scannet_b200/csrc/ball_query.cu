// Ball query for sm_100a.
//
// Replaces query_ball_point_kernel / _wrapper of the reference
// (/root/reference/pointnet2/_ext_src/src/ball_query_gpu.cu:9-54): one CTA per
// cloud, one THREAD per centre scanning up to n points from global memory.
//
// Here a WARP owns a centre: 32 lanes test 32 consecutive points,
// __ballot_sync + popc ranks the hits in ascending index order (the
// reference's order), and the warp leaves the scan as soon as nsample hits are
// found -- per centre, not per slowest-of-32-centres as in a thread-per-centre
// layout.  The result row is assembled in shared memory and written with one
// coalesced store, including the reference's padding (first hit repeated; all
// zeros when the ball is empty, ball_query.cpp:19-21 + ball_query_gpu.cu:34-38).
//
// Clouds of up to kBqStagePts points (96 KB of xyz) are staged once per CTA
// into shared memory with a 1-D TMA bulk copy (cp.async.bulk + mbarrier) and
// scanned from there; larger clouds (whole scenes) are scanned straight from
// L2 with coalesced loads.
#include "pn2_common.cuh"

namespace pn2 {

constexpr int kBqThreads = 256;
constexpr int kBqWarps = kBqThreads / 32;
constexpr int kBqStagePts = 8192;  // 96 KB -> two CTAs resident per SM

struct BqParams {
  int n, m, nsample;
  int centres_per_cta;
  int stage_floats;  // shared-memory floats reserved for the staged cloud
  float radius2;
};

// Scan points [0, n) of `pts` (shared or global, (n,3) fp32) for centre q.
// On return row[0..min(cnt,nsample)) holds the hits in ascending index order.
__device__ __forceinline__ void bq_scan(const float *pts, int n, float qx, float qy,
                                        float qz, float radius2, int nsample,
                                        int lane, int *row, int &cnt, int &first) {
  cnt = 0, first = 0;
  for (int base = 0; base < n && cnt < nsample; base += 32) {
    const int k = base + lane;
    bool hit = false;
    if (k < n)
      hit = sqdist_rn(qx, qy, qz, pts[3 * k], pts[3 * k + 1], pts[3 * k + 2]) < radius2;
    const unsigned mask = __ballot_sync(0xffffffffu, hit);
    if (mask) {
      if (cnt == 0) first = base + __ffs(mask) - 1;
      const int rank = cnt + __popc(mask & ((1u << lane) - 1u));
      if (hit && rank < nsample) row[rank] = k;
      cnt += __popc(mask);
    }
  }
}

template <bool kStaged>
__global__ void __launch_bounds__(kBqThreads, 2)
ball_query_kernel(BqParams p, const float *__restrict__ new_xyz,
                  const float *__restrict__ xyz, int *__restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ uint64_t s_bar;
  float *s_xyz = reinterpret_cast<float *>(smem_raw);  // [n][3] when staged
  int *s_rows = reinterpret_cast<int *>(smem_raw) + p.stage_floats;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cloud = blockIdx.y;
  const float *pts = xyz + static_cast<size_t>(cloud) * p.n * 3;
  const float *ctr = new_xyz + static_cast<size_t>(cloud) * p.m * 3;
  int *out = idx + static_cast<size_t>(cloud) * p.m * p.nsample;
  int *row = s_rows + warp * p.nsample;

  if (kStaged && p.n > 0) {
    const uint32_t bytes = static_cast<uint32_t>(p.n) * 12u;
    const bool bulk = (reinterpret_cast<uintptr_t>(pts) & 15u) == 0 && bytes % 16u == 0;
    if (bulk) {
      if (tid == 0) {
        mbar_init(&s_bar, 1);
        mbar_fence_init();
        mbar_expect_tx(&s_bar, bytes);
        tma_load_1d(s_xyz, pts, bytes, &s_bar);
      }
      __syncthreads();  // barrier initialised before anyone polls it
      mbar_wait(&s_bar, 0);
    } else {  // unaligned cloud base or odd size: plain coalesced copy
      for (int i = tid; i < p.n * 3; i += kBqThreads) s_xyz[i] = __ldg(pts + i);
      __syncthreads();
    }
  }
  const float *src = kStaged ? s_xyz : pts;

  const int c_begin = blockIdx.x * p.centres_per_cta;
  const int c_end = min(c_begin + p.centres_per_cta, p.m);
  for (int j = c_begin + warp; j < c_end; j += kBqWarps) {
    const float qx = __ldg(ctr + 3 * j), qy = __ldg(ctr + 3 * j + 1), qz = __ldg(ctr + 3 * j + 2);
    int cnt, first;
    bq_scan(src, p.n, qx, qy, qz, p.radius2, p.nsample, lane, row, cnt, first);
    __syncwarp();
    int *dst = out + static_cast<size_t>(j) * p.nsample;
    const int filled = min(cnt, p.nsample);
    for (int l = lane; l < p.nsample; l += 32) dst[l] = l < filled ? row[l] : first;
    __syncwarp();
  }
}

}  // namespace pn2

extern "C" int pn2_query_ball_point(int b, int n, int m, float radius,
                                    int nsample, const float *new_xyz,
                                    const float *xyz, int *idx, void *stream) {
  using namespace pn2;
  if (b < 0 || n < 0 || m < 0 || nsample < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || m == 0 || nsample == 0) return PN2_OK;
  PN2_CHECK_PTRS(new_xyz, idx);
  if (n > 0) PN2_CHECK_PTRS(xyz);
  if (nsample > 2048 || b > 65535) return PN2_ERR_UNSUPPORTED;

  BqParams p;
  p.n = n, p.m = m, p.nsample = nsample;
  p.radius2 = radius * radius;  // fp32 product, ball_query_gpu.cu:22
  // two CTAs resident per SM; at least one centre per warp of a CTA
  int ctas_per_cloud = (2 * kNumSMs + b - 1) / b;
  const int max_ctas = (m + kBqWarps - 1) / kBqWarps;
  if (ctas_per_cloud > max_ctas) ctas_per_cloud = max_ctas;
  p.centres_per_cta = (m + ctas_per_cloud - 1) / ctas_per_cloud;
  ctas_per_cloud = (m + p.centres_per_cta - 1) / p.centres_per_cta;

  const bool staged = n <= kBqStagePts;
  p.stage_floats = staged ? (3 * n + 3) / 4 * 4 : 0;  // keeps the row buffers 16-byte aligned
  const size_t smem = sizeof(float) * p.stage_floats + sizeof(int) * kBqWarps * nsample;
  auto kern = staged ? ball_query_kernel<true> : ball_query_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  kern<<<dim3(ctas_per_cloud, b), kBqThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      p, new_xyz, xyz, idx);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
