// Ball query for sm_100a.
//
// Replaces query_ball_point_kernel / _wrapper of the reference
// (/root/reference/pointnet2/_ext_src/src/ball_query_gpu.cu:9-54): one CTA per
// cloud, one THREAD per centre scanning up to n points from global memory.
//
// Here a WARP owns four centres: 32 lanes test 32 consecutive points against
// all four (packed fp32: two distances per instruction), __ballot_sync + popc
// ranks the hits in ascending index order (the reference's order), and the warp
// leaves the scan as soon as its centres have nsample hits each -- not when
// the slowest of 32 centres has, as in a thread-per-centre layout.  The result row is assembled in shared memory and written with one
// coalesced store, including the reference's padding (first hit repeated; all
// zeros when the ball is empty, ball_query.cpp:19-21 + ball_query_gpu.cu:34-38).
//
// Clouds of up to kBqStagePts points (96 KB of xyz) are staged once per CTA
// into shared memory with a 1-D TMA bulk copy (cp.async.bulk + mbarrier) and
// scanned from there; larger clouds (whole scenes) are scanned straight from
// L2 with coalesced loads.
#include <limits.h>
#include <stdlib.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kBqThreads = 512;
constexpr int kBqWarps = kBqThreads / 32;
constexpr int kBqStagePts = 8192;  // 96 KB -> two CTAs resident per SM

struct BqParams {
  int n, m, nsample;
  int centres_per_cta;
  int stage_floats;  // shared-memory floats reserved for the staged cloud
  float radius2;
};

constexpr int kBqQ = 4;  // centres a warp tests against every point it loads

// A warp scans points [0, n) of `pts` (shared or global, (n,3) fp32) for kBqQ centres at
// once: each lane loads ONE point per step and evaluates it against the four centres with
// two packed distance evaluations (centres 0|1 and 2|3 share an instruction), so the point
// load, the loop and the address arithmetic are amortised 4x.  rows[q][0..min(cnt,nsample))
// receive the hits of centre q in ascending index order.
__device__ __forceinline__ void bq_scan4(const float *pts, int n, const float2 (&qx)[2],
                                         const float2 (&qy)[2], const float2 (&qz)[2], float radius2,
                                         int nsample, int nq, int lane, int *rows, int (&cnt)[kBqQ],
                                         int (&first)[kBqQ]) {
#pragma unroll
  for (int q = 0; q < kBqQ; ++q) cnt[q] = q < nq ? 0 : nsample, first[q] = 0;
  const unsigned lt = (1u << lane) - 1u;
  for (int base = 0; base < n; base += 32) {
    const int k = base + lane;
    const bool in = k < n;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (in) px = pts[3 * k], py = pts[3 * k + 1], pz = pts[3 * k + 2];
    const float2 p2x = make_float2(px, px), p2y = make_float2(py, py), p2z = make_float2(pz, pz);
    const float2 d01 = sqdist2_rn(qx[0], qy[0], qz[0], p2x, p2y, p2z);
    const float2 d23 = sqdist2_rn(qx[1], qy[1], qz[1], p2x, p2y, p2z);
    const bool h0 = in && d01.x < radius2, h1 = in && d01.y < radius2;
    const bool h2 = in && d23.x < radius2, h3 = in && d23.y < radius2;
    // hits are rare (tens per centre among thousands of points): one vote decides the step
    if (!__any_sync(0xffffffffu, h0 | h1 | h2 | h3)) continue;
    const bool hit[kBqQ] = {h0, h1, h2, h3};
    bool open = false;
#pragma unroll
    for (int q = 0; q < kBqQ; ++q) {
      const unsigned mask = __ballot_sync(0xffffffffu, hit[q]);
      if (mask && cnt[q] < nsample) {
        if (cnt[q] == 0) first[q] = base + __ffs(mask) - 1;
        const int rank = cnt[q] + __popc(mask & lt);
        if (hit[q] && rank < nsample) rows[q * nsample + rank] = k;
        cnt[q] += __popc(mask);
      }
      open |= cnt[q] < nsample;
    }
    if (!open) break;  // every centre of this warp has its nsample hits
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
  int *rows = s_rows + warp * kBqQ * p.nsample;

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
  for (int j0 = c_begin + warp * kBqQ; j0 < c_end; j0 += kBqWarps * kBqQ) {
    const int nq = min(kBqQ, c_end - j0);
    float c[kBqQ][3];
#pragma unroll
    for (int q = 0; q < kBqQ; ++q) {
      const int j = j0 + (q < nq ? q : 0);
      c[q][0] = __ldg(ctr + 3 * j), c[q][1] = __ldg(ctr + 3 * j + 1), c[q][2] = __ldg(ctr + 3 * j + 2);
    }
    const float2 qx[2] = {make_float2(c[0][0], c[1][0]), make_float2(c[2][0], c[3][0])};
    const float2 qy[2] = {make_float2(c[0][1], c[1][1]), make_float2(c[2][1], c[3][1])};
    const float2 qz[2] = {make_float2(c[0][2], c[1][2]), make_float2(c[2][2], c[3][2])};
    int cnt[kBqQ], first[kBqQ];
    bq_scan4(src, p.n, qx, qy, qz, p.radius2, p.nsample, nq, lane, rows, cnt, first);
    __syncwarp();
#pragma unroll
    for (int q = 0; q < kBqQ; ++q) {
      if (q < nq) {
        int *dst = out + static_cast<size_t>(j0 + q) * p.nsample;
        const int filled = min(cnt[q], p.nsample);
        for (int l = lane; l < p.nsample; l += 32) dst[l] = l < filled ? rows[q * p.nsample + l] : first[q];
      }
    }
    __syncwarp();
  }
}

// ---------------------------------------------------------------------------------------------
// Grid variant (512 <= n <= 8192 points per cloud, the training chunks): the cloud is counting-
// sorted into a spatial hash of cell size 1.001 * radius in shared memory, and a centre only
// tests the points of its 27 neighbouring cells (tens of candidates instead of all n: the balls
// of the SSG levels hold ~5-15 points).  A warp owns a centre: lane l < 27 walks cell l's bucket,
// hits are appended to a per-warp list, and the nsample SMALLEST indices are extracted in
// ascending order with warp minima -- the reference's "first nsample hits in index order".  The
// distance is the same un-contracted fp32 expression, evaluated on a subset of the same pairs, so
// the result is bit-identical to the full scan (a list overflow falls back to the full scan).
// ---------------------------------------------------------------------------------------------
constexpr int kGridThreads = 1024, kGridWarps = kGridThreads / 32;
constexpr int kGridBuckets = 4096, kGridHitCap = 128;
constexpr float kMaxCellCoord = 4096.f;   // largest |coordinate / cell size| the 27-cell search is trusted for

__global__ void __launch_bounds__(kGridThreads, 1)
ball_query_grid_kernel(BqParams p, float inv, const float *__restrict__ new_xyz,
                       const float *__restrict__ xyz, int *__restrict__ idx) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float4 *s_pts = reinterpret_cast<float4 *>(smem_raw);                       // [n]
  int *s_start = reinterpret_cast<int *>(s_pts + p.n);                       // [T + 1]
  int *s_cnt = s_start + kGridBuckets + 1;                                    // [T]
  int *s_hits = s_cnt + kGridBuckets;                                         // [warps][cap]
  int *s_rows = s_hits + kGridWarps * kGridHitCap;                            // [warps][nsample]
  int *s_hitcnt = s_rows + kGridWarps * p.nsample;                            // [warps]
  int *s_scratch = s_hitcnt + kGridWarps;                                     // [32]

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int cloud = blockIdx.y;
  const float *pts = xyz + static_cast<size_t>(cloud) * p.n * 3;
  const float *ctr = new_xyz + static_cast<size_t>(cloud) * p.m * 3;
  int *out = idx + static_cast<size_t>(cloud) * p.m * p.nsample;
  // inv = 1 / (1.001 * radius): cell size just above the radius (computed on the host)

  for (int i = tid; i < kGridBuckets; i += kGridThreads) s_cnt[i] = 0;
  __syncthreads();
  grid_build<kGridThreads>(pts, p.n, inv, kGridBuckets, s_cnt, s_start, s_pts, s_scratch);
  // The 27-cell search is complete only while the fp32 rounding of x * inv stays inside the 0.1 % margin
  // of the cell size and no cell index is clamped: |x * inv| <= kMaxCellCoord (rounding error of one
  // product <= 4096 * 2^-24 = 2.4e-4 of a cell, two products < 1e-3).  A cloud (or a centre) outside that
  // range -- millimetre radii on metre-scale coordinates, chunks at large world offsets, non-finite input --
  // takes the reference's scan over all points instead, so the result never depends on the grid.
  bool far = false;
  for (int k = tid; k < p.n; k += kGridThreads) {
    const float4 q = s_pts[k];
    far |= !(fabsf(q.x * inv) <= kMaxCellCoord && fabsf(q.y * inv) <= kMaxCellCoord && fabsf(q.z * inv) <= kMaxCellCoord);
  }
  const bool scan_all = __syncthreads_or(far) != 0;

  const int j0 = blockIdx.x * p.centres_per_cta;
  const int j1 = min(j0 + p.centres_per_cta, p.m);
  int *hits = s_hits + warp * kGridHitCap;
  int *row = s_rows + warp * p.nsample;
  for (int j = j0 + warp; j < j1; j += kGridWarps) {
    const float cx = __ldg(ctr + 3 * j), cy = __ldg(ctr + 3 * j + 1), cz = __ldg(ctr + 3 * j + 2);
    const bool use_scan = scan_all || !(fabsf(cx * inv) <= kMaxCellCoord && fabsf(cy * inv) <= kMaxCellCoord &&
                                         fabsf(cz * inv) <= kMaxCellCoord);   // warp-uniform
    int tx = grid_cell(cx, inv), ty = grid_cell(cy, inv), tz = grid_cell(cz, inv);
    int s = 0, e = 0;
    if (lane < 27 && !use_scan) {
      tx += lane % 3 - 1, ty += (lane / 3) % 3 - 1, tz += lane / 9 - 1;
      const unsigned b = grid_hash(tx, ty, tz, kGridBuckets - 1);
      s = s_start[b], e = s_start[b + 1];
    }
    if (lane == 0) s_hitcnt[warp] = 0;
    __syncwarp();
    const int steps = __reduce_max_sync(0xffffffffu, e - s);
    for (int t = 0; t < steps; ++t) {
      if (s + t < e) {
        const float4 q = s_pts[s + t];
        // the bucket may also hold points of other cells (hash collisions): only this cell's count
        if (grid_cell(q.x, inv) == tx && grid_cell(q.y, inv) == ty && grid_cell(q.z, inv) == tz &&
            sqdist_rn(cx, cy, cz, q.x, q.y, q.z) < p.radius2) {
          const int pos = atomicAdd(&s_hitcnt[warp], 1);
          if (pos < kGridHitCap) hits[pos] = __float_as_int(q.w);
        }
      }
    }
    __syncwarp();
    const int H = use_scan ? kGridHitCap + 1 : s_hitcnt[warp];
    int *dst = out + static_cast<size_t>(j) * p.nsample;
    if (H > kGridHitCap) {
      // more hits than the list holds: the reference's scan over all points, in index order
      int cnt = 0, first = 0;
      const unsigned lt = (1u << lane) - 1u;
      for (int base = 0; base < p.n && cnt < p.nsample; base += 32) {
        const int k = base + lane;
        bool hit = false;
        if (k < p.n)
          hit = sqdist_rn(cx, cy, cz, __ldg(pts + 3 * k), __ldg(pts + 3 * k + 1), __ldg(pts + 3 * k + 2)) < p.radius2;
        const unsigned mask = __ballot_sync(0xffffffffu, hit);
        if (mask == 0u) continue;
        if (cnt == 0) first = base + __ffs(mask) - 1;
        const int slot = cnt + __popc(mask & lt);
        if (hit && slot < p.nsample) row[slot] = k;
        cnt += __popc(mask);
      }
      __syncwarp();
      const int filled = min(cnt, p.nsample);
      for (int l = lane; l < p.nsample; l += 32) dst[l] = l < filled ? row[l] : first;
    } else {
      int v[kGridHitCap / 32];
#pragma unroll
      for (int i = 0; i < kGridHitCap / 32; ++i) v[i] = lane + 32 * i < H ? hits[lane + 32 * i] : INT_MAX;
      const int filled = min(H, p.nsample);
      for (int l = 0; l < filled; ++l) {
        int mine = v[0];
#pragma unroll
        for (int i = 1; i < kGridHitCap / 32; ++i) mine = min(mine, v[i]);
        const int best = __reduce_min_sync(0xffffffffu, mine);
#pragma unroll
        for (int i = 0; i < kGridHitCap / 32; ++i)
          if (v[i] == best) v[i] = INT_MAX;  // indices are unique: exactly one lane holds it
        if (lane == 0) row[l] = best;
      }
      __syncwarp();
      const int first = filled > 0 ? row[0] : 0;  // empty ball: zeros (ball_query.cpp:19-21)
      for (int l = lane; l < p.nsample; l += 32) dst[l] = l < filled ? row[l] : first;
    }
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
  if (nsample > 512 || b > 65535) return PN2_ERR_UNSUPPORTED;

  BqParams p;
  p.n = n, p.m = m, p.nsample = nsample;
  p.radius2 = radius * radius;  // fp32 product, ball_query_gpu.cu:22
  // one wave: at most two CTAs resident per SM (shared-memory stage), never a second, partial wave
  int ctas_per_cloud = 2 * kNumSMs / b;
  if (ctas_per_cloud < 1) ctas_per_cloud = 1;
  const int max_ctas = (m + kBqWarps * kBqQ - 1) / (kBqWarps * kBqQ);
  if (ctas_per_cloud > max_ctas) ctas_per_cloud = max_ctas;
  p.centres_per_cta = (m + ctas_per_cloud - 1) / ctas_per_cloud;
  ctas_per_cloud = (m + p.centres_per_cta - 1) / p.centres_per_cta;

  // training-chunk sizes: spatial hash in shared memory, 27 cells per centre instead of all n points
  if (n >= 512 && n <= 8192 && nsample <= 128 && radius >= 1e-3f && radius <= 1e3f &&
      getenv("PN2_BQ_NO_GRID") == nullptr) {
    int splits = kNumSMs / b;  // CTAs per cloud (each rebuilds the grid: ~n/1024 steps)
    if (splits < 1) splits = 1;
    if (splits > (m + 4 * kGridWarps - 1) / (4 * kGridWarps)) splits = (m + 4 * kGridWarps - 1) / (4 * kGridWarps);
    p.centres_per_cta = (m + splits - 1) / splits;
    splits = (m + p.centres_per_cta - 1) / p.centres_per_cta;
    p.stage_floats = 0;
    const size_t gsmem = sizeof(float4) * n + sizeof(int) * (2 * kGridBuckets + 1 + kGridWarps * kGridHitCap +
                                                             kGridWarps * nsample + kGridWarps + 32);
    cudaError_t ge = cudaFuncSetAttribute(ball_query_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                          static_cast<int>(gsmem));
    if (ge != cudaSuccess) return static_cast<int>(ge);
    ball_query_grid_kernel<<<dim3(splits, b), kGridThreads, gsmem, static_cast<cudaStream_t>(stream)>>>(
        p, 1.0f / (radius * 1.001f), new_xyz, xyz, idx);
    count_launch();
    PN2_RETURN_LAUNCH_STATUS();
  }
  const bool staged = n <= kBqStagePts;
  p.stage_floats = staged ? (3 * n + 3) / 4 * 4 : 0;  // keeps the row buffers 16-byte aligned
  const size_t smem = sizeof(float) * p.stage_floats + sizeof(int) * kBqWarps * kBqQ * nsample;
  auto kern = staged ? ball_query_kernel<true> : ball_query_kernel<false>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  kern<<<dim3(ctas_per_cloud, b), kBqThreads, smem, static_cast<cudaStream_t>(stream)>>>(
      p, new_xyz, xyz, idx);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
