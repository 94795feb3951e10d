// Adjoints of the two row gathers as SEGMENTED GATHER-SUMS over an inverse map (sm_100a).
//
// The reference scatters: group_points_grad / three_interpolate_grad issue one global atomicAdd per
// gathered element into a zero-filled tensor (/root/reference/pointnet2/_ext_src/src/
// group_points_gpu.cu:43-75, interpolate_gpu.cu:116-154): nondeterministic summation order, a zero
// fill, and the atomic unit as the limiter.  The gathers' index tables depend on coordinates only, so
// their INVERSE -- for every source point, the sorted list of positions that read it -- can be built
// once per batch next to the ball query / 3-NN that produced them (geometry.build_plan, on the side
// stream).  The adjoint is then a gather again:
//
//     grad_src[b, k, :] = sum over e in segment(b, k), ascending:  w[e] * grad_rows[position(e), :]
//
// every output row written exactly once (no zero fill, no atomics), in a fixed order (bit-reproducible),
// every row of grad_rows read once with 128-bit loads by the warp that owns its source point.
//
// MEASURED (B200, B = 32, L2 flushed, tools/time_rows.py; profiles/r02_rows_inverse_vs_scatter.txt): the
// gather is NOT faster than the scatter it replaces.  FP0 (8192 <- 1024, 128 channels): 82 us against
// 89 us; SA2 (ct = 68): 55 against 28 us.  The rows a segment reads are scattered over a 70-140 MB
// tensor in 272-544 byte pieces (the points of a chunk are stored in sampling order, not in space), which
// HBM serves at ~1.9 TB/s, while the scatter streams grad_rows sequentially and its vector reductions hit
// an L2-resident 4-17 MB target.  The step therefore keeps the scatter by default
// (fused.INVERSE_ADJOINT = False); this path is what a caller that needs bit-reproducible gradients
// switches on, at +0.2 ms of inverse-map construction on the geometry stream.
//
//   pn2_inverse_map               keys [b, n_items] in [0, n_src) -> offsets [b, n_src + 1], items [b, n_items]
//                                 (item ids of every key, ascending): one CTA per cloud, counting sort in
//                                 shared memory with an ordered (schedule-independent) fill
//   pn2_sa_group_rows_grad_inv    adjoint of pn2_sa_group_rows   (QueryAndGroup, pointnet2_utils.py:317-359)
//   pn2_fp_interp_rows_grad_inv   adjoint of pn2_fp_interp_rows  (three_interpolate + cat,
//                                 pointnet2_modules.py:395-412)
#include "pn2_common.cuh"

namespace pn2 {
namespace inv {

constexpr int kBuildThreads = 1024;
constexpr int kMaxSrc = 8192;        // counters / offsets in shared memory
constexpr int kMaxItems = 40960;     // items in shared memory (160 KB)

// One CTA per cloud.  Shared memory: cnt[n_src] | start[n_src + 1] | items[n_items] | scratch[32].
__global__ void __launch_bounds__(kBuildThreads) inverse_map_kernel(int n_src, int n_items, const int *__restrict__ keys,
                                                                    int *__restrict__ offsets, int *__restrict__ items) {
  extern __shared__ int sm[];
  int *cnt = sm, *start = cnt + n_src, *it = start + n_src + 1, *scratch = it + n_items;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int *kc = keys + static_cast<size_t>(blockIdx.x) * n_items;
  for (int k = tid; k < n_src; k += kBuildThreads) cnt[k] = 0;
  __syncthreads();
  for (int e = tid; e < n_items; e += kBuildThreads) {
    const int k = __ldg(kc + e);
    if (k >= 0 && k < n_src) atomicAdd(&cnt[k], 1);
  }
  __syncthreads();
  // exclusive scan of cnt: every thread owns `per` consecutive keys
  const int per = (n_src + kBuildThreads - 1) / kBuildThreads;
  int local = 0;
  for (int i = 0; i < per; ++i) {
    const int k = tid * per + i;
    if (k < n_src) local += cnt[k];
  }
  int incl = local;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = scratch[lane];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    scratch[lane] = w;
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? scratch[warp - 1] : 0);
  for (int i = 0; i < per; ++i) {
    const int k = tid * per + i;
    if (k < n_src) {
      const int c = cnt[k];
      start[k] = run;
      cnt[k] = 0;  // becomes the fill cursor
      run += c;
    }
  }
  if (tid == kBuildThreads - 1) start[n_src] = run;  // == number of in-range items
  __syncthreads();
  // Ordered fill by ONE warp: items are visited in ascending order, 32 per step; lanes holding the same key
  // rank themselves with match_any, the lowest of them advances the key's cursor.  Every segment is then
  // in ascending item order whatever the schedule (an atomic fill would need a sort afterwards, and
  // ball-query padding makes some segments hundreds of items long).
  if (warp == 0) {
    for (int base = 0; base < n_items; base += 128) {
      int key[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = base + 32 * u + lane;
        key[u] = e < n_items ? __ldg(kc + e) : -1;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int e = base + 32 * u + lane;
        const bool ok = key[u] >= 0 && key[u] < n_src;
        const unsigned peers = __match_any_sync(0xffffffffu, ok ? key[u] : -1 - lane);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int first = 0;
        if (ok && lane == leader) {
          first = cnt[key[u]];
          cnt[key[u]] = first + __popc(peers);
        }
        first = __shfl_sync(0xffffffffu, first, leader);
        if (ok) it[start[key[u]] + first + rank] = e;
        __syncwarp();
      }
    }
  }
  __syncthreads();
  int *oc = offsets + static_cast<size_t>(blockIdx.x) * (n_src + 1);
  int *ic = items + static_cast<size_t>(blockIdx.x) * n_items;
  for (int k = tid; k <= n_src; k += kBuildThreads) oc[k] = start[k];
  const int total = start[n_src];
  for (int e = tid; e < n_items; e += kBuildThreads) ic[e] = e < total ? it[e] : -1;
}

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

// Warp per (cloud, source point): lanes own the float4 column groups c4 = lane, lane + 32, ... (kIters of
// them) of the output row.  The segment's items (and weights) are read 32 at a time, one per lane, and
// broadcast by shuffles; kU rows are in flight per step (the rows of a segment are scattered over the
// tensor: what bounds this kernel is how many independent row reads a warp keeps outstanding).
//   kWeighted: FP interpolation (item e = 3 * i + t reads row i with weight w[e]); else SA grouping
//   (item e = position j * ns + s reads row e with weight 1).
template <bool kWeighted, int kIters, int kU>
__global__ void __launch_bounds__(256)
seg_gather_sum_kernel(int total_src, int n_src, int n_items, int rows_per_cloud, int ld_rows4, int cols4,
                      int ld_out4, const float *__restrict__ grad_rows, const int *__restrict__ offsets,
                      const int *__restrict__ items, const float *__restrict__ weight, float *__restrict__ out) {
  const int w = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (w >= total_src) return;
  const int cloud = w / n_src, k = w - cloud * n_src;
  const int *oc = offsets + static_cast<size_t>(cloud) * (n_src + 1);
  const int s0 = __ldg(oc + k), s1 = __ldg(oc + k + 1);
  const int *ic = items + static_cast<size_t>(cloud) * n_items;
  const float *wc = kWeighted ? weight + static_cast<size_t>(cloud) * n_items : nullptr;
  const float4 *rows = reinterpret_cast<const float4 *>(grad_rows) + static_cast<size_t>(cloud) * rows_per_cloud * ld_rows4;
  float4 acc[kIters];
#pragma unroll
  for (int q = 0; q < kIters; ++q) acc[q] = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int base = s0; base < s1; base += 32) {
    const int cnt = min(32, s1 - base);
    int my_row = 0;
    float my_w = 1.f;
    if (lane < cnt) {
      const int e = __ldg(ic + base + lane);
      my_row = kWeighted ? e / 3 : e;
      if (kWeighted) my_w = __ldg(wc + e);
    }
    for (int j0 = 0; j0 < cnt; j0 += kU) {
      float4 v[kU][kIters];
      float ww[kU];
#pragma unroll
      for (int u = 0; u < kU; ++u) {
        if (j0 + u < cnt) {   // warp-uniform
          const int r = __shfl_sync(0xffffffffu, my_row, j0 + u);
          ww[u] = __shfl_sync(0xffffffffu, my_w, j0 + u);
#pragma unroll
          for (int q = 0; q < kIters; ++q) {
            const int c4 = lane + 32 * q;
            v[u][q] = c4 < cols4 ? __ldg(rows + static_cast<size_t>(r) * ld_rows4 + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      }
#pragma unroll
      for (int u = 0; u < kU; ++u)   // ascending item order
        if (j0 + u < cnt) {
#pragma unroll
          for (int q = 0; q < kIters; ++q) {
            acc[q].x = fmaf(ww[u], v[u][q].x, acc[q].x), acc[q].y = fmaf(ww[u], v[u][q].y, acc[q].y);
            acc[q].z = fmaf(ww[u], v[u][q].z, acc[q].z), acc[q].w = fmaf(ww[u], v[u][q].w, acc[q].w);
          }
        }
    }
  }
  float4 *dst = reinterpret_cast<float4 *>(out) + static_cast<size_t>(w) * ld_out4;
#pragma unroll
  for (int q = 0; q < kIters; ++q) {
    const int c4 = lane + 32 * q;
    if (c4 < cols4) dst[c4] = acc[q];
  }
}

// grad_skip[b, i, :c1] = grad_rows[b * n + i, c2 : c2 + c1]
__global__ void __launch_bounds__(256)
skip_grad_kernel(size_t total, int c1, int c2, int ct, const float *__restrict__ grad_rows, float *__restrict__ gs) {
  const size_t i = static_cast<size_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= total) return;
  const size_t p = i / c1;
  const int c = static_cast<int>(i - p * c1);
  gs[i] = __ldg(grad_rows + p * ct + c2 + c);
}

template <bool kWeighted>
int launch_seg(int total_src, int n_src, int n_items, int rows_per_cloud, int ld_rows4, int cols4, int ld_out4,
               const float *grad_rows, const int *offsets, const int *items, const float *weight, float *out,
               cudaStream_t s) {
  const int blocks = (total_src + 7) / 8;
  const int iters = (cols4 + 31) / 32;
#define PN2_SEG(IT, U)                                                                                     \
  seg_gather_sum_kernel<kWeighted, IT, U><<<blocks, 256, 0, s>>>(total_src, n_src, n_items, rows_per_cloud, \
                                                                 ld_rows4, cols4, ld_out4, grad_rows, offsets, \
                                                                 items, weight, out)
  // (kU = 4 everywhere: 16 rows in flight per warp measured 2x SLOWER at FP0 -- 128 registers, 2 CTAs per SM)
  if (iters <= 1) PN2_SEG(1, 4);
  else if (iters == 2) PN2_SEG(2, 4);
  else if (iters <= 4) PN2_SEG(4, 4);
  else if (iters <= 8) PN2_SEG(8, 2);
  else return PN2_ERR_UNSUPPORTED;
#undef PN2_SEG
  return 0;
}

}  // namespace inv
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_inverse_map_supported(int n_src, int n_items) {
  if (n_src <= 0 || n_items <= 0 || n_src > inv::kMaxSrc || n_items > inv::kMaxItems) return 0;
  return sizeof(int) * (static_cast<size_t>(2) * n_src + 1 + n_items + 32) <= 220 * 1024 ? 1 : 0;
}

extern "C" int pn2_inverse_map(int b, int n_src, int n_items, const int *keys, int *offsets, int *items,
                               void *stream) {
  if (b <= 0) return PN2_ERR_INVALID_ARG;
  if (!pn2_inverse_map_supported(n_src, n_items)) return PN2_ERR_UNSUPPORTED;
  PN2_CHECK_PTRS(keys, offsets, items);
  const size_t smem = sizeof(int) * (static_cast<size_t>(2) * n_src + 1 + n_items + 32);
  cudaError_t e = cudaFuncSetAttribute(inv::inverse_map_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  inv::inverse_map_kernel<<<b, inv::kBuildThreads, smem, static_cast<cudaStream_t>(stream)>>>(n_src, n_items, keys,
                                                                                             offsets, items);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_sa_group_rows_grad_inv(int b, int n, int npoint, int nsample, int ct, const float *grad_rows,
                                          const int *offsets, const int *items, float *grad_table, void *stream) {
  if (b <= 0 || n <= 0 || npoint <= 0 || nsample <= 0 || ct < 4 || ct % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(grad_rows, offsets, items, grad_table);
  if (!aligned16(grad_rows) || !aligned16(grad_table)) return PN2_ERR_UNSUPPORTED;
  const int st = inv::launch_seg<false>(b * n, n, npoint * nsample, npoint * nsample, ct / 4, ct / 4, ct / 4,
                                        grad_rows, offsets, items, nullptr, grad_table,
                                        static_cast<cudaStream_t>(stream));
  if (st != 0) return st;
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_fp_interp_rows_grad_inv(int b, int n, int m, int c2, int c1, const float *grad_rows,
                                           const int *offsets, const int *items, const float *weight,
                                           float *grad_known, float *grad_skip, void *stream) {
  if (b <= 0 || n <= 0 || m <= 0 || c2 <= 0 || c1 < 0 || c2 % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(grad_rows, offsets, items, weight, grad_known);
  const int ct = (c2 + c1 + 3) / 4 * 4;
  if (!aligned16(grad_known) || !aligned16(grad_rows)) return PN2_ERR_UNSUPPORTED;
  if (grad_skip != nullptr && !aligned4(grad_skip)) return PN2_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int st = inv::launch_seg<true>(b * m, m, 3 * n, n, ct / 4, c2 / 4, c2 / 4, grad_rows, offsets, items, weight,
                                       grad_known, s);
  if (st != 0) return st;
  count_launch();
  if (grad_skip != nullptr && c1 > 0) {
    const size_t total = static_cast<size_t>(b) * n * c1;
    inv::skip_grad_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, s>>>(total, c1, c2, ct, grad_rows,
                                                                                    grad_skip);
    count_launch();
  }
  PN2_RETURN_LAUNCH_STATUS();
}
