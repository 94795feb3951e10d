// Furthest point sampling for sm_100a.
//
// Replaces furthest_point_sampling_kernel / _wrapper of the reference
// (/root/reference/pointnet2/_ext_src/src/sampling_gpu.cu:69-229).
//
// Design (one CTA per cloud, the whole working set on chip):
//   * every thread owns P points -- coordinates AND running minimum distance
//     live in registers for all m-1 iterations (the reference re-reads xyz and
//     round-trips `temp` through global memory every iteration);
//   * one iteration = P fused distance/min/arg-max updates per thread, a
//     two-instruction warp arg-max (REDUX.MAX on the float bits, REDUX.MIN on
//     the tie-break position), one 8-byte shared-memory slot per warp, ONE
//     __syncthreads (slots are double buffered), a second REDUX pair over the
//     per-warp slots, and one broadcast LDS.128 that fetches the winner's
//     coordinates and original index;
//   * the reference's tie-break is a property of its thread layout (strided
//     scan + shared-memory tree, sampling_gpu.cu:59-65,90-167).  In closed
//     form the winner among equal maxima minimises
//         pos(k) = bitrev_{log2 bs}(k mod bs) * q + (k div bs),
//     bs = the reference's block size for this n, q = ceil(n / bs).
//     Points are laid out BY pos, so "lowest position wins" reproduces it.
//   * points the reference skips (|p|^2 <= 1e-3 compared in double,
//     sampling_gpu.cu:100-101) and padding positions carry the sentinel
//     minimum -1.0f: fminf(d, -1) stays -1, which never beats a real
//     candidate (>= 0) -- exactly the reference's `best = -1` initial value --
//     and if nothing else exists position 0 (= index 0) wins, as `besti = 0`.
#include <math.h>

#include "pn2_common.cuh"

namespace pn2 {

constexpr unsigned kFull = 0xffffffffu;

struct FpsGeom {
  int n, m;
  int bs;      // reference block size (power of two, <= 512)
  int log2bs;
  int q;       // ceil(n / bs): points per reference thread
};

__device__ __forceinline__ int fps_pos_to_index(int p, const FpsGeom &g) {
  const int trev = p / g.q;
  const int j = p - trev * g.q;
  const int tid = g.log2bs ? static_cast<int>(__brev(trev) >> (32 - g.log2bs)) : 0;
  return tid + j * g.bs;
}
__device__ __forceinline__ int fps_index_to_pos(int k, const FpsGeom &g) {
  const int t = k & (g.bs - 1);
  const int trev = g.log2bs ? static_cast<int>(__brev(t) >> (32 - g.log2bs)) : 0;
  return trev * g.q + (k >> g.log2bs);
}

// Block-wide arg-max of (value desc, position asc).  `v` = float bits of this
// thread's best minimum distance (as signed int: -1.0f < every float >= 0),
// `pos` = its position.  Returns the winning position to every thread.
__device__ __forceinline__ unsigned block_argmax(int v, unsigned pos, int2 *slots,
                                                 int lane, int warp, int nwarps) {
  const int wmax = __reduce_max_sync(kFull, v);
  const unsigned wpos = __reduce_min_sync(kFull, v == wmax ? pos : 0xffffffffu);
  if (lane == 0) slots[warp] = make_int2(wmax, static_cast<int>(wpos));
  __syncthreads();
  int2 r = make_int2(INT_MIN, -1);
  if (lane < nwarps) r = slots[lane];
  const int gmax = __reduce_max_sync(kFull, r.x);
  return __reduce_min_sync(kFull, r.x == gmax ? static_cast<unsigned>(r.y)
                                              : 0xffffffffu);
}

// ---------------------------------------------------------------------------
// register-resident kernel: blockDim.x * P >= bs * q positions
// ---------------------------------------------------------------------------
template <int P, int MAXT>
__global__ void __launch_bounds__(MAXT, 1)
fps_resident_kernel(FpsGeom g, const float *__restrict__ dataset,
                    int *__restrict__ idxs) {
  extern __shared__ float4 s_pts[];  // [npos] (x, y, z, bits(original index))
  __shared__ int2 s_slots[2][32];

  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const int npos = g.bs * g.q;
  const float *pts = dataset + static_cast<size_t>(blockIdx.x) * g.n * 3;
  int *out = idxs + static_cast<size_t>(blockIdx.x) * g.m;

  // points are held as PAIRS (slots 2h, 2h+1 of this thread) so that one packed FADD2/FFMA2
  // updates two running minima; odd P (= 1) keeps a dummy second half
  constexpr int H = (P + 1) / 2;
  float2 x[H], y[H], z[H], t[H];
#pragma unroll
  for (int i = 0; i < 2 * H; ++i) {
    const int p = tid * P + i;
    const int k = (i < P && p < npos) ? fps_pos_to_index(p, g) : g.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    bool live = false;
    if (k < g.n) {
      px = __ldg(pts + 3 * k + 0);
      py = __ldg(pts + 3 * k + 1);
      pz = __ldg(pts + 3 * k + 2);
      live = !(static_cast<double>(sqnorm_rn(px, py, pz)) <= 1e-3);
    }
    const float t0 = live ? 1e10f : -1.0f;
    if (i & 1) x[i / 2].y = px, y[i / 2].y = py, z[i / 2].y = pz, t[i / 2].y = t0;
    else x[i / 2].x = px, y[i / 2].x = py, z[i / 2].x = pz, t[i / 2].x = t0;
    if (i < P && p < npos) s_pts[p] = make_float4(px, py, pz, __int_as_float(k < g.n ? k : 0));
  }
  if (tid == 0) out[0] = 0;
  __syncthreads();

  float4 c = s_pts[0];  // position 0 <-> index 0
  for (int j = 1; j < g.m; ++j) {
    const float2 cx = make_float2(c.x, c.x), cy = make_float2(c.y, c.y), cz = make_float2(c.z, c.z);
    float best;
    int bi = 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float2 d = sqdist2_rn(x[h], y[h], z[h], cx, cy, cz);
      t[h].x = fminf(d.x, t[h].x);
      t[h].y = fminf(d.y, t[h].y);
      if (h == 0) {
        best = t[0].x;
      } else if (t[h].x > best) {  // strict: lowest position of a thread's ties
        best = t[h].x;
        bi = 2 * h;
      }
      if (2 * h + 1 < P && t[h].y > best) {
        best = t[h].y;
        bi = 2 * h + 1;
      }
    }
    const unsigned win =
        block_argmax(__float_as_int(best), static_cast<unsigned>(tid * P + bi),
                     s_slots[j & 1], lane, warp, nwarps);
    c = s_pts[win];
    if (tid == 0) out[j] = __float_as_int(c.w);
  }
}

// ---------------------------------------------------------------------------
// streaming fallback for n > PN2_FPS_MAX_RESIDENT_N: minima in the caller's
// `temp` like the reference.  1024 threads; thread t scans k = t, t+1024, ...
// (k mod bs is constant per thread because 1024 % bs == 0, so a thread's
// positions ascend and strict '>' keeps its lowest-position tie).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(1024, 1)
fps_streaming_kernel(FpsGeom g, const float *__restrict__ dataset,
                     float *__restrict__ temp, int *__restrict__ idxs) {
  __shared__ int2 s_slots[2][32];
  const int tid = threadIdx.x;
  const int lane = tid & 31, warp = tid >> 5, nwarps = blockDim.x >> 5;
  const float *pts = dataset + static_cast<size_t>(blockIdx.x) * g.n * 3;
  float *tmp = temp + static_cast<size_t>(blockIdx.x) * g.n;
  int *out = idxs + static_cast<size_t>(blockIdx.x) * g.m;

  for (int k = tid; k < g.n; k += blockDim.x) {
    const float px = pts[3 * k], py = pts[3 * k + 1], pz = pts[3 * k + 2];
    const bool live = !(static_cast<double>(sqnorm_rn(px, py, pz)) <= 1e-3);
    tmp[k] = live ? 1e10f : -1.0f;
  }
  if (tid == 0) out[0] = 0;
  __syncthreads();

  int cur = 0;
  for (int j = 1; j < g.m; ++j) {
    const float cx = pts[3 * cur], cy = pts[3 * cur + 1], cz = pts[3 * cur + 2];
    float best = -2.0f;
    int bk = 0;
    for (int k = tid; k < g.n; k += blockDim.x) {
      const float d = sqdist_rn(pts[3 * k], pts[3 * k + 1], pts[3 * k + 2], cx, cy, cz);
      const float v = fminf(d, tmp[k]);
      tmp[k] = v;
      if (v > best) best = v, bk = k;
    }
    const unsigned pos = best > -2.0f ? static_cast<unsigned>(fps_index_to_pos(bk, g))
                                      : 0xfffffffeu;
    const unsigned win =
        block_argmax(__float_as_int(best), pos, s_slots[j & 1], lane, warp, nwarps);
    cur = fps_pos_to_index(static_cast<int>(win), g);
    if (tid == 0) out[j] = cur;
  }
}

template <int P, int MAXT>
static cudaError_t launch_resident(const FpsGeom &g, int b, int threads,
                                   const float *dataset, int *idxs,
                                   cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(g.bs) * g.q * sizeof(float4);
  auto kern = fps_resident_kernel<P, MAXT>;
  cudaError_t e = cudaFuncSetAttribute(
      kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kern<<<b, threads, smem, stream>>>(g, dataset, idxs);
  count_launch();
  return cudaGetLastError();
}

}  // namespace pn2

extern "C" int pn2_furthest_point_sampling(int b, int n, int m,
                                           const float *dataset, float *temp,
                                           int *idxs, void *stream) {
  using namespace pn2;
  if (b < 0 || n < 0 || m < 0) return PN2_ERR_INVALID_ARG;
  if (b == 0 || m == 0) return PN2_OK;
  if (n == 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dataset, idxs);
  cudaStream_t s = static_cast<cudaStream_t>(stream);

  FpsGeom g;
  g.n = n, g.m = m;
  g.bs = ref_block_threads(n);
  g.log2bs = 0;
  while ((1 << g.log2bs) < g.bs) ++g.log2bs;
  g.q = (n + g.bs - 1) / g.bs;
  const int npos = g.bs * g.q;

  if (n <= PN2_FPS_MAX_RESIDENT_N) {
    // at least 4 warps (one per SM sub-partition), then as many points per
    // thread as the register file allows: fewer warps = cheaper block arg-max
    auto threads_for = [&](int P) {
      int t = (npos + P - 1) / P;
      t = (t + 31) / 32 * 32;
      return t < 128 ? (npos < 128 ? (npos + 31) / 32 * 32 : 128) : t;
    };
    cudaError_t e;
    if (npos <= 128 * 1) e = launch_resident<1, 1024>(g, b, threads_for(1), dataset, idxs, s);
    else if (npos <= 128 * 2) e = launch_resident<2, 1024>(g, b, threads_for(2), dataset, idxs, s);
    else if (npos <= 128 * 4) e = launch_resident<4, 1024>(g, b, threads_for(4), dataset, idxs, s);
    else if (npos <= 1024 * 8) e = launch_resident<8, 1024>(g, b, threads_for(8), dataset, idxs, s);
    else return PN2_ERR_UNSUPPORTED;
    return static_cast<int>(e);
  }
  if (temp == nullptr) return PN2_ERR_INVALID_ARG;
  if (!aligned4(temp)) return PN2_ERR_ALIGNMENT;
  fps_streaming_kernel<<<b, 1024, 0, s>>>(g, dataset, temp, idxs);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
