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
#include <cooperative_groups.h>
#include <math.h>
#include <stdlib.h>

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

namespace cg = cooperative_groups;

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
// pruned register-resident kernel (2048 < positions <= 8192)
//
// Same arithmetic, same answer, fewer updates.  Points are Morton-sorted inside the kernel
// (bitonic sort in shared memory, ~7 us) so that every WARP owns 256 spatially compact points
// and their bounding box.  After a new centre c is picked, a warp first evaluates the lower
// bound  lb = ((ex*ex)+(ey*ey))+(ez*ez),  e = max(lo - c, c - hi, 0) per axis, with the very
// operations the distance uses.  fp32 rounding is monotone, so lb <= d(p, c) for EVERY point p of
// the box, bit for bit; if lb >= the warp's current maximum of the running minima, no minimum
// in the warp can change (min(d, t) = t for all of them) and the warp's arg-max candidate from the
// previous iteration is still exact: the warp skips the iteration.  Once a few dozen samples
// are placed the running minima are small and a new centre only touches nearby warps.
//
// The tie-break can no longer be encoded in the register order (slots are spatial now): a thread
// resolves equal minima by comparing the reference positions pos(k) kept in shared memory --
// rare, except in the first iterations when every minimum is still 1e10.
// ---------------------------------------------------------------------------
__device__ __forceinline__ unsigned morton_expand10(unsigned v) {  // 10 bits -> every third bit
  v = (v | (v << 16)) & 0x030000FFu;
  v = (v | (v << 8)) & 0x0300F00Fu;
  v = (v | (v << 4)) & 0x030C30C3u;
  v = (v | (v << 2)) & 0x09249249u;
  return v;
}

#ifdef PN2_FPS_DEBUG
__device__ unsigned long long g_fps_active[1024];  // per-iteration count of non-skipping warps
#endif

template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_pruned_kernel(FpsGeom g, const float *__restrict__ dataset, int *__restrict__ idxs) {
  constexpr int M = T * P;          // sorted slots (power of two)
  constexpr int NW = T / 32;
  extern __shared__ __align__(16) unsigned char fps_smem[];
  float4 *s_pts = reinterpret_cast<float4 *>(fps_smem);                              // [M] by reference position
  unsigned long long *s_keys = reinterpret_cast<unsigned long long *>(s_pts + M);    // [M] (morton << 32) | position
  float4 *s_box = reinterpret_cast<float4 *>(s_keys + M);                            // [NW][2] lo, hi
  __shared__ int2 s_slots[2][32];
  __shared__ float s_ext[6][32];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int npos = g.bs * g.q;
  const float *pts = dataset + static_cast<size_t>(blockIdx.x) * g.n * 3;
  int *out = idxs + static_cast<size_t>(blockIdx.x) * g.m;

  // ---- 1. load by reference position, cloud bounding box over the live points
  float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
  unsigned state[P];  // 0 live, 1 real but skipped by the reference (never a candidate), 2 padding
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int p = tid * P + i;
    const int k = p < npos ? fps_pos_to_index(p, g) : g.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    state[i] = 2;
    if (k < g.n) {
      px = __ldg(pts + 3 * k), py = __ldg(pts + 3 * k + 1), pz = __ldg(pts + 3 * k + 2);
      state[i] = (static_cast<double>(sqnorm_rn(px, py, pz)) <= 1e-3) ? 1 : 0;
      if (state[i] == 0) {
        lo[0] = fminf(lo[0], px), lo[1] = fminf(lo[1], py), lo[2] = fminf(lo[2], pz);
        hi[0] = fmaxf(hi[0], px), hi[1] = fmaxf(hi[1], py), hi[2] = fmaxf(hi[2], pz);
      }
    }
    s_pts[p] = make_float4(px, py, pz, __int_as_float(k < g.n ? k : 0));
  }
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    for (int o = 16; o > 0; o >>= 1) {
      lo[a] = fminf(lo[a], __shfl_xor_sync(kFull, lo[a], o));
      hi[a] = fmaxf(hi[a], __shfl_xor_sync(kFull, hi[a], o));
    }
    if (lane == 0) s_ext[a][warp] = lo[a], s_ext[3 + a][warp] = hi[a];
  }
  __syncthreads();
#pragma unroll
  for (int a = 0; a < 3; ++a) {
    float l = lane < NW ? s_ext[a][lane] : 3.0e38f, h = lane < NW ? s_ext[3 + a][lane] : -3.0e38f;
    for (int o = 16; o > 0; o >>= 1) {
      l = fminf(l, __shfl_xor_sync(kFull, l, o));
      h = fmaxf(h, __shfl_xor_sync(kFull, h, o));
    }
    lo[a] = l, hi[a] = h;
  }
  // ---- 2. sort keys: Morton code of the live points; skipped and padding slots go last
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const int p = tid * P + i;
    unsigned code = state[i] == 0 ? 0u : (state[i] == 1 ? 0xFFFFFFFEu : 0xFFFFFFFFu);
    if (state[i] == 0) {
      const float4 v = s_pts[p];
      const float q[3] = {v.x, v.y, v.z};
      unsigned c3[3];
#pragma unroll
      for (int a = 0; a < 3; ++a) {
        const float span = hi[a] - lo[a];
        const float u = span > 0.f ? __fdividef(q[a] - lo[a], span) : 0.f;  // sort key only: no parity role
        c3[a] = static_cast<unsigned>(fminf(fmaxf(u * 1023.f, 0.f), 1023.f));
      }
      code = morton_expand10(c3[0]) | (morton_expand10(c3[1]) << 1) | (morton_expand10(c3[2]) << 2);
    }
    s_keys[p] = (static_cast<unsigned long long>(code) << 32) |
                static_cast<unsigned>(state[i] == 2 ? 0x7fffffff : p);
  }
  __syncthreads();
  for (int size = 2; size <= M; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      for (int e = tid; e < M / 2; e += T) {
        const int i0 = 2 * e - (e & (stride - 1));  // lower element of the pair
        const int i1 = i0 + stride;
        const bool up = (i0 & size) == 0;
        const unsigned long long a = s_keys[i0], b = s_keys[i1];
        if ((a > b) == up) s_keys[i0] = b, s_keys[i1] = a;
      }
      __syncthreads();
    }
  }
  // ---- 3. this thread's P sorted slots -> registers (pairs), warp bounding box
  constexpr int H = P / 2;
  float2 x[H], y[H], z[H], t[H];
  float wl[3] = {3.0e38f, 3.0e38f, 3.0e38f}, wh[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
#pragma unroll
  for (int i = 0; i < P; ++i) {
    const unsigned long long key = s_keys[tid * P + i];
    const unsigned code = static_cast<unsigned>(key >> 32), pos = static_cast<unsigned>(key);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (pos < static_cast<unsigned>(npos)) v = s_pts[pos];
    const bool live = code < 0xFFFFFFFEu;
    if (live) {
      wl[0] = fminf(wl[0], v.x), wl[1] = fminf(wl[1], v.y), wl[2] = fminf(wl[2], v.z);
      wh[0] = fmaxf(wh[0], v.x), wh[1] = fmaxf(wh[1], v.y), wh[2] = fmaxf(wh[2], v.z);
    }
    const float t0 = live ? 1e10f : -1.0f;
    if (i & 1) x[i / 2].y = v.x, y[i / 2].y = v.y, z[i / 2].y = v.z, t[i / 2].y = t0;
    else x[i / 2].x = v.x, y[i / 2].x = v.y, z[i / 2].x = v.z, t[i / 2].x = t0;
  }
#pragma unroll
  for (int a = 0; a < 3; ++a)
    for (int o = 16; o > 0; o >>= 1) {
      wl[a] = fminf(wl[a], __shfl_xor_sync(kFull, wl[a], o));
      wh[a] = fmaxf(wh[a], __shfl_xor_sync(kFull, wh[a], o));
    }
  if (lane == 0) {
    s_box[2 * warp] = make_float4(wl[0], wl[1], wl[2], 0.f);
    s_box[2 * warp + 1] = make_float4(wh[0], wh[1], wh[2], 0.f);
  }
  if (tid == 0) out[0] = 0;
  __syncthreads();

  const unsigned *s_pos = reinterpret_cast<const unsigned *>(s_keys);  // low word of key e: 2*e
  auto pos_of = [&](int i) { return s_pos[2 * (tid * P + i)]; };
  int wval = __float_as_int(1e10f);  // the warp's cached arg-max candidate: (value bits, position)
  unsigned wpos = 0xffffffffu;
  bool fresh = false;                // becomes true once the warp has evaluated an iteration
  float4 c = s_pts[0];               // position 0 <-> index 0
  for (int j = 1; j < g.m; ++j) {
    // exact lower bound of the distance from c to any point of this warp's box
    const float4 bl = s_box[2 * warp], bh = s_box[2 * warp + 1];
    const float ex = fmaxf(fmaxf(__fsub_rn(bl.x, c.x), __fsub_rn(c.x, bh.x)), 0.f);
    const float ey = fmaxf(fmaxf(__fsub_rn(bl.y, c.y), __fsub_rn(c.y, bh.y)), 0.f);
    const float ez = fmaxf(fmaxf(__fsub_rn(bl.z, c.z), __fsub_rn(c.z, bh.z)), 0.f);
    const float lb = __fadd_rn(__fadd_rn(__fmul_rn(ex, ex), __fmul_rn(ey, ey)), __fmul_rn(ez, ez));
    if (!(fresh && lb >= __int_as_float(wval))) {
#ifdef PN2_FPS_DEBUG
      if (lane == 0 && j < 1024) atomicAdd(&g_fps_active[j], 1ull);
#endif
      // Only a few warps survive the bounding-box test, so what matters here is the length of the
      // dependency chain, not the instruction count: all P updates are independent, the maximum
      // is a tree, and the tie-break positions are fetched only when a thread really has a tie.
      const float2 cx = make_float2(c.x, c.x), cy = make_float2(c.y, c.y), cz = make_float2(c.z, c.z);
#pragma unroll
      for (int h = 0; h < H; ++h) {
        const float2 d = sqdist2_rn(x[h], y[h], z[h], cx, cy, cz);
        t[h].x = fminf(d.x, t[h].x);
        t[h].y = fminf(d.y, t[h].y);
      }
      float pm[H];
#pragma unroll
      for (int h = 0; h < H; ++h) pm[h] = fmaxf(t[h].x, t[h].y);
#pragma unroll
      for (int w = H; w > 1; w = (w + 1) / 2)
#pragma unroll
        for (int h = 0; h < w / 2; ++h) pm[h] = fmaxf(pm[h], pm[w - 1 - h]);
      const float best = pm[0];
      int bi = 0, ties = 0;
#pragma unroll
      for (int i = P - 1; i >= 0; --i) {  // lowest slot holding the maximum, and how many hold it
        const bool e = ((i & 1) ? t[i / 2].y : t[i / 2].x) == best;
        bi = e ? i : bi;
        ties += e ? 1 : 0;
      }
      if (ties > 1) {  // equal minima inside one thread: the reference position decides
        unsigned bp = 0xffffffffu;
#pragma unroll
        for (int i = 0; i < P; ++i)
          if (((i & 1) ? t[i / 2].y : t[i / 2].x) == best) {
            const unsigned q = pos_of(i);
            if (q < bp) bp = q, bi = i;
          }
      }
      const int v = __float_as_int(best);
      wval = __reduce_max_sync(kFull, v);
      wpos = __reduce_min_sync(kFull, v == wval ? pos_of(bi) : 0xffffffffu);
      fresh = true;
    }
    if (lane == 0) s_slots[j & 1][warp] = make_int2(wval, static_cast<int>(wpos));
    __syncthreads();
    int2 r = make_int2(INT_MIN, -1);
    if (lane < NW) r = s_slots[j & 1][lane];
    const int gmax = __reduce_max_sync(kFull, r.x);
    const unsigned win = __reduce_min_sync(kFull, r.x == gmax ? static_cast<unsigned>(r.y) : 0xffffffffu);
    c = s_pts[win < static_cast<unsigned>(npos) ? win : 0];
    if (tid == 0) out[j] = __float_as_int(c.w);
  }
}

template <int P, int T>
static cudaError_t launch_pruned(const FpsGeom &g, int b, const float *dataset, int *idxs,
                                 cudaStream_t stream) {
  const size_t smem = static_cast<size_t>(T) * P * (sizeof(float4) + sizeof(unsigned long long)) +
                      sizeof(float4) * 2 * (T / 32);
  auto kern = fps_pruned_kernel<P, T>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  kern<<<b, T, smem, stream>>>(g, dataset, idxs);
  count_launch();
  return cudaGetLastError();
}

// ---------------------------------------------------------------------------
// thread-block-cluster kernel: one cloud per CLUSTER of C CTAs (8192 < positions <= 16 * T * P,
// i.e. whole scenes of up to ~160k points, or smaller clouds split for latency).
//
// CTA r keeps positions [r*S, (r+1)*S), S = T*P, in registers exactly like the resident kernel.
// Per iteration: local block arg-max (one __syncthreads), then lane r of warp 0 writes this CTA's
// candidate -- (value bits, position, x, y, z, original index), 32 bytes -- into slot [parity][rank]
// of CTA r's shared memory (distributed shared memory store), ONE cluster barrier, and every warp
// of every CTA reduces the C candidates with the same REDUX pair.  The exchanged record carries the
// winner's coordinates, so no CTA ever reads another CTA's point table.  Slots are double buffered
// by iteration parity, which makes one cluster barrier per iteration sufficient (a CTA can only run
// one barrier ahead of the slowest reader).
// ---------------------------------------------------------------------------
struct __align__(16) FpsCand {
  int val;        // float bits of the running minimum (as int: -1.0f sorts below every value >= 0)
  unsigned pos;   // reference position (tie-break key)
  int idx;        // original point index
  int pad;
  float x, y, z, w;
};

template <int P, int T>
__global__ void __launch_bounds__(T, 1)
fps_cluster_kernel(FpsGeom g, int csize, const float *__restrict__ dataset, int *__restrict__ idxs) {
  constexpr int S = T * P;
  extern __shared__ float4 s_pts[];  // [S] this CTA's slice (x, y, z, bits(index)), by local position
  __shared__ int2 s_slots[2][32];
  __shared__ FpsCand s_xchg[2][16];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int cloud = blockIdx.x / csize;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nwarps = T / 32;
  const int npos = g.bs * g.q;
  const float *pts = dataset + static_cast<size_t>(cloud) * g.n * 3;
  int *out = idxs + static_cast<size_t>(cloud) * g.m;
  const int base = rank * S;

  constexpr int H = (P + 1) / 2;
  float2 x[H], y[H], z[H], t[H];
#pragma unroll
  for (int i = 0; i < 2 * H; ++i) {
    const int p = base + tid * P + i;
    const int k = (i < P && p < npos) ? fps_pos_to_index(p, g) : g.n;
    float px = 0.f, py = 0.f, pz = 0.f;
    bool live = false;
    if (k < g.n) {
      px = __ldg(pts + 3 * k), py = __ldg(pts + 3 * k + 1), pz = __ldg(pts + 3 * k + 2);
      live = !(static_cast<double>(sqnorm_rn(px, py, pz)) <= 1e-3);
    }
    const float t0 = live ? 1e10f : -1.0f;
    if (i & 1) x[i / 2].y = px, y[i / 2].y = py, z[i / 2].y = pz, t[i / 2].y = t0;
    else x[i / 2].x = px, y[i / 2].x = py, z[i / 2].x = pz, t[i / 2].x = t0;
    if (i < P) s_pts[tid * P + i] = make_float4(px, py, pz, __int_as_float(k < g.n ? k : 0));
  }
  if (rank == 0 && tid == 0) out[0] = 0;
  // first centre = index 0 = position 0, owned by CTA 0: fetch it from global memory
  float cx0 = __ldg(pts + 0), cy0 = __ldg(pts + 1), cz0 = __ldg(pts + 2);
  cluster.sync();  // every CTA's shared memory (exchange slots) exists before anyone writes to it

  float c_x = cx0, c_y = cy0, c_z = cz0;
  for (int j = 1; j < g.m; ++j) {
    const float2 cx = make_float2(c_x, c_x), cy = make_float2(c_y, c_y), cz = make_float2(c_z, c_z);
    float best;
    int bi = 0;
#pragma unroll
    for (int h = 0; h < H; ++h) {
      const float2 d = sqdist2_rn(x[h], y[h], z[h], cx, cy, cz);
      t[h].x = fminf(d.x, t[h].x);
      t[h].y = fminf(d.y, t[h].y);
      if (h == 0) {
        best = t[0].x;
      } else if (t[h].x > best) {
        best = t[h].x, bi = 2 * h;
      }
      if (2 * h + 1 < P && t[h].y > best) best = t[h].y, bi = 2 * h + 1;
    }
    // local arg-max (value, lowest global position)
    const int v = __float_as_int(best);
    const int wmax = __reduce_max_sync(kFull, v);
    const unsigned wpos = __reduce_min_sync(kFull, v == wmax ? static_cast<unsigned>(base + tid * P + bi) : 0xffffffffu);
    if (lane == 0) s_slots[j & 1][warp] = make_int2(wmax, static_cast<int>(wpos));
    __syncthreads();
    if (warp == 0) {
      int2 r = make_int2(INT_MIN, -1);
      if (lane < nwarps) r = s_slots[j & 1][lane];
      const int lmax = __reduce_max_sync(kFull, r.x);
      const unsigned lpos = __reduce_min_sync(kFull, r.x == lmax ? static_cast<unsigned>(r.y) : 0xffffffffu);
      if (lane < csize) {  // lane r publishes this CTA's candidate in CTA r's exchange slot
        const float4 pw = s_pts[lpos - base];
        FpsCand cand;
        cand.val = lmax, cand.pos = lpos, cand.idx = __float_as_int(pw.w), cand.pad = 0;
        cand.x = pw.x, cand.y = pw.y, cand.z = pw.z, cand.w = 0.f;
        FpsCand *remote = cluster.map_shared_rank(&s_xchg[j & 1][rank], lane);
        *remote = cand;
      }
    }
    cluster.sync();  // release the stores / acquire everyone else's
    int cv = INT_MIN;
    unsigned cp = 0xffffffffu;
    if (lane < csize) cv = s_xchg[j & 1][lane].val, cp = s_xchg[j & 1][lane].pos;
    const int gmax = __reduce_max_sync(kFull, cv);
    const unsigned gpos = __reduce_min_sync(kFull, cv == gmax ? cp : 0xffffffffu);
    const unsigned who = __ballot_sync(kFull, cv == gmax && cp == gpos);
    const FpsCand &win = s_xchg[j & 1][__ffs(who) - 1];
    c_x = win.x, c_y = win.y, c_z = win.z;
    if (rank == 0 && tid == 0) out[j] = win.idx;
  }
  cluster.sync();  // no CTA may exit while a peer can still write into its shared memory
}

template <int P, int T>
static cudaError_t launch_cluster(const FpsGeom &g, int b, int csize, const float *dataset, int *idxs,
                                  cudaStream_t stream) {
  auto kern = fps_cluster_kernel<P, T>;
  const size_t smem = static_cast<size_t>(T) * P * sizeof(float4);
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return e;
  if (csize > 8) {
    e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) return e;
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(b * csize), cfg.blockDim = dim3(T), cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = csize, attr[0].val.clusterDim.y = 1, attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr, cfg.numAttrs = 1;
  e = cudaLaunchKernelEx(&cfg, kern, g, csize, dataset, idxs);
  count_launch();
  return e != cudaSuccess ? e : cudaGetLastError();
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

#ifdef PN2_FPS_DEBUG
extern "C" __attribute__((visibility("default"))) int pn2_debug_fps_active(unsigned long long *host1024) {
  cudaDeviceSynchronize();
  cudaError_t e = cudaMemcpyFromSymbol(host1024, pn2::g_fps_active, sizeof(unsigned long long) * 1024);
  unsigned long long zeros[1024] = {0};
  cudaMemcpyToSymbol(pn2::g_fps_active, zeros, sizeof(zeros));
  return static_cast<int>(e);
}
#endif

static int fps_dispatch(int b, int n, int m, const float *dataset, float *temp, int *idxs, void *stream,
                        bool force_pruned) {
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
    // Measured on B200 (8192 -> 1024, B = 32): the bounding-box variant leaves only 19% of the warps
    // active per iteration yet runs 708 us against 635 us for the plain resident kernel -- at this
    // size the iteration is bound by what EVERY warp executes (winner fetch, box test, slot,
    // barrier, second-stage arg-max), not by the distance updates.  It is therefore opt-in
    // (pn2_furthest_point_sampling_pruned / PN2_FPS_PRUNE=1); see DESIGN.md section 4.
    static const bool prune = getenv("PN2_FPS_PRUNE") != nullptr;
    // PN2_FPS_SPLIT=c (2, 4, 8): split every resident-size cloud over a cluster of c CTAs (more SMs
    // per cloud, fewer points per thread, one cluster barrier per iteration extra).  Measured, not
    // the default: see DESIGN.md section 4.
    const int split = getenv("PN2_FPS_SPLIT") != nullptr ? atoi(getenv("PN2_FPS_SPLIT")) : 0;
    if (split >= 2 && split <= 8 && !force_pruned && npos > 1024 && m > 1) {
      const int per = (npos + split - 1) / split;
      if (per <= 8 * 128) return static_cast<int>(launch_cluster<8, 128>(g, b, split, dataset, idxs, s));
      if (per <= 8 * 256) return static_cast<int>(launch_cluster<8, 256>(g, b, split, dataset, idxs, s));
      if (per <= 8 * 512) return static_cast<int>(launch_cluster<8, 512>(g, b, split, dataset, idxs, s));
    }
    if ((prune || force_pruned) && npos > 4096 && m > 2) e = launch_pruned<8, 1024>(g, b, dataset, idxs, s);
    else if ((prune || force_pruned) && npos > 2048 && m > 2) e = launch_pruned<8, 512>(g, b, dataset, idxs, s);
    else if (npos <= 128 * 1) e = launch_resident<1, 1024>(g, b, threads_for(1), dataset, idxs, s);
    else if (npos <= 128 * 2) e = launch_resident<2, 1024>(g, b, threads_for(2), dataset, idxs, s);
    else if (npos <= 128 * 4) e = launch_resident<4, 1024>(g, b, threads_for(4), dataset, idxs, s);
    else if (npos <= 1024 * 8) e = launch_resident<8, 1024>(g, b, threads_for(8), dataset, idxs, s);
    else return PN2_ERR_UNSUPPORTED;
    return static_cast<int>(e);
  }
  // whole scenes: one CLUSTER per cloud, the slices register-resident in 2..16 CTAs
  const bool no_cluster = getenv("PN2_FPS_NO_CLUSTER") != nullptr;
  if (!no_cluster && npos <= 16 * 768 * 13) {
    int csize = 2;
    while (csize < 16 && csize * 8192 < npos) csize *= 2;
    cudaError_t e = csize * 8192 >= npos ? launch_cluster<8, 1024>(g, b, csize, dataset, idxs, s)
                                         : launch_cluster<13, 768>(g, b, 16, dataset, idxs, s);
    return static_cast<int>(e);
  }
  if (temp == nullptr) return PN2_ERR_INVALID_ARG;
  if (!aligned4(temp)) return PN2_ERR_ALIGNMENT;
  fps_streaming_kernel<<<b, 1024, 0, s>>>(g, dataset, temp, idxs);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_furthest_point_sampling(int b, int n, int m, const float *dataset, float *temp,
                                           int *idxs, void *stream) {
  return fps_dispatch(b, n, m, dataset, temp, idxs, stream, false);
}

// Same result through the bounding-box-pruned kernel where it applies (2048 < n <= 8192).
extern "C" int pn2_furthest_point_sampling_pruned(int b, int n, int m, const float *dataset, float *temp,
                                                  int *idxs, void *stream) {
  return fps_dispatch(b, n, m, dataset, temp, idxs, stream, true);
}
