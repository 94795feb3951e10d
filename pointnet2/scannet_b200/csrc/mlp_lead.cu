// First layer of a shared-MLP block applied BEFORE the gather (training mode), sm_100a.
//
// PointnetSAModule(MSG) / PointnetFPModule feed their SharedMLP with GATHERED rows
// (/root/reference/pointnet2/pointnet2_modules.py:60-75: QueryAndGroup -> Conv2d 1x1 ...;
//  :395-412: three_interpolate -> cat -> Conv2d 1x1 ...).  A 1x1 convolution is linear in its input and
// the gather / the interpolation only select and weight rows, so the two commute:
//
//   SA:  z[g, s, :] = W [x_j - c_g | f_j]      = U[j, :] + W_x (x_j - c_g),      U = F W_f^T   (j = idx[g, s])
//   FP:  z[r, :]    = W [sum_t w_t k_{i_t} | s_r] = sum_t w_t G[i_t, :] + W_s s_r,   G = K W_k^T
//
// U / G are products over the N (or m) SOURCE points -- 4-8x fewer rows than the gathered positions -- on the
// tcgen05 layer kernel (mlp_gemm.cu); the kernels here produce the first layer's pre-BatchNorm output z
// straight from them.  The gathered rows tensor (B*npoint*nsample x (3 + C), B*n x (C2 + C1)) is never
// written or re-read, forward or backward; the coordinate term stays in fp32 on the centred differences
// exactly as the reference forms them (pointnet2_utils.py:348-350), so it loses nothing to TF32.
//
//   pn2_sa_lead_fwd   z = U[idx] + W_x (xyz[idx] - centre)          + BatchNorm statistics partials of z
//   pn2_fp_lead_fwd   z = ((w1 G[i1]) + (w2 G[i2])) + (w3 G[i3]) + W_s skip   + the same partials
//   pn2_sa_lead_bwd   dZ = a (relu'(.) dH - c1 - zhat c2) formed on the fly from (dH, z);
//                     dU[idx] += dZ (128-bit vector reductions), dW_x partials = sum dZ (x_j - c_g)^T
//   pn2_fp_lead_bwd   dG[i_t] += w_t dZ, dW_s partials = sum dZ skip^T
//
// after which dW_f = dU^T F, dF = dU W_f (and dW_k = dG^T K, dK = dG W_k) are products over the source points
// again (mlp_wgrad.cu, mlp_gemm.cu).  Statistics partials are [blocks][2][C] (column sums of z and z^2),
// one row per CTA over a fixed row range, summed in fixed order by pn2_bn_finalize: schedule-independent.
#include <cstdlib>

#include "pn2_common.cuh"

namespace pn2 {
namespace lead {

constexpr int kThreads = 256;
// Grid: up to 8 CTAs per SM's worth of row ranges (the kernels keep 3-4 resident per SM).  Finer than one
// resident wave on purpose: these kernels share the GPU with the next batch's furthest-point sampling (32 SMs
// held for ~0.7 ms), and a one-wave grid of statically assigned ranges doubled its time whenever a fifth of its
// CTAs had to wait for a free SM (318 vs 186 us measured); short ranges let the hardware scheduler balance.
// The forward kernels' partials are BatchNorm statistics that a one-warp-per-channel finalise sums (its time grows
// with their number: 14 us at 1184 partials against 6 at 296) and they never run beside the sampling kernel: ONE
// resident wave (4 / 3 CTAs per SM; FP0 forward measured 75 us with 444 CTAs, 100 us with 592).
enum Kind { kSaFwd = 0, kSaBwd = 1, kFpFwd = 2, kFpBwd = 3 };
static inline int blocks_per_sm(int kind) { return kind == kSaFwd ? 4 : kind == kFpFwd ? 3 : kind == kSaBwd ? 6 : 8; }
constexpr int kMaxSkip = 8;   // skip channels handled in registers by the FP kernels

struct Geom {
  int P, C, C4, row_lanes, rows_per_cta;
};

// a thread owns one float4 column group and a row lane; a CTA a fixed range of rows
static Geom geom(int P, int C, int kind, int *nblk) {
  Geom g;
  g.P = P, g.C = C, g.C4 = C / 4;
  g.row_lanes = kThreads / g.C4 > 0 ? kThreads / g.C4 : 1;
  int nb = (P + g.row_lanes * 8 - 1) / (g.row_lanes * 8);
  if (nb > blocks_per_sm(kind) * kNumSMs) nb = blocks_per_sm(kind) * kNumSMs;
  if (nb < 1) nb = 1;
  g.rows_per_cta = (P + nb - 1) / nb;
  *nblk = (P + g.rows_per_cta - 1) / g.rows_per_cta;
  return g;
}

struct FastDiv {
  unsigned d, mul, shr;
};
static FastDiv fast_div(unsigned d) {
  FastDiv f{d, 0u, 0u};
  if (d <= 1) return f;
  unsigned l = 0;
  while ((1ull << l) < d) ++l;
  f.mul = static_cast<unsigned>(((1ull << (31 + l)) + d - 1) / d);
  f.shr = l - 1;
  return f;
}
__device__ __forceinline__ unsigned fdiv(unsigned x, const FastDiv &f) {
  return f.d <= 1 ? x : (__umulhi(x, f.mul) >> f.shr);
}

__device__ __forceinline__ float4 ld4(const float *p, size_t i4) {
  return __ldg(reinterpret_cast<const float4 *>(p) + i4);
}
__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ void fma4(float4 &acc, const float4 &v, float s) {
  acc.x = fmaf(v.x, s, acc.x), acc.y = fmaf(v.y, s, acc.y), acc.z = fmaf(v.z, s, acc.z), acc.w = fmaf(v.w, s, acc.w);
}

// Sum of NV float4 accumulators over the CTA's row lanes in fixed order -> partials[blockIdx.x][NV][C].
// Every thread of the CTA calls (threads past row_lanes * C4 only take part in the barrier).
template <int NV>
__device__ __forceinline__ void cta_reduce(const Geom &g, int lane, int c4, float4 (&acc)[NV],
                                           float *__restrict__ partials) {
  extern __shared__ float4 s_red[];  // [row_lanes][NV][C4]
  if (lane < g.row_lanes) {
#pragma unroll
    for (int v = 0; v < NV; ++v) s_red[(lane * NV + v) * g.C4 + c4] = acc[v];
  }
  __syncthreads();
  // one thread per (vector, column group) adds the row lanes in ascending order (fixed order: schedule-independent)
  float4 *dst = reinterpret_cast<float4 *>(partials + static_cast<size_t>(blockIdx.x) * NV * g.C);
  for (int e = threadIdx.x; e < NV * g.C4; e += kThreads) {
    const int v = e / g.C4, cc = e - v * g.C4;
    float4 t = s_red[v * g.C4 + cc];
    for (int l = 1; l < g.row_lanes; ++l) {
      const float4 u = s_red[(l * NV + v) * g.C4 + cc];
      t.x += u.x, t.y += u.y, t.z += u.z, t.w += u.w;
    }
    dst[v * g.C4 + cc] = t;
  }
}

// dZ of a hidden Conv + BatchNorm(train) + ReLU layer for one float4:
//   a (relu'(a z + b) dH - c1 - zhat c2) = a dY + (p + q z),  q = -a invstd c2,  p = -a c1 - q mean
struct Adjoint {
  float4 a, b, p, q;
  int relu;
  __device__ __forceinline__ void load(int c4, int C4, int relu_, const float *pa, const float *pb, const float *pm,
                                       const float *pi, const float *c1c2) {
    a = ld4(pa, c4), b = ld4(pb, c4);
    const float4 mu = ld4(pm, c4), is = ld4(pi, c4), k1 = ld4(c1c2, c4), k2 = ld4(c1c2, C4 + c4);
    q = make_float4(-a.x * is.x * k2.x, -a.y * is.y * k2.y, -a.z * is.z * k2.z, -a.w * is.w * k2.w);
    p = make_float4(-a.x * k1.x - q.x * mu.x, -a.y * k1.y - q.y * mu.y, -a.z * k1.z - q.z * mu.z, -a.w * k1.w - q.w * mu.w);
    relu = relu_;
  }
  __device__ __forceinline__ float one(float d, float v, float aa, float bb, float pp, float qq) const {
    const float dy = (!relu || fmaf(aa, v, bb) > 0.f) ? d : 0.f;
    return fmaf(aa, dy, fmaf(qq, v, pp));
  }
  __device__ __forceinline__ float4 dz(const float4 &d, const float4 &v) const {
    return make_float4(one(d.x, v.x, a.x, b.x, p.x, q.x), one(d.y, v.y, a.y, b.y, p.y, q.y),
                       one(d.z, v.z, a.z, b.z, p.z, q.z), one(d.w, v.w, a.w, b.w, p.w, q.w));
  }
};

// Every kernel below walks its CTA's row range in steps of U rows per thread; a thread owns one float4 column
// group c4 of its rows.  The C/4 threads of a row need the same per-row scalars (neighbour index, centred
// coordinates; interpolation indices / weights, skip channels).  They sit in ONE register per row: lane t of the
// row's group of G = min(C/4, 32) lanes loads scalar t, and the value is handed round by warp shuffles where it
// is used -- one load instruction per row instead of 7-14 redundant ones, and no register arrays of scalars
// (the first version of these kernels spilled its accumulators; ncu: 70-80 % long_scoreboard stalls).  All loads
// of a step are issued before the first use, and the one load a gather address depends on -- the neighbour
// index -- is fetched one step AHEAD, so a step costs one memory round trip.  C is a power of two (>= 16), so
// a CTA's 256 threads tile rows exactly and every lane takes part in every shuffle.

__device__ __forceinline__ float shfl_f(float v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }
__device__ __forceinline__ int shfl_i(int v, int src, int width) { return __shfl_sync(0xffffffffu, v, src, width); }

// ------------------------------------------------------------------------------ SA, forward
// rows are (cloud, centre, sample): r / ns = the centre's index into new_xyz, r / (npoint * ns) = the cloud
template <int U>
__global__ void __launch_bounds__(kThreads, 4)
sa_lead_fwd_kernel(Geom g, FastDiv np_ns, FastDiv ns, int n, const float *__restrict__ u,
                   const float *__restrict__ xyz, int ldx, const float *__restrict__ new_xyz, const int *__restrict__ idx,
                   const float *__restrict__ wx, int ldw, float *__restrict__ z, float *__restrict__ partials) {
  pdl_enter();
  const int c4 = threadIdx.x % g.C4, lane = threadIdx.x / g.C4;
  const int G = min(g.C4, 32), sub = threadIdx.x & (G - 1);
  float4 acc[2];
  acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  float w[4][3];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int k = 0; k < 3; ++k) w[j][k] = __ldg(wx + static_cast<size_t>(4 * c4 + j) * ldw + k);
  const int r0 = blockIdx.x * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.P);
  const int step = g.row_lanes * U;
  int k[U];   // neighbour index of the step's rows: valid in lane 0 of the row's group
#pragma unroll
  for (int i = 0; i < U; ++i) k[i] = sub == 0 ? __ldg(idx + min(r0 + lane + i * g.row_lanes, r1 - 1)) : 0;
  for (int rb = r0; rb < r1; rb += step) {   // (uniform trip count over the CTA: the shuffles need every lane)
    const int r = rb + lane;
    float4 uv[U];
    float pc[U], cc[U];   // lanes 0..2 of the group: x / y / z of the neighbour and of the centre
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int rr = min(r + i * g.row_lanes, r1 - 1);
      const size_t src = static_cast<size_t>(fdiv(rr, np_ns)) * n + shfl_i(k[i], 0, G);
      uv[i] = ld4(u, src * g.C4 + c4);
      pc[i] = cc[i] = 0.f;
      if (sub < 3) {
        pc[i] = __ldg(xyz + src * ldx + sub);
        cc[i] = __ldg(new_xyz + static_cast<size_t>(fdiv(rr, ns)) * 3 + sub);
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i)   // next step's indices
      k[i] = sub == 0 ? __ldg(idx + min(r + step + i * g.row_lanes, r1 - 1)) : 0;
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const float dd = pc[i] - cc[i];
      const float dx = shfl_f(dd, 0, G), dy = shfl_f(dd, 1, G), dz = shfl_f(dd, 2, G);
      if (r + i * g.row_lanes < r1) {
        float4 v = uv[i];
        v.x += fmaf(w[0][2], dz, fmaf(w[0][1], dy, w[0][0] * dx));
        v.y += fmaf(w[1][2], dz, fmaf(w[1][1], dy, w[1][0] * dx));
        v.z += fmaf(w[2][2], dz, fmaf(w[2][1], dy, w[2][0] * dx));
        v.w += fmaf(w[3][2], dz, fmaf(w[3][1], dy, w[3][0] * dx));
        reinterpret_cast<float4 *>(z)[static_cast<size_t>(r + i * g.row_lanes) * g.C4 + c4] = v;
        acc[0].x += v.x, acc[0].y += v.y, acc[0].z += v.z, acc[0].w += v.w;
        acc[1].x = fmaf(v.x, v.x, acc[1].x), acc[1].y = fmaf(v.y, v.y, acc[1].y);
        acc[1].z = fmaf(v.z, v.z, acc[1].z), acc[1].w = fmaf(v.w, v.w, acc[1].w);
      }
    }
  }
  cta_reduce<2>(g, lane, c4, acc, partials);
}

// ------------------------------------------------------------------------------ SA, backward
template <int U>
__global__ void __launch_bounds__(kThreads, 3)
sa_lead_bwd_kernel(Geom g, FastDiv np_ns, FastDiv ns, int n, int relu, const float *__restrict__ dh,
                   const float *__restrict__ z, const float *__restrict__ a, const float *__restrict__ b,
                   const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ c1c2,
                   const float *__restrict__ xyz, int ldx, const float *__restrict__ new_xyz,
                   const int *__restrict__ idx, float *__restrict__ du, float *__restrict__ partials) {
  pdl_enter();
  const int c4 = threadIdx.x % g.C4, lane = threadIdx.x / g.C4;
  const int G = min(g.C4, 32), sub = threadIdx.x & (G - 1);
  float4 acc[3];
  acc[0] = acc[1] = acc[2] = make_float4(0.f, 0.f, 0.f, 0.f);
  Adjoint adj;
  adj.load(c4, g.C4, relu, a, b, mean, invstd, c1c2);
  const int r0 = blockIdx.x * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.P);
  const int step = g.row_lanes * U;
  int k[U];
#pragma unroll
  for (int i = 0; i < U; ++i) k[i] = sub == 0 ? __ldg(idx + min(r0 + lane + i * g.row_lanes, r1 - 1)) : 0;
  for (int rb = r0; rb < r1; rb += step) {
    const int r = rb + lane;
    unsigned src[U];   // b * n < 2^31 (checked by the launcher)
    float4 d[U], v[U];
    float pc[U], cc[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const int rr = min(r + i * g.row_lanes, r1 - 1);
      d[i] = ld4(dh, static_cast<size_t>(rr) * g.C4 + c4);
      v[i] = ld4(z, static_cast<size_t>(rr) * g.C4 + c4);
      src[i] = fdiv(rr, np_ns) * n + shfl_i(k[i], 0, G);
      pc[i] = cc[i] = 0.f;
      if (sub < 3) {
        pc[i] = __ldg(xyz + static_cast<size_t>(src[i]) * ldx + sub);
        cc[i] = __ldg(new_xyz + static_cast<size_t>(fdiv(rr, ns)) * 3 + sub);
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i)   // next step's indices
      k[i] = sub == 0 ? __ldg(idx + min(r + step + i * g.row_lanes, r1 - 1)) : 0;
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const float dd = pc[i] - cc[i];
      const float dx = shfl_f(dd, 0, G), dy = shfl_f(dd, 1, G), dzz = shfl_f(dd, 2, G);
      if (r + i * g.row_lanes < r1) {
        const float4 gz = adj.dz(d[i], v[i]);
        red_add_v4(du + (static_cast<size_t>(src[i]) * g.C4 + c4) * 4, gz);
        fma4(acc[0], gz, dx), fma4(acc[1], gz, dy), fma4(acc[2], gz, dzz);
      }
    }
  }
  cta_reduce<3>(g, lane, c4, acc, partials);
}

// ------------------------------------------------------------------------------ FP
// the 6 + c1 scalars of row r: lane t of the row's group holds  t < 3: idx[r, t] (bits) | t < 6: weight[r, t - 3]
// | t < 6 + c1: skip[r, t - 6]   (G >= 16 >= 6 + kMaxSkip lanes per group)
__device__ __forceinline__ float fp_row_scalar(int sub, int c1, int rr, const int *__restrict__ idx,
                                               const float *__restrict__ weight, const float *__restrict__ skip,
                                               int lds) {
  const size_t q = static_cast<size_t>(rr) * 3;
  if (sub < 3) return __int_as_float(__ldg(idx + q + sub));
  if (sub < 6) return __ldg(weight + q + (sub - 3));
  if (sub < 6 + c1) return __ldg(skip + static_cast<size_t>(rr) * lds + (sub - 6));
  return 0.f;
}

// (forward: measured faster with the weights / skip channels loaded by every thread -- uniform addresses, L1
// broadcast, no dependent load -- than with the shuffle hand-round the backward kernel uses: 75 vs 97 us at FP0)
template <int U>
__global__ void __launch_bounds__(kThreads, 3)
fp_lead_fwd_kernel(Geom g, FastDiv nn, int m, int c1, const float *__restrict__ gk, const float *__restrict__ skip,
                   int lds, const int *__restrict__ idx, const float *__restrict__ weight, const float *__restrict__ ws,
                   int ldw, float *__restrict__ z, float *__restrict__ partials) {
  pdl_enter();
  const int c4 = threadIdx.x % g.C4, lane = threadIdx.x / g.C4;
  float4 acc[2];
  acc[0] = acc[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  // W_s as [c1][C4] float4 (w[k][c4] = W_s[4 c4 .. 4 c4 + 3, k]) in shared memory, behind cta_reduce's area
  extern __shared__ float4 s_all[];
  float4 *s_w = s_all + 2 * g.row_lanes * g.C4;
  for (int e = threadIdx.x; e < c1 * g.C4; e += kThreads) {
    const int k = e / g.C4, cc = e - k * g.C4;
    const float *p = ws + static_cast<size_t>(4 * cc) * ldw + k;
    s_w[e] = make_float4(__ldg(p), __ldg(p + ldw), __ldg(p + 2 * ldw), __ldg(p + 3 * ldw));
  }
  __syncthreads();
  const int r0 = blockIdx.x * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.P);
  const int step = g.row_lanes * U;
  int it[U][3];
#pragma unroll
  for (int i = 0; i < U; ++i) {
    const size_t q = static_cast<size_t>(min(r0 + lane + i * g.row_lanes, r1 - 1)) * 3;
#pragma unroll
    for (int t = 0; t < 3; ++t) it[i][t] = __ldg(idx + q + t);
  }
  for (int r = r0 + lane; r < r1; r += step) {
    int rr[U];
    float wt[U][3], sk[U][kMaxSkip];
    float4 k0[U], k1[U], k2[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {
      rr[i] = min(r + i * g.row_lanes, r1 - 1);
      const size_t row0 = static_cast<size_t>(fdiv(rr[i], nn)) * m;
      k0[i] = ld4(gk, (row0 + it[i][0]) * g.C4 + c4);
      k1[i] = ld4(gk, (row0 + it[i][1]) * g.C4 + c4);
      k2[i] = ld4(gk, (row0 + it[i][2]) * g.C4 + c4);
      const size_t q = static_cast<size_t>(rr[i]) * 3;
#pragma unroll
      for (int t = 0; t < 3; ++t) wt[i][t] = __ldg(weight + q + t);
      const float *sr = skip + static_cast<size_t>(rr[i]) * lds;
#pragma unroll
      for (int k = 0; k < kMaxSkip; ++k) sk[i][k] = k < c1 ? __ldg(sr + k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {   // next step's neighbour indices
      const size_t q = static_cast<size_t>(min(r + step + i * g.row_lanes, r1 - 1)) * 3;
#pragma unroll
      for (int t = 0; t < 3; ++t) it[i][t] = __ldg(idx + q + t);
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      if (r + i * g.row_lanes >= r1) break;
      float4 v;
      v.x = fmaf(k2[i].x, wt[i][2], fmaf(k1[i].x, wt[i][1], k0[i].x * wt[i][0]));
      v.y = fmaf(k2[i].y, wt[i][2], fmaf(k1[i].y, wt[i][1], k0[i].y * wt[i][0]));
      v.z = fmaf(k2[i].z, wt[i][2], fmaf(k1[i].z, wt[i][1], k0[i].z * wt[i][0]));
      v.w = fmaf(k2[i].w, wt[i][2], fmaf(k1[i].w, wt[i][1], k0[i].w * wt[i][0]));
#pragma unroll
      for (int k = 0; k < kMaxSkip; ++k)
        if (k < c1) fma4(v, s_w[k * g.C4 + c4], sk[i][k]);
      reinterpret_cast<float4 *>(z)[static_cast<size_t>(rr[i]) * g.C4 + c4] = v;
      acc[0].x += v.x, acc[0].y += v.y, acc[0].z += v.z, acc[0].w += v.w;
      acc[1].x = fmaf(v.x, v.x, acc[1].x), acc[1].y = fmaf(v.y, v.y, acc[1].y);
      acc[1].z = fmaf(v.z, v.z, acc[1].z), acc[1].w = fmaf(v.w, v.w, acc[1].w);
    }
  }
  cta_reduce<2>(g, lane, c4, acc, partials);
}

template <int U>
__global__ void __launch_bounds__(kThreads, 3)
fp_lead_bwd_kernel(Geom g, FastDiv nn, int m, int c1, int relu, const float *__restrict__ dh,
                   const float *__restrict__ z, const float *__restrict__ a, const float *__restrict__ b,
                   const float *__restrict__ mean, const float *__restrict__ invstd, const float *__restrict__ c1c2,
                   const float *__restrict__ skip, int lds, const int *__restrict__ idx,
                   const float *__restrict__ weight, float *__restrict__ dg, float *__restrict__ partials) {
  pdl_enter();
  const int c4 = threadIdx.x % g.C4, lane = threadIdx.x / g.C4;
  const int G = min(g.C4, 32), sub = threadIdx.x & (G - 1);
  float4 acc[kMaxSkip];
#pragma unroll
  for (int k = 0; k < kMaxSkip; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
  Adjoint adj;
  adj.load(c4, g.C4, relu, a, b, mean, invstd, c1c2);
  const int r0 = blockIdx.x * g.rows_per_cta, r1 = min(r0 + g.rows_per_cta, g.P);
  for (int rb = r0; rb < r1; rb += g.row_lanes * U) {
    const int r = rb + lane;
    float sc[U];
    float4 d[U], v[U];
    size_t row0[U];
#pragma unroll
    for (int i = 0; i < U; ++i) {   // nothing here depends on another load: one round trip per step
      const int rr = min(r + i * g.row_lanes, r1 - 1);
      d[i] = ld4(dh, static_cast<size_t>(rr) * g.C4 + c4);
      v[i] = ld4(z, static_cast<size_t>(rr) * g.C4 + c4);
      sc[i] = fp_row_scalar(sub, c1, rr, idx, weight, skip, lds);
      row0[i] = static_cast<size_t>(fdiv(rr, nn)) * m;
    }
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const bool ok = r + i * g.row_lanes < r1;
      const float4 gz = adj.dz(d[i], v[i]);
      const int si = __float_as_int(sc[i]);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        const int it = shfl_i(si, t, G);
        const float wt = shfl_f(sc[i], 3 + t, G);
        if (ok)
          red_add_v4(dg + ((row0[i] + it) * g.C4 + c4) * 4, make_float4(gz.x * wt, gz.y * wt, gz.z * wt, gz.w * wt));
      }
#pragma unroll
      for (int k = 0; k < kMaxSkip; ++k)
        if (k < c1) {
          const float sk = shfl_f(sc[i], 6 + k, G);
          if (ok) fma4(acc[k], gz, sk);
        }
    }
  }
  cta_reduce<kMaxSkip>(g, lane, c4, acc, partials);
}

// widths: powers of two, 16 .. 1024 (a CTA's 256 threads then tile rows exactly: every lane takes part in the shuffles)
static inline bool shape_ok(long long p, int c) { return p > 0 && p < (1ll << 31) && c >= 16 && c <= 1024 && (c & (c - 1)) == 0; }

}  // namespace lead
}  // namespace pn2

using namespace pn2;

extern "C" int pn2_lead_blocks(int p, int c, int kind) {
  if (!lead::shape_ok(p, c) || kind < 0 || kind > 3) return 0;
  int nblk;
  lead::geom(p, c, kind, &nblk);
  return nblk;
}
extern "C" int pn2_fp_lead_max_skip(void) { return lead::kMaxSkip; }

extern "C" int pn2_sa_lead_fwd(int b, int n, int npoint, int nsample, int c, const float *u, const float *xyz,
                               int ldx, const float *new_xyz, const int *idx, const float *wx, int ldw, float *z,
                               float *partials, void *stream) {
  const long long p = static_cast<long long>(b) * npoint * nsample;
  if (b <= 0 || n <= 0 || npoint <= 0 || nsample <= 0 || ldw < 3 || ldx < 3 || !lead::shape_ok(p, c))
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(u, xyz, new_xyz, idx, wx, z, partials);
  if (!aligned16(u) || !aligned16(z) || !aligned16(partials)) return PN2_ERR_UNSUPPORTED;
  int nblk;
  const lead::Geom g = lead::geom(static_cast<int>(p), c, lead::kSaFwd, &nblk);
  const size_t smem = sizeof(float4) * 2 * g.row_lanes * g.C4;
  pn2::launch(lead::sa_lead_fwd_kernel<4>, nblk, lead::kThreads, smem, static_cast<cudaStream_t>(stream), g, lead::fast_div(npoint * nsample), lead::fast_div(nsample), n, u, xyz, ldx, new_xyz, idx, wx, ldw, z, partials);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_sa_lead_bwd(int b, int n, int npoint, int nsample, int c, int relu, const float *dh,
                               const float *z, const float *a, const float *bb, const float *mean,
                               const float *invstd, const float *c1c2, const float *xyz, int ldx,
                               const float *new_xyz, const int *idx, float *du, int zero_fill, float *partials,
                               void *stream) {
  const long long p = static_cast<long long>(b) * npoint * nsample;
  if (b <= 0 || n <= 0 || npoint <= 0 || nsample <= 0 || ldx < 3 || !lead::shape_ok(p, c) ||
      static_cast<long long>(b) * n >= (1ll << 31))
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dh, z, a, bb, mean, invstd, c1c2, xyz, new_xyz, idx, du, partials);
  if (!aligned16(dh) || !aligned16(z) || !aligned16(a) || !aligned16(bb) || !aligned16(mean) || !aligned16(invstd) ||
      !aligned16(c1c2) || !aligned16(du) || !aligned16(partials))
    return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (zero_fill) {   // (0: the caller zeroed du ahead of time, off the critical path)
    cudaError_t e = cudaMemsetAsync(du, 0, sizeof(float) * static_cast<size_t>(b) * n * c, s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int nblk;
  const lead::Geom g = lead::geom(static_cast<int>(p), c, lead::kSaBwd, &nblk);
  const size_t smem = sizeof(float4) * 3 * g.row_lanes * g.C4;
  pn2::launch(lead::sa_lead_bwd_kernel<4>, nblk, lead::kThreads, smem, s, g, lead::fast_div(npoint * nsample), lead::fast_div(nsample), n, relu, dh, z, a, bb, mean, invstd, c1c2, xyz,
      ldx, new_xyz, idx, du, partials);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_fp_lead_fwd(int b, int n, int m, int c, int c1, const float *gk, const float *skip, int lds,
                               const int *idx, const float *weight, const float *ws, int ldw, float *z,
                               float *partials, void *stream) {
  const long long p = static_cast<long long>(b) * n;
  if (b <= 0 || n <= 0 || m <= 0 || c1 < 0 || c1 > lead::kMaxSkip || lds < c1 || c < 64 || !lead::shape_ok(p, c))
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(gk, idx, weight, z, partials);
  if (c1 > 0) PN2_CHECK_PTRS(skip, ws);
  if (!aligned16(gk) || !aligned16(z) || !aligned16(partials)) return PN2_ERR_UNSUPPORTED;
  int nblk;
  const lead::Geom g = lead::geom(static_cast<int>(p), c, lead::kFpFwd, &nblk);
  const size_t smem = sizeof(float4) * (2 * g.row_lanes * g.C4 + c1 * g.C4);
  pn2::launch(lead::fp_lead_fwd_kernel<2>, nblk, lead::kThreads, smem, static_cast<cudaStream_t>(stream), g, lead::fast_div(n), m, c1, gk, skip, lds, idx, weight, ws, ldw, z, partials);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_fp_lead_bwd(int b, int n, int m, int c, int c1, int relu, const float *dh, const float *z,
                               const float *a, const float *bb, const float *mean, const float *invstd,
                               const float *c1c2, const float *skip, int lds, const int *idx, const float *weight,
                               float *dg, int zero_fill, float *partials, void *stream) {
  const long long p = static_cast<long long>(b) * n;
  if (b <= 0 || n <= 0 || m <= 0 || c1 < 0 || c1 > lead::kMaxSkip || lds < c1 || c < 64 || !lead::shape_ok(p, c))
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dh, z, a, bb, mean, invstd, c1c2, idx, weight, dg, partials);
  if (c1 > 0) PN2_CHECK_PTRS(skip);
  if (!aligned16(dh) || !aligned16(z) || !aligned16(a) || !aligned16(bb) || !aligned16(mean) || !aligned16(invstd) ||
      !aligned16(c1c2) || !aligned16(dg) || !aligned16(partials))
    return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  if (zero_fill) {
    cudaError_t e = cudaMemsetAsync(dg, 0, sizeof(float) * static_cast<size_t>(b) * m * c, s);
    if (e != cudaSuccess) return static_cast<int>(e);
  }
  int nblk;
  const lead::Geom g = lead::geom(static_cast<int>(p), c, lead::kFpBwd, &nblk);
  const size_t smem = sizeof(float4) * lead::kMaxSkip * g.row_lanes * g.C4;
  pn2::launch(lead::fp_lead_bwd_kernel<2>, nblk, lead::kThreads, smem, s, g, lead::fast_div(n), m, c1, relu, dh, z, a, bb, mean, invstd, c1c2, skip, lds, idx, weight, dg, partials);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
