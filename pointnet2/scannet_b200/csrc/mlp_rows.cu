// Shared-MLP block in POSITION-MAJOR layout for sm_100a: everything around the 1x1
// convolutions of PointnetSAModule(MSG) / PointnetFPModule
// (/root/reference/pointnet2/pointnet2_modules.py:60-75, 395-416;
//  /root/reference/pointnet2/pytorch_utils.py:11-36,87-120).
//
// The reference materialises (B, C, npoint, nsample) tensors and runs cuDNN's NCHW
// BatchNorm / ReLU / max_pool2d over them: on B200 55% of its training step is
// cudnn::bn_fw_tr / bn_bw (profiles/r01_step_breakdown_baseline.md).  Here an activation
// is a row-major [P, C] matrix -- P = B*npoint*nsample grouped positions (SA) or B*n
// points (FP), channels contiguous -- so a 1x1 conv is a plain GEMM Z = H * W^T and
// every kernel below streams rows with 128-bit accesses:
//
//   sa_group_rows        X0[p,:] = table[b, idx[p], :] - (centre, 0..)   (QueryAndGroup)
//   bn_stats             per-channel batch mean / biased variance of Z (two-stage,
//                        deterministic: per-CTA partials, then one fp64 finalise that also
//                        folds BN into  y = a*z + b  and updates the running statistics)
//   bn_relu_apply        H = relu(a*Z + b)
//   bn_relu_groupmax     out[g,c] = max_s relu(a*Z[g*ns+s,c] + b), first arg-max kept
//   pool_bn_bwd          adjoint of bn_relu_groupmax through train-mode BN -> dense dZ
//   relu_bn_bwd          adjoint of bn_relu_apply through train-mode BN    -> dense dZ
//   sa_group_rows_grad   scatter-add of dX0 into the point table (128-bit vector reds)
//
// Train-mode BN adjoint per channel (P rows, zhat = (z-mean)*invstd, a = gamma*invstd):
//   dbeta = sum dY,  dgamma = sum dY*zhat,  dZ = a * (dY - dbeta/P - zhat*dgamma/P)
#include <cstdlib>

#include "pn2_common.cuh"

namespace pn2 {

constexpr int kRowThreads = 256;
constexpr int kMaxPartials = 4 * kNumSMs;  // CTAs of a column reduction (4 per SM)

// ---------------------------------------------------------------------------------
// generic per-channel reduction over the rows of [P, C]: every thread owns one float4
// column group and a row lane; a CTA reduces its slab to one partial per channel
// ---------------------------------------------------------------------------------
struct ColGeom {
  int P, C, C4, row_lanes, rows_per_cta;
};

static ColGeom col_geom(int P, int C, int *nblk) {
  ColGeom g;
  g.P = P, g.C = C, g.C4 = C / 4;
  g.row_lanes = kRowThreads / g.C4 > 0 ? kRowThreads / g.C4 : 1;
  int nb = (P + g.row_lanes * 8 - 1) / (g.row_lanes * 8);
  if (nb > kMaxPartials) nb = kMaxPartials;
  if (nb < 1) nb = 1;
  g.rows_per_cta = (P + nb - 1) / nb;
  *nblk = (P + g.rows_per_cta - 1) / g.rows_per_cta;
  return g;
}

// F(row, col4, acc0, acc1): accumulates two float4 statistics for one (row, column group)
template <class F>
__device__ __forceinline__ void col_reduce2(const ColGeom &g, float *__restrict__ partials, F f) {
  extern __shared__ float4 s_red[];  // [row_lanes][2][C4]
  const int r0 = blockIdx.x * g.rows_per_cta;
  const int r1 = min(r0 + g.rows_per_cta, g.P);
  for (int c4 = threadIdx.x % g.C4; c4 < g.C4; c4 += kRowThreads) {  // C4 > 256 only if C > 1024
    const int lane = threadIdx.x / g.C4;
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    if (lane < g.row_lanes) {
      // four rows per trip: their loads are independent of each other and of the accumulators, so the compiler
      // issues all of them before the first add (same order of additions as the plain loop); one row per trip
      // left a single global-load round trip in flight per thread
      int r = r0 + lane;
      for (; r + 3 * g.row_lanes < r1; r += 4 * g.row_lanes) {
        float4 a0 = make_float4(0.f, 0.f, 0.f, 0.f), a1 = a0, b0 = a0, b1 = a0, c0 = a0, c1 = a0, d0 = a0, d1 = a0;
        f(r, c4, a0, a1), f(r + g.row_lanes, c4, b0, b1), f(r + 2 * g.row_lanes, c4, c0, c1), f(r + 3 * g.row_lanes, c4, d0, d1);
        s0.x += a0.x, s0.y += a0.y, s0.z += a0.z, s0.w += a0.w, s1.x += a1.x, s1.y += a1.y, s1.z += a1.z, s1.w += a1.w;
        s0.x += b0.x, s0.y += b0.y, s0.z += b0.z, s0.w += b0.w, s1.x += b1.x, s1.y += b1.y, s1.z += b1.z, s1.w += b1.w;
        s0.x += c0.x, s0.y += c0.y, s0.z += c0.z, s0.w += c0.w, s1.x += c1.x, s1.y += c1.y, s1.z += c1.z, s1.w += c1.w;
        s0.x += d0.x, s0.y += d0.y, s0.z += d0.z, s0.w += d0.w, s1.x += d1.x, s1.y += d1.y, s1.z += d1.z, s1.w += d1.w;
      }
      for (; r < r1; r += g.row_lanes) f(r, c4, s0, s1);
    }
    if (lane < g.row_lanes) {
      s_red[(lane * 2 + 0) * g.C4 + c4] = s0;
      s_red[(lane * 2 + 1) * g.C4 + c4] = s1;
    }
    __syncthreads();
    if (lane == 0) {
      for (int l = 1; l < g.row_lanes; ++l) {
        const float4 a = s_red[(l * 2 + 0) * g.C4 + c4], b = s_red[(l * 2 + 1) * g.C4 + c4];
        s0.x += a.x, s0.y += a.y, s0.z += a.z, s0.w += a.w;
        s1.x += b.x, s1.y += b.y, s1.z += b.z, s1.w += b.w;
      }
      float4 *dst = reinterpret_cast<float4 *>(partials + static_cast<size_t>(blockIdx.x) * 2 * g.C);
      dst[c4] = s0;
      dst[g.C4 + c4] = s1;
    }
    __syncthreads();
  }
}

__device__ __forceinline__ float4 ld4(const float *p, size_t i4) {
  return __ldg(reinterpret_cast<const float4 *>(p) + i4);
}

// ------------------------------------------------------------------ forward statistics
__global__ void __launch_bounds__(kRowThreads)
bn_stats_partial_kernel(ColGeom g, const float *__restrict__ z, float *__restrict__ partials) {
  pdl_enter();
  col_reduce2(g, partials, [&](int r, int c4, float4 &s, float4 &q) {
    const float4 v = ld4(z, static_cast<size_t>(r) * g.C4 + c4);
    s.x += v.x, s.y += v.y, s.z += v.z, s.w += v.w;
    q.x += v.x * v.x, q.y += v.y * v.y, q.z += v.z * v.z, q.w += v.w * v.w;
  });
}

// Fixed-order fp64 sum of the per-CTA partials of one channel by one warp: lane l adds
// partials l, l+32, ... and a shuffle tree combines the lanes (same order every run).
__device__ __forceinline__ void warp_sum_partials(int ld, int c, int nblk, const float *__restrict__ partials,
                                                  double &s, double &q) {
  const int lane = threadIdx.x & 31;
  s = 0.0, q = 0.0;
  // partials: [nblk][2][ld].  All of the lane's loads are issued before the first add (same order of
  // additions as a plain loop, hence the same bits): a dependent loop costs one L2 round trip per
  // iteration, ~10 of them for the ~300 partials of a tensor-core layer.
  constexpr int kBatch = 10;
  for (int k0 = lane; k0 < nblk; k0 += 32 * kBatch) {
    float vs[kBatch], vq[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i) {
      const int k = k0 + 32 * i;
      const bool ok = k < nblk;
      vs[i] = ok ? __ldg(partials + static_cast<size_t>(k) * 2 * ld + c) : 0.f;
      vq[i] = ok ? __ldg(partials + static_cast<size_t>(k) * 2 * ld + ld + c) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      if (k0 + 32 * i < nblk) s += vs[i], q += vq[i];
  }
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    q += __shfl_xor_sync(0xffffffffu, q, o);
  }
}

constexpr int kFinalizeWarps = 8;  // channels per CTA of the finalise kernels

// one WARP per channel, then the BN fold
__global__ void __launch_bounds__(32 * kFinalizeWarps)
bn_stats_finalize_kernel(int P, int C, int ld, int nblk, const float *__restrict__ partials,
                         const float *__restrict__ gamma, const float *__restrict__ beta,
                         float eps, float momentum, float *running_mean,
                         float *running_var, long long *num_batches_tracked, float *__restrict__ a,
                         float *__restrict__ b, float *__restrict__ mean_out,
                         float *__restrict__ invstd_out) {
  pdl_enter();
  const int c = blockIdx.x * kFinalizeWarps + (threadIdx.x >> 5);
  if (blockIdx.x == 0 && threadIdx.x == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (c >= C) return;
  double s, q;
  warp_sum_partials(ld, c, nblk, partials, s, q);
  if ((threadIdx.x & 31) != 0) return;
  // fp64 only where cancellation needs it (q/P - mean^2), and no fp64 division or square root:
  // Blackwell's fp64 rate makes a handful of DP divides the cost of this whole kernel.
  const double inv_p = static_cast<double>(1.0f / static_cast<float>(P)) ;
  const double ip = inv_p * (2.0 - static_cast<double>(P) * inv_p);  // one Newton step: 1/P to fp64 accuracy
  const double mean = s * ip;
  double var = q * ip - mean * mean;  // biased, as BatchNorm normalises with
  if (var < 0.0) var = 0.0;
  const float varf = static_cast<float>(var);
  const float invstd = 1.0f / sqrtf(varf + eps);
  const float ga = gamma[c] * invstd;
  a[c] = ga;
  b[c] = beta[c] - static_cast<float>(mean) * ga;
  mean_out[c] = static_cast<float>(mean);
  invstd_out[c] = invstd;
  if (running_mean != nullptr) {  // torch: running = (1-m)*running + m*stat, unbiased variance
    const float unbiased = P > 1 ? varf * (static_cast<float>(P) / static_cast<float>(P - 1)) : varf;
    running_mean[c] = (1.f - momentum) * running_mean[c] + momentum * static_cast<float>(mean);
    running_var[c] = (1.f - momentum) * running_var[c] + momentum * unbiased;
  }
}

// eval mode: fold the running statistics
__global__ void bn_fold_eval_kernel(int C, const float *__restrict__ gamma, const float *__restrict__ beta,
                                    const float *__restrict__ running_mean,
                                    const float *__restrict__ running_var, float eps,
                                    float *__restrict__ a, float *__restrict__ b,
                                    float *__restrict__ mean_out, float *__restrict__ invstd_out) {
  pdl_enter();
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const float invstd = 1.0f / sqrtf(running_var[c] + eps);
  const float ga = gamma[c] * invstd;
  a[c] = ga;
  b[c] = beta[c] - running_mean[c] * ga;
  mean_out[c] = running_mean[c];
  invstd_out[c] = invstd;
}

// -------------------------------------------------------------------------- apply
__global__ void __launch_bounds__(kRowThreads)
bn_relu_apply_kernel(size_t total4, int C4, const float *__restrict__ z, const float *__restrict__ a,
                     const float *__restrict__ b, float *__restrict__ h, int relu) {
  pdl_enter();
  const size_t stride = static_cast<size_t>(gridDim.x) * kRowThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x; i < total4; i += stride) {
    const int c4 = static_cast<int>(i % C4);
    const float4 v = ld4(z, i), aa = ld4(a, c4), bb = ld4(b, c4);
    float4 y = make_float4(fmaf(aa.x, v.x, bb.x), fmaf(aa.y, v.y, bb.y), fmaf(aa.z, v.z, bb.z),
                           fmaf(aa.w, v.w, bb.w));
    if (relu) y = make_float4(fmaxf(y.x, 0.f), fmaxf(y.y, 0.f), fmaxf(y.z, 0.f), fmaxf(y.w, 0.f));
    reinterpret_cast<float4 *>(h)[i] = y;
  }
}

// out[g, c] = max over the ns rows of group g of relu(a*z + b); first maximal row kept
__global__ void __launch_bounds__(kRowThreads)
bn_relu_groupmax_kernel(int G, int ns, int C4, const float *__restrict__ z, const float *__restrict__ a,
                        const float *__restrict__ b, float *__restrict__ out, int *__restrict__ argmax) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= static_cast<size_t>(G) * C4) return;
  const int c4 = static_cast<int>(i % C4);
  const size_t g = i / C4;
  const float4 aa = ld4(a, c4), bb = ld4(b, c4);
  float4 best = make_float4(-1.f, -1.f, -1.f, -1.f);  // relu output is >= 0
  int4 arg = make_int4(0, 0, 0, 0);
  for (int s = 0; s < ns; ++s) {
    const float4 v = ld4(z, (g * ns + s) * C4 + c4);
    const float y0 = fmaxf(fmaf(aa.x, v.x, bb.x), 0.f), y1 = fmaxf(fmaf(aa.y, v.y, bb.y), 0.f);
    const float y2 = fmaxf(fmaf(aa.z, v.z, bb.z), 0.f), y3 = fmaxf(fmaf(aa.w, v.w, bb.w), 0.f);
    if (y0 > best.x) best.x = y0, arg.x = s;
    if (y1 > best.y) best.y = y1, arg.y = s;
    if (y2 > best.z) best.z = y2, arg.z = s;
    if (y3 > best.w) best.w = y3, arg.w = s;
  }
  reinterpret_cast<float4 *>(out)[i] = best;
  reinterpret_cast<int4 *>(argmax)[i] = arg;
}

// out[g, c] and its arg-max row from the group extrema of Z that the tensor-core layer left
// (csrc/mlp_gemm.cu, pool_ns): max_s relu(a*z_s + b) = relu(a * (a > 0 ? max z : min z) + b), a
// monotone map of z; the row is the first one attaining that extremum -- row 0 when the result
// is not positive, as bn_relu_groupmax_kernel's strict comparison leaves it.
__global__ void __launch_bounds__(kRowThreads)
groupmax_from_extrema_kernel(size_t total, int C, const float *__restrict__ zmax, const float *__restrict__ zmin,
                             const int *__restrict__ imax, const int *__restrict__ imin,
                             const float *__restrict__ a, const float *__restrict__ b, float *__restrict__ out,
                             int *__restrict__ argmax, float *__restrict__ zsel) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= total) return;
  const int c = static_cast<int>(i % C);
  const float aa = __ldg(a + c), bb = __ldg(b + c);
  const bool up = aa > 0.f;
  const float zs = up ? __ldg(zmax + i) : __ldg(zmin + i);
  const float y = fmaf(aa, zs, bb);
  const bool pos = y > 0.f && aa != 0.f;
  if (zsel != nullptr) zsel[i] = zs;   // = z at the arg-max row whenever that row carries a gradient (a != 0)
  out[i] = fmaxf(y, 0.f);
  argmax[i] = pos ? (up ? __ldg(imax + i) : __ldg(imin + i)) : 0;
}

// ------------------------------------------------------------- adjoint of the pooled layer
// sparse statistics: only the arg-max row of every (group, channel) carries a gradient
__global__ void __launch_bounds__(kRowThreads)
pool_bwd_stats_kernel(ColGeom g, int ns, const float *__restrict__ dout, const int *__restrict__ argmax,
                      const float *__restrict__ z, const float *__restrict__ zsel, const float *__restrict__ a,
                      const float *__restrict__ b, const float *__restrict__ mean,
                      const float *__restrict__ invstd, float *__restrict__ partials) {
  pdl_enter();
  // rows here are GROUPS (g.P = G)
  col_reduce2(g, partials, [&](int grp, int c4, float4 &s, float4 &q) {
    const size_t i = static_cast<size_t>(grp) * g.C4 + c4;
    const float4 d = ld4(dout, i);
    const int4 am = __ldg(reinterpret_cast<const int4 *>(argmax) + i);
    const float4 aa = ld4(a, c4), bb = ld4(b, c4), mu = ld4(mean, c4), is = ld4(invstd, c4);
    float z0, z1, z2, z3;
    if (zsel != nullptr && aa.x != 0.f && aa.y != 0.f && aa.z != 0.f && aa.w != 0.f) {
      // the forward pass kept z at the arg-max row (pn2_groupmax_from_extrema): no gather over z.  Where the
      // group's maximum is not positive the kept value differs from z at row 0, but that entry is masked.
      const float4 zs = ld4(zsel, i);
      z0 = zs.x, z1 = zs.y, z2 = zs.z, z3 = zs.w;
    } else {
      const size_t base = static_cast<size_t>(grp) * ns;
      z0 = __ldg(z + (base + am.x) * g.C + 4 * c4 + 0);
      z1 = __ldg(z + (base + am.y) * g.C + 4 * c4 + 1);
      z2 = __ldg(z + (base + am.z) * g.C + 4 * c4 + 2);
      z3 = __ldg(z + (base + am.w) * g.C + 4 * c4 + 3);
    }
    const float d0 = fmaf(aa.x, z0, bb.x) > 0.f ? d.x : 0.f, d1 = fmaf(aa.y, z1, bb.y) > 0.f ? d.y : 0.f;
    const float d2 = fmaf(aa.z, z2, bb.z) > 0.f ? d.z : 0.f, d3 = fmaf(aa.w, z3, bb.w) > 0.f ? d.w : 0.f;
    s.x += d0, s.y += d1, s.z += d2, s.w += d3;
    q.x += d0 * (z0 - mu.x) * is.x, q.y += d1 * (z1 - mu.y) * is.y;
    q.z += d2 * (z2 - mu.z) * is.z, q.w += d3 * (z3 - mu.w) * is.w;
  });
}

// sums the partials (fp64, fixed order, one warp per channel): dbeta, dgamma, and the two
// per-channel constants c1 = dbeta/P, c2 = dgamma/P used by the dense dZ kernels (0 in eval mode)
__global__ void __launch_bounds__(32 * kFinalizeWarps)
bwd_stats_finalize_kernel(int P, int C, int ld, int nblk, int training, const float *__restrict__ partials,
                          float *__restrict__ dgamma, float *__restrict__ dbeta, float *__restrict__ c1,
                          float *__restrict__ c2) {
  pdl_enter();
  const int c = blockIdx.x * kFinalizeWarps + (threadIdx.x >> 5);
  if (c >= C) return;
  double s, q;
  warp_sum_partials(ld, c, nblk, partials, s, q);
  if ((threadIdx.x & 31) != 0) return;
  dbeta[c] = static_cast<float>(s);
  dgamma[c] = static_cast<float>(q);
  const float inv_p = 1.0f / static_cast<float>(P);  // no fp64 division (see bn_stats_finalize_kernel)
  c1[c] = training ? static_cast<float>(s) * inv_p : 0.f;
  c2[c] = training ? static_cast<float>(q) * inv_p : 0.f;
}

// dense dZ of the pooled layer: dZ[p,c] = a*(dY - c1 - zhat*c2), dY nonzero at the arg-max row
__global__ void __launch_bounds__(kRowThreads)
pool_bwd_dz_kernel(int G, int ns, int C4, const float *__restrict__ dout, const int *__restrict__ argmax,
                   const float *__restrict__ z, const float *__restrict__ a, const float *__restrict__ b,
                   const float *__restrict__ mean, const float *__restrict__ invstd,
                   const float *__restrict__ c1, const float *__restrict__ c2, float *__restrict__ dz) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= static_cast<size_t>(G) * C4) return;
  const int c4 = static_cast<int>(i % C4);
  const size_t g = i / C4;
  const float4 d = ld4(dout, i);
  const int4 am = __ldg(reinterpret_cast<const int4 *>(argmax) + i);
  const float4 aa = ld4(a, c4), bb = ld4(b, c4), mu = ld4(mean, c4), is = ld4(invstd, c4);
  const float4 k1 = ld4(c1, c4), k2 = ld4(c2, c4);
  for (int s = 0; s < ns; ++s) {
    const size_t r = (g * ns + s) * C4 + c4;
    const float4 v = ld4(z, r);
    float4 o;
    {
      const float dy = (s == am.x && fmaf(aa.x, v.x, bb.x) > 0.f) ? d.x : 0.f;
      o.x = aa.x * (dy - k1.x - (v.x - mu.x) * is.x * k2.x);
    }
    {
      const float dy = (s == am.y && fmaf(aa.y, v.y, bb.y) > 0.f) ? d.y : 0.f;
      o.y = aa.y * (dy - k1.y - (v.y - mu.y) * is.y * k2.y);
    }
    {
      const float dy = (s == am.z && fmaf(aa.z, v.z, bb.z) > 0.f) ? d.z : 0.f;
      o.z = aa.z * (dy - k1.z - (v.z - mu.z) * is.z * k2.z);
    }
    {
      const float dy = (s == am.w && fmaf(aa.w, v.w, bb.w) > 0.f) ? d.w : 0.f;
      o.w = aa.w * (dy - k1.w - (v.w - mu.w) * is.w * k2.w);
    }
    reinterpret_cast<float4 *>(dz)[r] = o;
  }
}

// ------------------------------------------------------- adjoint of a hidden BN+ReLU layer
__global__ void __launch_bounds__(kRowThreads)
relu_bwd_stats_kernel(ColGeom g, int relu, const float *__restrict__ dh, const float *__restrict__ z,
                      const float *__restrict__ a, const float *__restrict__ b,
                      const float *__restrict__ mean, const float *__restrict__ invstd,
                      float *__restrict__ partials) {
  pdl_enter();
  col_reduce2(g, partials, [&](int r, int c4, float4 &s, float4 &q) {
    const size_t i = static_cast<size_t>(r) * g.C4 + c4;
    const float4 d = ld4(dh, i), v = ld4(z, i);
    const float4 aa = ld4(a, c4), bb = ld4(b, c4), mu = ld4(mean, c4), is = ld4(invstd, c4);
    const float d0 = (!relu || fmaf(aa.x, v.x, bb.x) > 0.f) ? d.x : 0.f;
    const float d1 = (!relu || fmaf(aa.y, v.y, bb.y) > 0.f) ? d.y : 0.f;
    const float d2 = (!relu || fmaf(aa.z, v.z, bb.z) > 0.f) ? d.z : 0.f;
    const float d3 = (!relu || fmaf(aa.w, v.w, bb.w) > 0.f) ? d.w : 0.f;
    s.x += d0, s.y += d1, s.z += d2, s.w += d3;
    q.x += d0 * (v.x - mu.x) * is.x, q.y += d1 * (v.y - mu.y) * is.y;
    q.z += d2 * (v.z - mu.z) * is.z, q.w += d3 * (v.w - mu.w) * is.w;
  });
}

__global__ void __launch_bounds__(kRowThreads)
relu_bwd_dz_kernel(size_t total4, int C4, int relu, const float *__restrict__ dh, const float *__restrict__ z,
                   const float *__restrict__ a, const float *__restrict__ b, const float *__restrict__ mean,
                   const float *__restrict__ invstd, const float *__restrict__ c1,
                   const float *__restrict__ c2, float *__restrict__ dz) {
  pdl_enter();
  const size_t stride = static_cast<size_t>(gridDim.x) * kRowThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x; i < total4; i += stride) {
    const int c4 = static_cast<int>(i % C4);
    const float4 d = ld4(dh, i), v = ld4(z, i);
    const float4 aa = ld4(a, c4), bb = ld4(b, c4), mu = ld4(mean, c4), is = ld4(invstd, c4);
    const float4 k1 = ld4(c1, c4), k2 = ld4(c2, c4);
    float4 o;
    o.x = aa.x * (((!relu || fmaf(aa.x, v.x, bb.x) > 0.f) ? d.x : 0.f) - k1.x - (v.x - mu.x) * is.x * k2.x);
    o.y = aa.y * (((!relu || fmaf(aa.y, v.y, bb.y) > 0.f) ? d.y : 0.f) - k1.y - (v.y - mu.y) * is.y * k2.y);
    o.z = aa.z * (((!relu || fmaf(aa.z, v.z, bb.z) > 0.f) ? d.z : 0.f) - k1.z - (v.z - mu.z) * is.z * k2.z);
    o.w = aa.w * (((!relu || fmaf(aa.w, v.w, bb.w) > 0.f) ? d.w : 0.f) - k1.w - (v.w - mu.w) * is.w * k2.w);
    reinterpret_cast<float4 *>(dz)[i] = o;
  }
}

// --------------------------------------------------------------- grouping in row layout
// rows[(b*np + j)*ns + s, :] = table[b, idx[b,j,s], :]; the first 3 columns (xyz) are centred
// on new_xyz[b, j] (pointnet2_utils.py:348-350).  One thread per (row, float4 column group).
__global__ void __launch_bounds__(kRowThreads)
sa_group_rows_kernel(size_t total4, int n, int np_ns, int ns, int ct4, const float *__restrict__ table,
                     const float *__restrict__ new_xyz, const int *__restrict__ idx,
                     float *__restrict__ rows) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % ct4);
  const size_t p = i / ct4;
  const size_t cloud = p / np_ns;
  const int k = __ldg(idx + p);
  float4 v = ld4(table, (cloud * n + k) * ct4 + c4);
  if (c4 == 0) {
    const float *ctr = new_xyz + (p / ns) * 3;
    v.x -= __ldg(ctr), v.y -= __ldg(ctr + 1), v.z -= __ldg(ctr + 2);
  }
  reinterpret_cast<float4 *>(rows)[i] = v;
}

__device__ __forceinline__ void red_add_v4(float *p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}

// grad_table[b, idx[p], :] += grad_rows[p, :]   (grad_table zero-filled by the launcher)
__global__ void __launch_bounds__(kRowThreads)
sa_group_rows_grad_kernel(size_t total4, int n, int np_ns, int ct4, const float *__restrict__ grad_rows,
                          const int *__restrict__ idx, float *__restrict__ grad_table) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % ct4);
  const size_t p = i / ct4;
  const size_t cloud = p / np_ns;
  const int k = __ldg(idx + p);
  red_add_v4(grad_table + ((cloud * n + k) * ct4 + c4) * 4, ld4(grad_rows, i));
}

// ------------------------------------------------------ feature propagation in row layout
// rows[(b, j), :] = [ ((k1*w1) + (k2*w2)) + (k3*w3) | skip[b, j, :] | 0 ]   with k_t = known[b, idx[b,j,t], :]
// (three_interpolate + torch.cat of PointnetFPModule.forward, pointnet2_modules.py:395-412).
// known (b, m, c2) and skip (b, n, c1) are point-major; c2 % 4 == 0; ct = pad4(c2 + c1).
__global__ void __launch_bounds__(kRowThreads)
fp_interp_rows_kernel(size_t total4, int n, int m, int c2, int c1, int ct4, const float *__restrict__ known,
                      const float *__restrict__ skip, const int *__restrict__ idx,
                      const float *__restrict__ weight, float *__restrict__ rows) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % ct4);
  const size_t p = i / ct4;  // = b*n + j
  float4 v;
  if (c4 * 4 < c2) {
    const size_t cloud = p / n;
    const int i1 = __ldg(idx + 3 * p), i2 = __ldg(idx + 3 * p + 1), i3 = __ldg(idx + 3 * p + 2);
    const float w1 = __ldg(weight + 3 * p), w2 = __ldg(weight + 3 * p + 1), w3 = __ldg(weight + 3 * p + 2);
    const int k4 = c2 / 4;
    const float4 a = ld4(known, (cloud * m + i1) * k4 + c4);
    const float4 b = ld4(known, (cloud * m + i2) * k4 + c4);
    const float4 c = ld4(known, (cloud * m + i3) * k4 + c4);
    v.x = __fadd_rn(__fadd_rn(__fmul_rn(a.x, w1), __fmul_rn(b.x, w2)), __fmul_rn(c.x, w3));
    v.y = __fadd_rn(__fadd_rn(__fmul_rn(a.y, w1), __fmul_rn(b.y, w2)), __fmul_rn(c.y, w3));
    v.z = __fadd_rn(__fadd_rn(__fmul_rn(a.z, w1), __fmul_rn(b.z, w2)), __fmul_rn(c.z, w3));
    v.w = __fadd_rn(__fadd_rn(__fmul_rn(a.w, w1), __fmul_rn(b.w, w2)), __fmul_rn(c.w, w3));
  } else {
    const int c = c4 * 4 - c2;
    const float *sr = skip + p * c1;
    v.x = c + 0 < c1 ? __ldg(sr + c + 0) : 0.f;
    v.y = c + 1 < c1 ? __ldg(sr + c + 1) : 0.f;
    v.z = c + 2 < c1 ? __ldg(sr + c + 2) : 0.f;
    v.w = c + 3 < c1 ? __ldg(sr + c + 3) : 0.f;
  }
  reinterpret_cast<float4 *>(rows)[i] = v;
}

// adjoint: grad_known[b, idx_t, :] += w_t * grad_rows[p, :c2] (vector reds into a zeroed buffer);
// grad_skip[b, j, :] = grad_rows[p, c2:c2+c1]
__global__ void __launch_bounds__(kRowThreads)
fp_interp_rows_grad_kernel(size_t total4, int n, int m, int c2, int c1, int ct4,
                           const float *__restrict__ grad_rows, const int *__restrict__ idx,
                           const float *__restrict__ weight, float *__restrict__ grad_known,
                           float *__restrict__ grad_skip) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;
  if (i >= total4) return;
  const int c4 = static_cast<int>(i % ct4);
  const size_t p = i / ct4;
  const float4 g = ld4(grad_rows, i);
  if (c4 * 4 < c2) {
    const size_t cloud = p / n;
    const int k4 = c2 / 4;
#pragma unroll
    for (int t = 0; t < 3; ++t) {
      const int it = __ldg(idx + 3 * p + t);
      const float w = __ldg(weight + 3 * p + t);
      red_add_v4(grad_known + ((cloud * m + it) * k4 + c4) * 4,
                 make_float4(__fmul_rn(g.x, w), __fmul_rn(g.y, w), __fmul_rn(g.z, w), __fmul_rn(g.w, w)));
    }
  } else if (grad_skip != nullptr) {
    const int c = c4 * 4 - c2;
    float *sr = grad_skip + p * c1;
    if (c + 0 < c1) sr[c + 0] = g.x;
    if (c + 1 < c1) sr[c + 1] = g.y;
    if (c + 2 < c1) sr[c + 2] = g.z;
    if (c + 3 < c1) sr[c + 3] = g.w;
  }
}

// out[p, :] = [in[p, 0 .. c_in) | 0 ...]: the network's input rows widened to a multiple of 4 columns -- the level-0
// point table [xyz | features | pad] in one launch (torch: contiguous copy of the features, zeros, cat)
__global__ void __launch_bounds__(kRowThreads)
pad_rows_kernel(size_t total4, int c_in, int c4_out, const float *__restrict__ in, float *__restrict__ out) {
  pdl_enter();
  const size_t stride = static_cast<size_t>(gridDim.x) * kRowThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x; i < total4; i += stride) {
    const size_t p = i / c4_out;
    const int c = static_cast<int>(i - p * c4_out) * 4;
    const float *src = in + p * c_in + c;
    float4 v;
    v.x = c + 0 < c_in ? __ldg(src + 0) : 0.f;
    v.y = c + 1 < c_in ? __ldg(src + 1) : 0.f;
    v.z = c + 2 < c_in ? __ldg(src + 2) : 0.f;
    v.w = c + 3 < c_in ? __ldg(src + 3) : 0.f;
    reinterpret_cast<float4 *>(out)[i] = v;
  }
}

// ---- the four gather / scatter kernels above, U float4 elements per thread -----------------
// One element per thread leaves a single dependent chain (index -> gathered row -> store) in
// flight per thread and spends most of its issue slots on 64-bit divisions.  Here a thread owns
// U elements kRowThreads apart (every access of a warp stays coalesced), all index loads are
// issued first, then all gathers, then the stores; the flat element id is split with
// multiply-shift divisions (FastDiv, exact below 2^31 elements -- larger launches take the
// one-element kernels).
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

template <int U>
__global__ void __launch_bounds__(kRowThreads)
sa_group_rows_kernel_u(unsigned total4, FastDiv ct4, FastDiv np_ns, FastDiv ns, int n,
                       const float *__restrict__ table, const float *__restrict__ new_xyz,
                       const int *__restrict__ idx, float *__restrict__ rows) {
  pdl_enter();
  const unsigned base = blockIdx.x * (kRowThreads * U) + threadIdx.x;
  unsigned p[U], c4[U];
  int k[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned e = min(base + u * kRowThreads, total4 - 1);
    p[u] = fdiv(e, ct4);
    c4[u] = e - p[u] * ct4.d;
    k[u] = __ldg(idx + p[u]);
  }
  float4 v[U];
  float cx[U], cy[U], cz[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned cloud = fdiv(p[u], np_ns);
    v[u] = ld4(table, (static_cast<size_t>(cloud) * n + k[u]) * ct4.d + c4[u]);
    cx[u] = cy[u] = cz[u] = 0.f;
    if (c4[u] == 0) {
      const float *ctr = new_xyz + static_cast<size_t>(fdiv(p[u], ns)) * 3;
      cx[u] = __ldg(ctr), cy[u] = __ldg(ctr + 1), cz[u] = __ldg(ctr + 2);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned e = base + u * kRowThreads;
    v[u].x -= cx[u], v[u].y -= cy[u], v[u].z -= cz[u];
    if (e < total4) reinterpret_cast<float4 *>(rows)[e] = v[u];
  }
}

template <int U>
__global__ void __launch_bounds__(kRowThreads)
sa_group_rows_grad_kernel_u(unsigned total4, FastDiv ct4, FastDiv np_ns, int n,
                            const float *__restrict__ grad_rows, const int *__restrict__ idx,
                            float *__restrict__ grad_table) {
  pdl_enter();
  const unsigned base = blockIdx.x * (kRowThreads * U) + threadIdx.x;
  unsigned p[U], c4[U];
  int k[U];
  float4 g[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned e = min(base + u * kRowThreads, total4 - 1);
    p[u] = fdiv(e, ct4);
    c4[u] = e - p[u] * ct4.d;
    k[u] = __ldg(idx + p[u]);
    g[u] = ld4(grad_rows, e);
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned cloud = fdiv(p[u], np_ns);
    if (base + u * kRowThreads < total4)
      red_add_v4(grad_table + ((static_cast<size_t>(cloud) * n + k[u]) * ct4.d + c4[u]) * 4, g[u]);
  }
}

template <int U>
__global__ void __launch_bounds__(kRowThreads)
fp_interp_rows_kernel_u(unsigned total4, FastDiv ct4, FastDiv nn, int m, int c2, int c1,
                        const float *__restrict__ known, const float *__restrict__ skip,
                        const int *__restrict__ idx, const float *__restrict__ weight,
                        float *__restrict__ rows) {
  pdl_enter();
  const unsigned base = blockIdx.x * (kRowThreads * U) + threadIdx.x;
  const unsigned k4 = c2 >> 2;
  unsigned p[U], c4[U];
  int i1[U], i2[U], i3[U];
  float w1[U], w2[U], w3[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned e = min(base + u * kRowThreads, total4 - 1);
    p[u] = fdiv(e, ct4);
    c4[u] = e - p[u] * ct4.d;
    if (c4[u] < k4) {
      const size_t q = static_cast<size_t>(p[u]) * 3;
      i1[u] = __ldg(idx + q), i2[u] = __ldg(idx + q + 1), i3[u] = __ldg(idx + q + 2);
      w1[u] = __ldg(weight + q), w2[u] = __ldg(weight + q + 1), w3[u] = __ldg(weight + q + 2);
    }
  }
  float4 a[U], b[U], c[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (c4[u] < k4) {
      const size_t row0 = static_cast<size_t>(fdiv(p[u], nn)) * m;
      a[u] = ld4(known, (row0 + i1[u]) * k4 + c4[u]);
      b[u] = ld4(known, (row0 + i2[u]) * k4 + c4[u]);
      c[u] = ld4(known, (row0 + i3[u]) * k4 + c4[u]);
    } else {
      const int ch = static_cast<int>(c4[u] * 4) - c2;
      const float *sr = skip + static_cast<size_t>(p[u]) * c1;
      a[u].x = ch + 0 < c1 ? __ldg(sr + ch + 0) : 0.f;
      a[u].y = ch + 1 < c1 ? __ldg(sr + ch + 1) : 0.f;
      a[u].z = ch + 2 < c1 ? __ldg(sr + ch + 2) : 0.f;
      a[u].w = ch + 3 < c1 ? __ldg(sr + ch + 3) : 0.f;
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    float4 v = a[u];
    if (c4[u] < k4) {  // ((k1*w1) + (k2*w2)) + (k3*w3), one rounding per operation (interpolate_gpu.cu:95-98)
      v.x = __fadd_rn(__fadd_rn(__fmul_rn(a[u].x, w1[u]), __fmul_rn(b[u].x, w2[u])), __fmul_rn(c[u].x, w3[u]));
      v.y = __fadd_rn(__fadd_rn(__fmul_rn(a[u].y, w1[u]), __fmul_rn(b[u].y, w2[u])), __fmul_rn(c[u].y, w3[u]));
      v.z = __fadd_rn(__fadd_rn(__fmul_rn(a[u].z, w1[u]), __fmul_rn(b[u].z, w2[u])), __fmul_rn(c[u].z, w3[u]));
      v.w = __fadd_rn(__fadd_rn(__fmul_rn(a[u].w, w1[u]), __fmul_rn(b[u].w, w2[u])), __fmul_rn(c[u].w, w3[u]));
    }
    const unsigned e = base + u * kRowThreads;
    if (e < total4) reinterpret_cast<float4 *>(rows)[e] = v;
  }
}

template <int U>
__global__ void __launch_bounds__(kRowThreads)
fp_interp_rows_grad_kernel_u(unsigned total4, FastDiv ct4, FastDiv nn, int m, int c2, int c1,
                             const float *__restrict__ grad_rows, const int *__restrict__ idx,
                             const float *__restrict__ weight, float *__restrict__ grad_known,
                             float *__restrict__ grad_skip) {
  pdl_enter();
  const unsigned base = blockIdx.x * (kRowThreads * U) + threadIdx.x;
  const unsigned k4 = c2 >> 2;
  unsigned p[U], c4[U];
  int it[U][3];
  float w[U][3];
  float4 g[U];
#pragma unroll
  for (int u = 0; u < U; ++u) {
    const unsigned e = min(base + u * kRowThreads, total4 - 1);
    p[u] = fdiv(e, ct4);
    c4[u] = e - p[u] * ct4.d;
    g[u] = ld4(grad_rows, e);
    if (c4[u] < k4) {
      const size_t q = static_cast<size_t>(p[u]) * 3;
#pragma unroll
      for (int t = 0; t < 3; ++t) it[u][t] = __ldg(idx + q + t), w[u][t] = __ldg(weight + q + t);
    }
  }
#pragma unroll
  for (int u = 0; u < U; ++u) {
    if (base + u * kRowThreads >= total4) continue;
    if (c4[u] < k4) {
      const size_t row0 = static_cast<size_t>(fdiv(p[u], nn)) * m;
#pragma unroll
      for (int t = 0; t < 3; ++t)
        red_add_v4(grad_known + ((row0 + it[u][t]) * k4 + c4[u]) * 4,
                   make_float4(__fmul_rn(g[u].x, w[u][t]), __fmul_rn(g[u].y, w[u][t]),
                               __fmul_rn(g[u].z, w[u][t]), __fmul_rn(g[u].w, w[u][t])));
    } else if (grad_skip != nullptr) {
      const int ch = static_cast<int>(c4[u] * 4) - c2;
      float *sr = grad_skip + static_cast<size_t>(p[u]) * c1;
      if (ch + 0 < c1) sr[ch + 0] = g[u].x;
      if (ch + 1 < c1) sr[ch + 1] = g[u].y;
      if (ch + 2 < c1) sr[ch + 2] = g[u].z;
      if (ch + 3 < c1) sr[ch + 3] = g[u].w;
    }
  }
}

// elements per thread of the kernels above: 2 (measured best at the SSG shapes: tools/time_rows.py; 4 costs
// occupancy) unless PN2_ROWS_U says 1 (the one-element kernels) or 4
static int rows_unroll(size_t total4) {
  static const int env = [] {
    const char *e = getenv("PN2_ROWS_U");
    const int v = e != nullptr ? atoi(e) : 2;
    return (v == 1 || v == 2 || v == 4) ? v : 2;
  }();
  return total4 < (1ull << 31) ? env : 1;
}
static inline unsigned blocks_u(size_t total4, int u) {
  return static_cast<unsigned>((total4 + static_cast<size_t>(kRowThreads) * u - 1) / (static_cast<size_t>(kRowThreads) * u));
}

static inline int blocks_for(size_t total, int cap = 0) {
  size_t nb = (total + kRowThreads - 1) / kRowThreads;
  if (cap > 0 && nb > static_cast<size_t>(cap)) nb = cap;
  return static_cast<int>(nb < 1 ? 1 : nb);
}

static inline bool rows_ok(int C, const void *p0) { return C > 0 && C % 4 == 0 && C <= 1024 && aligned16(p0); }

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_bn_stats_workspace_floats(int c) { return kMaxPartials * 2 * c; }

extern "C" int pn2_bn_stats(int p, int c, const float *z, const float *gamma, const float *beta, float eps,
                            float momentum, float *running_mean, float *running_var,
                            long long *num_batches_tracked, float *a, float *b, float *mean, float *invstd,
                            float *workspace, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(z, gamma, beta, a, b, mean, invstd, workspace);
  if (!rows_ok(c, z) || !aligned16(workspace)) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblk;
  const ColGeom g = col_geom(p, c, &nblk);
  const size_t smem = sizeof(float4) * 2 * g.row_lanes * g.C4;
  pn2::launch(bn_stats_partial_kernel, nblk, kRowThreads, smem, s, g, z, workspace);
  pn2::launch(bn_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, s, p, c, c, nblk, workspace, gamma, beta, eps, momentum,
                                                           running_mean, running_var, num_batches_tracked, a, b, mean, invstd);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_bn_fold_eval(int c, const float *gamma, const float *beta, const float *running_mean,
                                const float *running_var, float eps, float *a, float *b, float *mean,
                                float *invstd, void *stream) {
  if (c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(gamma, beta, running_mean, running_var, a, b, mean, invstd);
  pn2::launch(bn_fold_eval_kernel, (c + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream), c, gamma, beta, running_mean, running_var, eps, a, b, mean, invstd);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_bn_relu_apply(int p, int c, const float *z, const float *a, const float *b, float *h,
                                 int relu, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(z, a, b, h);
  if (!rows_ok(c, z) || !aligned16(h) || !aligned16(a) || !aligned16(b)) return PN2_ERR_UNSUPPORTED;
  const size_t total4 = static_cast<size_t>(p) * (c / 4);
  pn2::launch(bn_relu_apply_kernel, blocks_for(total4, 16 * kNumSMs), kRowThreads, 0, static_cast<cudaStream_t>(stream), total4, c / 4, z, a, b, h, relu);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_bn_relu_groupmax(int g, int ns, int c, const float *z, const float *a, const float *b,
                                    float *out, int *argmax, void *stream) {
  if (g <= 0 || ns <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(z, a, b, out, argmax);
  if (!rows_ok(c, z) || !aligned16(out) || !aligned16(argmax) || !aligned16(a) || !aligned16(b))
    return PN2_ERR_UNSUPPORTED;
  const size_t total4 = static_cast<size_t>(g) * (c / 4);
  pn2::launch(bn_relu_groupmax_kernel, blocks_for(total4), kRowThreads, 0, static_cast<cudaStream_t>(stream), g, ns, c / 4, z, a, b, out, argmax);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_groupmax_from_extrema(int g, int c, const float *zmax, const float *zmin, const int *imax,
                                         const int *imin, const float *a, const float *b, float *out,
                                         int *argmax, float *zsel, void *stream) {
  if (g <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(zmax, zmin, imax, imin, a, b, out, argmax);
  const size_t total = static_cast<size_t>(g) * c;
  pn2::launch(groupmax_from_extrema_kernel, blocks_for(total), kRowThreads, 0, static_cast<cudaStream_t>(stream), total, c, zmax, zmin, imax, imin, a, b, out, argmax, zsel);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_pool_bn_bwd(int g, int ns, int c, int training, const float *dout, const int *argmax,
                               const float *z, const float *a, const float *b, const float *mean,
                               const float *invstd, float *dz, float *dgamma, float *dbeta,
                               float *workspace, void *stream) {
  if (g <= 0 || ns <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dout, argmax, z, a, b, mean, invstd, dz, dgamma, dbeta, workspace);
  if (!rows_ok(c, z) || !aligned16(dout) || !aligned16(argmax) || !aligned16(dz) || !aligned16(workspace))
    return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblk;
  const ColGeom geom = col_geom(g, c, &nblk);
  float *c1 = workspace + static_cast<size_t>(kMaxPartials) * 2 * c, *c2 = c1 + c;
  const size_t smem = sizeof(float4) * 2 * geom.row_lanes * geom.C4;
  pn2::launch(pool_bwd_stats_kernel, nblk, kRowThreads, smem, s, geom, ns, dout, argmax, z, nullptr, a, b, mean, invstd, workspace);
  pn2::launch(bwd_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, s, g * ns, c, c, nblk, training, workspace, dgamma,
                                                            dbeta, c1, c2);
  const size_t total4 = static_cast<size_t>(g) * (c / 4);
  pn2::launch(pool_bwd_dz_kernel, blocks_for(total4), kRowThreads, 0, s, g, ns, c / 4, dout, argmax, z, a, b, mean,
                                                               invstd, c1, c2, dz);
  count_launch(3);
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_relu_bn_bwd(int p, int c, int training, int relu, const float *dh, const float *z,
                               const float *a, const float *b, const float *mean, const float *invstd,
                               float *dz, float *dgamma, float *dbeta, float *workspace, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dh, z, a, b, mean, invstd, dz, dgamma, dbeta, workspace);
  if (!rows_ok(c, z) || !aligned16(dh) || !aligned16(dz) || !aligned16(workspace)) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblk;
  const ColGeom geom = col_geom(p, c, &nblk);
  float *c1 = workspace + static_cast<size_t>(kMaxPartials) * 2 * c, *c2 = c1 + c;
  const size_t smem = sizeof(float4) * 2 * geom.row_lanes * geom.C4;
  pn2::launch(relu_bwd_stats_kernel, nblk, kRowThreads, smem, s, geom, relu, dh, z, a, b, mean, invstd, workspace);
  pn2::launch(bwd_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, s, p, c, c, nblk, training, workspace, dgamma, dbeta,
                                                            c1, c2);
  const size_t total4 = static_cast<size_t>(p) * (c / 4);
  pn2::launch(relu_bwd_dz_kernel, blocks_for(total4, 16 * kNumSMs), kRowThreads, 0, s, total4, c / 4, relu, dh, z, a, b,
                                                                             mean, invstd, c1, c2, dz);
  count_launch(3);
  PN2_RETURN_LAUNCH_STATUS();
}

// The dZ half of pn2_relu_bn_bwd alone, for a layer whose statistics c1c2 = [dbeta / P | dgamma / P] came from
// elsewhere (the data-gradient product's epilogue, pn2_mlp_dgrad_stats + pn2_bwd_stats_finalize).
extern "C" int pn2_relu_bn_bwd_dz(int p, int c, int relu, const float *dh, const float *z, const float *a,
                                  const float *b, const float *mean, const float *invstd, const float *c1c2,
                                  float *dz, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dh, z, a, b, mean, invstd, c1c2, dz);
  if (!rows_ok(c, z) || !aligned16(dh) || !aligned16(dz) || !aligned16(c1c2)) return PN2_ERR_UNSUPPORTED;
  const size_t total4 = static_cast<size_t>(p) * (c / 4);
  pn2::launch(relu_bwd_dz_kernel, blocks_for(total4, 16 * kNumSMs), kRowThreads, 0, static_cast<cudaStream_t>(stream), total4, c / 4, relu, dh, z, a, b, mean, invstd, c1c2, c1c2 + c, dz);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

// Statistics halves of the two adjoints above, for the fused tensor-core backward layer
// (mlp_bwd.cu), which forms dZ on chip: dgamma, dbeta and c1c2 = [dbeta / P | dgamma / P].
extern "C" int pn2_pool_bn_bwd_stats(int g, int ns, int c, int training, const float *dout, const int *argmax,
                                     const float *z, const float *zsel, const float *a, const float *b, const float *mean,
                                     const float *invstd, float *dgamma, float *dbeta, float *c1c2,
                                     float *workspace, void *stream) {
  if (g <= 0 || ns <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dout, argmax, z, a, b, mean, invstd, dgamma, dbeta, c1c2, workspace);
  if (!rows_ok(c, z) || !aligned16(dout) || !aligned16(argmax) || !aligned16(workspace)) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblk;
  const ColGeom geom = col_geom(g, c, &nblk);
  const size_t smem = sizeof(float4) * 2 * geom.row_lanes * geom.C4;
  if (zsel != nullptr && !aligned16(zsel)) return PN2_ERR_UNSUPPORTED;
  pn2::launch(pool_bwd_stats_kernel, nblk, kRowThreads, smem, s, geom, ns, dout, argmax, z, zsel, a, b, mean, invstd, workspace);
  pn2::launch(bwd_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, s, g * ns, c, c, nblk, training, workspace, dgamma, dbeta, c1c2, c1c2 + c);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_relu_bn_bwd_stats(int p, int c, int training, int relu, const float *dh, const float *z,
                                     const float *a, const float *b, const float *mean, const float *invstd,
                                     float *dgamma, float *dbeta, float *c1c2, float *workspace, void *stream) {
  if (p <= 0 || c <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(dh, z, a, b, mean, invstd, dgamma, dbeta, c1c2, workspace);
  if (!rows_ok(c, z) || !aligned16(dh) || !aligned16(workspace)) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  int nblk;
  const ColGeom geom = col_geom(p, c, &nblk);
  const size_t smem = sizeof(float4) * 2 * geom.row_lanes * geom.C4;
  pn2::launch(relu_bwd_stats_kernel, nblk, kRowThreads, smem, s, geom, relu, dh, z, a, b, mean, invstd, workspace);
  pn2::launch(bwd_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, s, p, c, c, nblk, training, workspace, dgamma, dbeta, c1c2, c1c2 + c);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}

// partials [nblk][2][ld] (column sums of dY and dY * zhat from pn2_mlp_bwd_layer's epilogue) ->
// dgamma, dbeta, c1c2 of the layer below
extern "C" int pn2_bwd_stats_finalize(int p, int c, int ld, int nblk, int training, const float *partials,
                                      float *dgamma, float *dbeta, float *c1c2, void *stream) {
  if (p <= 0 || c <= 0 || ld < c || nblk <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(partials, dgamma, dbeta, c1c2);
  pn2::launch(bwd_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, static_cast<cudaStream_t>(stream), p, c, ld, nblk, training, partials, dgamma,
                                                                   dbeta, c1c2, c1c2 + c);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_pad_rows(long long p, int c_in, int c_out, const float *in, float *out, void *stream) {
  if (p <= 0 || c_in <= 0 || c_out < c_in || c_out % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(in, out);
  if (!aligned16(out)) return PN2_ERR_UNSUPPORTED;
  const size_t total4 = static_cast<size_t>(p) * (c_out / 4);
  pn2::launch(pad_rows_kernel, blocks_for(total4, 16 * kNumSMs), kRowThreads, 0, static_cast<cudaStream_t>(stream), total4,
              c_in, c_out / 4, in, out);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_sa_group_rows(int b, int n, int npoint, int nsample, int ct, const float *table,
                                 const float *new_xyz, const int *idx, float *rows, void *stream) {
  if (b <= 0 || n <= 0 || npoint <= 0 || nsample <= 0 || ct < 4) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(table, new_xyz, idx, rows);
  if (!rows_ok(ct, table) || !aligned16(rows)) return PN2_ERR_UNSUPPORTED;
  const size_t total4 = static_cast<size_t>(b) * npoint * nsample * (ct / 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int u = rows_unroll(total4);
  const FastDiv dct = fast_div(ct / 4), dcl = fast_div(npoint * nsample), dns = fast_div(nsample);
  if (u == 4)
    pn2::launch(sa_group_rows_kernel_u<4>, blocks_u(total4, 4), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dcl, dns, n, table, new_xyz, idx, rows);
  else if (u == 2)
    pn2::launch(sa_group_rows_kernel_u<2>, blocks_u(total4, 2), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dcl, dns, n, table, new_xyz, idx, rows);
  else
    pn2::launch(sa_group_rows_kernel, blocks_for(total4), kRowThreads, 0, s, total4, n, npoint * nsample, nsample, ct / 4, table, new_xyz, idx, rows);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_sa_group_rows_grad(int b, int n, int npoint, int nsample, int ct, const float *grad_rows,
                                      const int *idx, float *grad_table, void *stream) {
  if (b <= 0 || n <= 0 || npoint <= 0 || nsample <= 0 || ct < 4) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(grad_rows, idx, grad_table);
  if (!rows_ok(ct, grad_rows) || !aligned16(grad_table)) return PN2_ERR_UNSUPPORTED;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(grad_table, 0, sizeof(float) * static_cast<size_t>(b) * n * ct, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t total4 = static_cast<size_t>(b) * npoint * nsample * (ct / 4);
  const int u = rows_unroll(total4);
  const FastDiv dct = fast_div(ct / 4), dcl = fast_div(npoint * nsample);
  if (u == 4)
    pn2::launch(sa_group_rows_grad_kernel_u<4>, blocks_u(total4, 4), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dcl, n, grad_rows, idx, grad_table);
  else if (u == 2)
    pn2::launch(sa_group_rows_grad_kernel_u<2>, blocks_u(total4, 2), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dcl, n, grad_rows, idx, grad_table);
  else
    pn2::launch(sa_group_rows_grad_kernel, blocks_for(total4), kRowThreads, 0, s, total4, n, npoint * nsample, ct / 4,
                                                                        grad_rows, idx, grad_table);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_fp_interp_rows(int b, int n, int m, int c2, int c1, const float *known, const float *skip,
                                  const int *idx, const float *weight, float *rows, void *stream) {
  if (b <= 0 || n <= 0 || m <= 0 || c2 <= 0 || c1 < 0 || c2 % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(known, idx, weight, rows);
  if (c1 > 0) PN2_CHECK_PTRS(skip);
  const int ct = (c2 + c1 + 3) / 4 * 4;
  if (!aligned16(known) || !aligned16(rows) || ct > 4096) return PN2_ERR_UNSUPPORTED;
  const size_t total4 = static_cast<size_t>(b) * n * (ct / 4);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const int u = rows_unroll(total4);
  const FastDiv dct = fast_div(ct / 4), dn = fast_div(n);
  if (u == 4)
    pn2::launch(fp_interp_rows_kernel_u<4>, blocks_u(total4, 4), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dn, m, c2, c1, known, skip, idx, weight, rows);
  else if (u == 2)
    pn2::launch(fp_interp_rows_kernel_u<2>, blocks_u(total4, 2), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dn, m, c2, c1, known, skip, idx, weight, rows);
  else
    pn2::launch(fp_interp_rows_kernel, blocks_for(total4), kRowThreads, 0, s, total4, n, m, c2, c1, ct / 4, known, skip, idx, weight, rows);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_fp_interp_rows_grad(int b, int n, int m, int c2, int c1, const float *grad_rows,
                                       const int *idx, const float *weight, float *grad_known,
                                       float *grad_skip, void *stream) {
  if (b <= 0 || n <= 0 || m <= 0 || c2 <= 0 || c1 < 0 || c2 % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(grad_rows, idx, weight, grad_known);
  const int ct = (c2 + c1 + 3) / 4 * 4;
  if (!aligned16(grad_known) || !aligned16(grad_rows) || ct > 4096) return PN2_ERR_UNSUPPORTED;
  if (grad_skip != nullptr && !aligned4(grad_skip)) return PN2_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  cudaError_t e = cudaMemsetAsync(grad_known, 0, sizeof(float) * static_cast<size_t>(b) * m * c2, s);
  if (e != cudaSuccess) return static_cast<int>(e);
  const size_t total4 = static_cast<size_t>(b) * n * (ct / 4);
  const int u = rows_unroll(total4);
  const FastDiv dct = fast_div(ct / 4), dn = fast_div(n);
  if (u == 4)
    pn2::launch(fp_interp_rows_grad_kernel_u<4>, blocks_u(total4, 4), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dn, m, c2, c1, grad_rows, idx, weight, grad_known, grad_skip);
  else if (u == 2)
    pn2::launch(fp_interp_rows_grad_kernel_u<2>, blocks_u(total4, 2), kRowThreads, 0, s, static_cast<unsigned>(total4), dct, dn, m, c2, c1, grad_rows, idx, weight, grad_known, grad_skip);
  else
    pn2::launch(fp_interp_rows_grad_kernel, blocks_for(total4), kRowThreads, 0, s, total4, n, m, c2, c1, ct / 4, grad_rows,
                                                                         idx, weight, grad_known, grad_skip);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

// Finalise BatchNorm batch statistics from per-CTA partial sums produced elsewhere (the
// tensor-core GEMM's epilogue): partials [nblk][2][ld], first c of every ld columns valid.
extern "C" int pn2_bn_finalize(int p, int c, int ld, int nblk, const float *partials, const float *gamma,
                               const float *beta, float eps, float momentum, float *running_mean,
                               float *running_var, long long *num_batches_tracked, float *a, float *b,
                               float *mean, float *invstd, void *stream) {
  if (p <= 0 || c <= 0 || ld < c || nblk <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(partials, gamma, beta, a, b, mean, invstd);
  pn2::launch(bn_stats_finalize_kernel, (c + kFinalizeWarps - 1) / kFinalizeWarps, 32 * kFinalizeWarps, 0, static_cast<cudaStream_t>(stream), p, c, ld, nblk, partials, gamma, beta, eps,
                                                                  momentum, running_mean, running_var,
                                                                  num_batches_tracked, a, b, mean, invstd);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

// ---------------------------------------------------------------------------------------
// Small training-step utilities on flat fp32 buffers (one launch each instead of one per tensor)
// ---------------------------------------------------------------------------------------
// out[j] = sum_s in[s, j]  (fixed order): the S partial weight gradients of a split-K GEMM.
// One WARP per float4 of the output: lane l adds partials l, l+32, ... and a shuffle tree
// combines the lanes, so even a 32x32 weight gradient summed over ~300 slabs uses 256 warps.
__global__ void __launch_bounds__(kRowThreads)
sum_partials_kernel(int S, int M4, const float *__restrict__ in, float *__restrict__ out) {
  pdl_enter();
  const int j = blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= M4) return;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kBatch = 5;  // loads of a batch are issued together; additions keep the plain loop's order
  for (int s0 = lane; s0 < S; s0 += 32 * kBatch) {
    float4 v[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      v[i] = s0 + 32 * i < S ? ld4(in, static_cast<size_t>(s0 + 32 * i) * M4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      if (s0 + 32 * i < S) acc.x += v[i].x, acc.y += v[i].y, acc.z += v[i].z, acc.w += v[i].w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o), acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o), acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (lane == 0) reinterpret_cast<float4 *>(out)[j] = acc;
}

// Few output columns, many partials (the [~900][3][C] coordinate-weight partials of mlp_lead.cu): one CTA per float4
// column, its 8 warps take contiguous eighths of the partials, then warp 0 adds the eight results in order
// (one batch of loads per lane instead of six dependent ones: 16 -> ~4 us at 888 x 96 floats).
__global__ void __launch_bounds__(kRowThreads)
sum_partials_wide_kernel(int S, int M4, const float *__restrict__ in, float *__restrict__ out) {
  pdl_enter();
  __shared__ float4 s_w[kRowThreads / 32];
  const int j = blockIdx.x, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kRowThreads / 32;
  const int chunk = (S + kWarps - 1) / kWarps;
  const int s_lo = warp * chunk, s_hi = min(S, s_lo + chunk);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kBatch = 5;
  for (int s0 = s_lo + lane; s0 < s_hi; s0 += 32 * kBatch) {
    float4 v[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      v[i] = s0 + 32 * i < s_hi ? ld4(in, static_cast<size_t>(s0 + 32 * i) * M4 + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      if (s0 + 32 * i < s_hi) acc.x += v[i].x, acc.y += v[i].y, acc.z += v[i].z, acc.w += v[i].w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o), acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o), acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (lane == 0) s_w[warp] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float4 t = s_w[0];
    for (int w = 1; w < kWarps; ++w) t.x += s_w[w].x, t.y += s_w[w].y, t.z += s_w[w].z, t.w += s_w[w].w;
    reinterpret_cast<float4 *>(out)[j] = t;
  }
}

extern "C" int pn2_sum_partials(int s, int m, const float *in, float *out, void *stream) {
  if (s <= 0 || m <= 0 || m % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(in, out);
  if (!aligned16(in) || !aligned16(out)) return PN2_ERR_UNSUPPORTED;
  const int warps_per_cta = kRowThreads / 32;
  if (m / 4 <= 2 * kNumSMs && s >= 256)   // few columns, many partials: a CTA per column
    pn2::launch(sum_partials_wide_kernel, m / 4, kRowThreads, 0, static_cast<cudaStream_t>(stream), s, m / 4, in, out);
  else
    pn2::launch(sum_partials_kernel, (m / 4 + warps_per_cta - 1) / warps_per_cta, kRowThreads, 0, static_cast<cudaStream_t>(stream), s, m / 4, in, out);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

// The same sum for padded weight-gradient images: in[s][rows_in][cols_in] (cols_in % 4 == 0) -> the dense
// leading block out[rows_out][cols_out] (any cols_out <= cols_in): the tensor a parameter's .grad takes as
// it is (no strided view that autograd would have to copy).  Same warp-per-float4 scheme and order.
__global__ void __launch_bounds__(kRowThreads)
sum_partials_block_kernel(int S, int rows_in, int cols_in, int rows_out, int cols_out,
                          const float *__restrict__ in, float *__restrict__ out) {
  pdl_enter();
  const int c4n = (cols_out + 3) / 4;
  const int j = blockIdx.x * (kRowThreads / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (j >= rows_out * c4n) return;
  const int r = j / c4n, c4 = j - r * c4n;
  const size_t image4 = static_cast<size_t>(rows_in) * cols_in / 4, off4 = (static_cast<size_t>(r) * cols_in) / 4 + c4;
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  constexpr int kBatch = 5;
  for (int s0 = lane; s0 < S; s0 += 32 * kBatch) {
    float4 v[kBatch];
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      v[i] = s0 + 32 * i < S ? ld4(in, static_cast<size_t>(s0 + 32 * i) * image4 + off4) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int i = 0; i < kBatch; ++i)
      if (s0 + 32 * i < S) acc.x += v[i].x, acc.y += v[i].y, acc.z += v[i].z, acc.w += v[i].w;
  }
  for (int o = 16; o > 0; o >>= 1) {
    acc.x += __shfl_xor_sync(0xffffffffu, acc.x, o), acc.y += __shfl_xor_sync(0xffffffffu, acc.y, o);
    acc.z += __shfl_xor_sync(0xffffffffu, acc.z, o), acc.w += __shfl_xor_sync(0xffffffffu, acc.w, o);
  }
  if (lane == 0) {
    float *dst = out + static_cast<size_t>(r) * cols_out + 4 * c4;
    const float v[4] = {acc.x, acc.y, acc.z, acc.w};
    for (int e = 0; e < 4; ++e)
      if (4 * c4 + e < cols_out) dst[e] = v[e];
  }
}

extern "C" int pn2_sum_partials_block(int s, int rows_in, int cols_in, int rows_out, int cols_out, const float *in,
                                      float *out, void *stream) {
  if (s <= 0 || rows_out <= 0 || cols_out <= 0 || rows_out > rows_in || cols_out > cols_in || cols_in % 4 != 0)
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(in, out);
  if (!aligned16(in)) return PN2_ERR_UNSUPPORTED;
  const int warps_per_cta = kRowThreads / 32;
  const int items = rows_out * ((cols_out + 3) / 4);
  pn2::launch(sum_partials_block_kernel, (items + warps_per_cta - 1) / warps_per_cta, kRowThreads, 0, static_cast<cudaStream_t>(stream), s, rows_in, cols_in, rows_out, cols_out, in, out);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

// Dropout keep mask (one byte per element) for the classifier's nn.Dropout inside the fused chain
// (pointnet2_semseg.py:156): counter-based -- element i of launch t keeps iff the i-th 16-bit lane of
// splitmix64(seed, t, i / 4) is below keep * 65536 -- so the launch is CUDA-graph replayable (the launch
// counter lives on the device and is bumped by the launch) and one pass at store bandwidth replaces
// torch's Philox bernoulli_ over bytes (47 us for the head's [262144, 128] mask).
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}
__global__ void __launch_bounds__(kRowThreads)
dropout_mask_kernel(size_t n16, size_t n, unsigned threshold, const long long *__restrict__ state,
                    unsigned char *__restrict__ mask) {
  pdl_enter();
  const size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x;  // 16 mask bytes per thread
  if (i >= n16) return;
  const unsigned long long key = splitmix64(static_cast<unsigned long long>(state[0]) ^
                                            (static_cast<unsigned long long>(state[1]) * 0xD1342543DE82EF95ull));
  unsigned w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const unsigned long long h = splitmix64(key + (i * 4 + k));
    w[k] = ((h & 0xffffull) < threshold ? 1u : 0u) | (((h >> 16) & 0xffffull) < threshold ? 0x100u : 0u) |
           (((h >> 32) & 0xffffull) < threshold ? 0x10000u : 0u) | ((h >> 48) < threshold ? 0x1000000u : 0u);
  }
  if (i * 16 + 16 <= n) {
    reinterpret_cast<uint4 *>(mask)[i] = make_uint4(w[0], w[1], w[2], w[3]);
  } else {
    for (size_t b = i * 16; b < n; ++b) mask[b] = static_cast<unsigned char>((w[(b >> 2) & 3] >> (8 * (b & 3))) & 1u);
  }
}
__global__ void bump_counter_kernel(long long *state) {
  pdl_enter(); state[1] += 1; }

extern "C" int pn2_dropout_mask(long long n, float keep, long long *state, unsigned char *mask, void *stream) {
  if (n <= 0 || !(keep >= 0.f && keep <= 1.f)) return PN2_ERR_INVALID_ARG;
  if (state == nullptr || mask == nullptr) return PN2_ERR_INVALID_ARG;
  if (!aligned16(mask) || (reinterpret_cast<uintptr_t>(state) & 7u) != 0) return PN2_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const size_t n16 = (static_cast<size_t>(n) + 15) / 16;
  const unsigned threshold = static_cast<unsigned>(static_cast<double>(keep) * 65536.0 + 0.5);
  pn2::launch(dropout_mask_kernel, blocks_for(n16), kRowThreads, 0, s, n16, static_cast<size_t>(n), threshold, state, mask);
  pn2::launch(bump_counter_kernel, 1, 1, 0, s, state);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}

// Adam (torch.optim.Adam semantics: L2 weight decay added to the gradient, bias correction,
// eps outside the square root) over ONE flat parameter buffer; the step count lives on the
// device so that the launch can be replayed from a CUDA graph.  grad_scale multiplies the
// gradient first (1/world after a summed all-reduce).
__global__ void __launch_bounds__(kRowThreads)
adam_flat_kernel(size_t n, float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
                 float *__restrict__ v, const long long *__restrict__ step, float lr, float beta1, float beta2,
                 float eps, float weight_decay, float grad_scale) {
  pdl_enter();
  const double t = static_cast<double>(*step);
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), t));
  const float bc2s = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(beta2), t)));
  const float step_size = lr / bc1;
  const size_t stride = static_cast<size_t>(gridDim.x) * kRowThreads;
  for (size_t i = static_cast<size_t>(blockIdx.x) * kRowThreads + threadIdx.x; i < n; i += stride) {
    float grad = g[i] * grad_scale;
    const float w = p[i];
    if (weight_decay != 0.f) grad = fmaf(weight_decay, w, grad);
    const float mi = beta1 * m[i] + (1.f - beta1) * grad;
    const float vi = beta2 * v[i] + (1.f - beta2) * grad * grad;
    m[i] = mi, v[i] = vi;
    p[i] = w - step_size * mi / (sqrtf(vi) / bc2s + eps);
  }
}
__global__ void bump_step_kernel(long long *step) {
  pdl_enter(); *step += 1; }

extern "C" int pn2_adam_flat(long long n, float *param, const float *grad, float *exp_avg, float *exp_avg_sq,
                             long long *step, float lr, float beta1, float beta2, float eps,
                             float weight_decay, float grad_scale, void *stream) {
  if (n <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(param, grad, exp_avg, exp_avg_sq);
  if (step == nullptr) return PN2_ERR_INVALID_ARG;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  pn2::launch(bump_step_kernel, 1, 1, 0, s, step);
  pn2::launch(adam_flat_kernel, blocks_for(static_cast<size_t>(n), 8 * kNumSMs), kRowThreads, 0, s, static_cast<size_t>(n), param, grad, exp_avg, exp_avg_sq, step, lr, beta1, beta2, eps, weight_decay,
      grad_scale);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}
