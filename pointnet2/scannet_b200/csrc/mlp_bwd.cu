// Fused backward of one shared-MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM).
//
// One Conv2d(1x1) -> BatchNorm2d(train) -> ReLU unit of the reference's SharedMLP
// (/root/reference/pointnet2/pytorch_utils.py:11-36, used by pointnet2_modules.py:65-71, 414) in
// rows layout, Z[P, N] = H_prev[P, K] * W[N, K]^T.  torch.autograd runs its adjoint as separate
// kernels (threshold_backward, cudnn bn_bw, two GEMMs), each re-reading [P, C] tensors from HBM.
// Here the whole layer adjoint is ONE pass over the rows:
//
//   dZ      = a * (dY - c1 - zhat * c2)      dY = dH masked by this layer's ReLU (or, for the
//                                            pooled last layer, the arg-max scatter of dOut)
//   dH_prev = dZ * W                          GEMM 1  (M = 128 rows, N = K channels, reduce over N)
//   dW     += dZ^T * H_prev                   GEMM 2  (M = N, N = K, reduce over the 128 rows)
//   s, q   += column sums of dY_prev, dY_prev * zhat_prev    (the NEXT layer adjoint's BN constants)
//
// dZ is formed on chip from (dH, Z) and never written; H_prev = relu(a_in * Z_prev + b_in) is
// recomputed from Z_prev (the forward pass does not store it); dW accumulates in tensor memory over
// all of a CTA's tiles and leaves the SM once.  HBM traffic per layer: read dH, Z, Z_prev, write
// dH_prev -- against 7 reads/writes of [P, N] plus 2 of [P, K] for the unfused sequence.
//
// GEMM 2 reduces over ROWS, so its operands are read MN-major straight from row-major slabs
// ([128 rows][32 fp32], slab stride = leading byte offset): dZ^T and H_prev are never transposed in
// memory.  MN-major TF32 operands must use the 32-byte-atom 128-byte swizzle, GEMM 1's K-major A
// operand the ordinary 16-byte-atom one, so a dZ slab is written twice by its loader: once into the
// tile's dZ set (GEMM 2) and once into the group's small staging slot (GEMM 1, freed one step later).
//
// Structure (persistent CTAs, one per SM, 17 warps):
//   warps 0-3   epilogue: tcgen05.ld of dH_prev, row stores, ReLU mask of the previous layer from
//               Z_prev and the butterfly column sums for its BN adjoint; at the end dW TMEM -> HBM
//   warp  4     one elected thread issues every tcgen05.mma (kind::tf32) and commit
//   warps 5-16  three loader groups of 128 threads taking the tile's slabs round-robin: 16-byte
//               LDGs (dH and Z, or Z_prev) -> registers -> BN adjoint / fold + ReLU -> TF32 ->
//               swizzled st.shared -> fence.proxy.async -> mbarrier arrive; W^T tiles by 1-D TMA.
//
// Precision: TF32 operands (round-to-nearest), fp32 accumulation -- the arithmetic class of the
// reference's own convolutions (torch.backends.cudnn.allow_tf32 defaults to True).
#include "tc_common.cuh"

namespace pn2 {
namespace bwd {

using namespace tc;

constexpr int BM = 128;                 // rows per tile
constexpr int BK = 32;                  // channels per slab = one 128-byte swizzle row
constexpr int SLAB = BM * BK * 4;       // 16 KB
constexpr int kEpiWarps = 4, kGroups = 3, kGroupWarps = 4;
constexpr int kMmaWarp = kEpiWarps;     // warp 4
constexpr int kFirstLoadWarp = kEpiWarps + 1;
constexpr int kThreads = 32 * (kEpiWarps + 1 + kGroups * kGroupWarps);  // 544
constexpr int kMaxNB = 4;               // N <= 128
constexpr int kMaxKB = 5;               // K <= 160
// Every loader group owns ONE H_prev slab and ONE W^T tile slot: a slot then has a single, in-order
// producer, so its phase-parity waits can never alias (with a shared ring a group running two uses
// ahead of a lagging group would see the parity it expects and overwrite a live slab).

struct Params {
  // this layer, N output channels
  const float *dH;       // [P, N] upstream gradient wrt the layer's activated output (dense mode)
  const float *dout;     // [P / ns, N] gradient of the pooled output (pooled mode: dH == nullptr)
  const int *argmax;     // [P / ns, N] arg-max row of every (group, channel)
  const float *Z;        // [P, N] pre-BN output of this layer
  const float *a, *b, *mean, *invstd;  // [N] BN fold and statistics of this layer
  const float *c1c2;     // [2][N] dbeta / P, dgamma / P (0 in eval mode)
  // the layer below, K channels
  const float *Aprev;    // [P, lda]: Z of the previous layer (a_in != nullptr) or the block's input rows
  const float *a_in, *b_in, *mean_in, *invstd_in;  // [K] fold / statistics of the previous layer
  const float *WTp;      // packed W^T tiles [nb][kpad16][32] (nullptr: dH_prev not wanted)
  float *dHprev;         // [P, lda]
  float *dWpart;         // [gridDim.x][nb * 32][kb * 32]
  float *stats_part;     // [gridDim.x][2][kb * 32]  (nullptr: no BN below)
  int P, N, nb, K, lda, kb, kpad16;
  int ns, relu, relu_in;
  int groups, sets, w_slot_bytes, tmem_cols;  // loader groups in use; dZ sets (2, or 1 when N > 64)
};

struct __align__(8) Barriers {
  uint64_t dz_full[2][kMaxNB];   // slab c of set s holds dZ           (128 arrivals: one loader group)
  uint64_t dz_empty[4];          // both products of tile t retired: barrier t & 3 (1 commit).  Four
                                 // barriers for <= two sets keep a loader at most one phase behind.
  uint64_t w_full[kGroups];      // W^T tile landed in the group's slot (1 arrival + tx bytes)
  uint64_t w_empty[kGroups];     // GEMM 1 step retired: W^T slot and K-major dZ staging slot free (1 commit)
  uint64_t h_full[kGroups];      // H_prev slab ready                   (128 arrivals)
  uint64_t h_empty[kGroups];     // GEMM 2 on the slab retired          (1 commit)
  uint64_t acc_full[2];          // dH_prev accumulator complete        (1 commit)
  uint64_t acc_empty[2];         // drained by the epilogue             (128 arrivals)
  uint64_t dw_full;              // every MMA of the CTA retired        (1 commit)
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 tf32x4(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}

__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_layer_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int nb = p.nb, kb = p.kb, G = p.groups;
  const int nb32 = nb * 32, kb32 = kb * 32;
  unsigned char *s_dz = smem;                                   // [sets][nb] slabs, MN-major image (GEMM 2)
  unsigned char *s_h = s_dz + p.sets * nb * SLAB;               // [G] H_prev slabs, one per loader group
  unsigned char *s_a1 = s_h + G * SLAB;                         // [G] K-major dZ staging slabs (GEMM 1)
  unsigned char *s_w = s_a1 + G * SLAB;                         // [G] W^T tiles
  float *s_tab = reinterpret_cast<float *>(s_w + G * p.w_slot_bytes);
  float *t_a = s_tab, *t_b = t_a + nb32, *t_mu = t_b + nb32, *t_is = t_mu + nb32, *t_c1 = t_is + nb32,
        *t_c2 = t_c1 + nb32;
  float *u_a = t_c2 + nb32, *u_b = u_a + kb32, *u_mu = u_b + kb32, *u_is = u_mu + kb32;
  float *s_stats = u_is + kb32;                                 // [kEpiWarps][2][kb32]
  Barriers *bars = reinterpret_cast<Barriers *>(s_stats + kEpiWarps * 2 * kb32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = (p.P + BM - 1) / BM;
  const int my_tiles = blockIdx.x < m_tiles ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const bool dgrad = p.WTp != nullptr;
  const bool stats = p.stats_part != nullptr;

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      for (int c = 0; c < kMaxNB; ++c) mbar_init(&bars->dz_full[s][c], 32 * kGroupWarps);
      mbar_init(&bars->acc_full[s], 1);
      mbar_init(&bars->acc_empty[s], 32 * kEpiWarps);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&bars->dz_empty[s], 1);
    for (int s = 0; s < kGroups; ++s) {
      mbar_init(&bars->w_full[s], 1), mbar_init(&bars->w_empty[s], 1);
      mbar_init(&bars->h_full[s], 32 * kGroupWarps), mbar_init(&bars->h_empty[s], 1);
    }
    mbar_init(&bars->dw_full, 1);
    mbar_fence_init();
  }
  if (warp == kMmaWarp) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&bars->tmem_base)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < nb32; i += kThreads) {
    const bool ok = i < p.N;
    t_a[i] = ok ? p.a[i] : 0.f, t_b[i] = ok ? p.b[i] : 0.f;
    t_mu[i] = ok ? p.mean[i] : 0.f, t_is[i] = ok ? p.invstd[i] : 0.f;
    t_c1[i] = ok ? p.c1c2[i] : 0.f, t_c2[i] = ok ? p.c1c2[p.N + i] : 0.f;
  }
  for (int i = tid; i < kb32; i += kThreads) {
    const bool ok = p.a_in != nullptr && i < p.K;
    u_a[i] = ok ? p.a_in[i] : 1.f, u_b[i] = ok ? p.b_in[i] : 0.f;
    u_mu[i] = ok && p.mean_in != nullptr ? p.mean_in[i] : 0.f;
    u_is[i] = ok && p.invstd_in != nullptr ? p.invstd_in[i] : 0.f;
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  // tensor memory: [acc 0 | acc 1] (kb32 columns each, only when dgrad) then dW (kb32 columns)
  const uint32_t tm_acc = tmem_base, tm_dw = tmem_base + (dgrad ? 2 * kb32 : 0);

  if (warp >= kFirstLoadWarp) {
    // ======================================= loaders =======================================
    const int lw = warp - kFirstLoadWarp, grp = lw / kGroupWarps;
    const int gt = (lw % kGroupWarps) * 32 + lane;  // 0..127 within the group
    const int chunk = gt & 7, row0 = gt >> 3;       // 16-byte chunk of the slab row; rows row0 + 16 i
    const int ipt = nb + kb;                         // items (slabs) per tile: dZ slabs, then H slabs
    int tl = grp < G ? 0 : my_tiles, j = grp;  // groups beyond G idle
    while (j >= ipt) j -= ipt, ++tl;
    int w_use = 0, h_use = 0;                  // uses of this group's own W^T / H_prev slot so far
    while (tl < my_tiles) {
      const int tile = blockIdx.x + tl * gridDim.x;
      if (j < nb) {
        // ---------------- dZ slab c: columns [32c, 32c+32) of this layer ----------------
        const int c = j, col = c * BK + chunk * 4;
        const bool colok = col < p.N;
        float4 d[8], z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int grow = tile * BM + row0 + 16 * i;
          const bool ok = colok && grow < p.P;
          d[i] = z[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (ok) {
            z[i] = ldg4(p.Z + static_cast<size_t>(grow) * p.N + col);
            if (p.dH != nullptr) {
              d[i] = ldg4(p.dH + static_cast<size_t>(grow) * p.N + col);
            } else {  // pooled layer: the gradient lives on the arg-max row of every (group, channel)
              const int g = grow / p.ns, s = grow - g * p.ns;
              const float4 dd = ldg4(p.dout + static_cast<size_t>(g) * p.N + col);
              const int4 am = __ldg(reinterpret_cast<const int4 *>(p.argmax + static_cast<size_t>(g) * p.N + col));
              d[i] = make_float4(am.x == s ? dd.x : 0.f, am.y == s ? dd.y : 0.f, am.z == s ? dd.z : 0.f,
                                 am.w == s ? dd.w : 0.f);
            }
          }
        }
        const float4 fa = *reinterpret_cast<const float4 *>(t_a + col), fb = *reinterpret_cast<const float4 *>(t_b + col);
        const float4 mu = *reinterpret_cast<const float4 *>(t_mu + col), is = *reinterpret_cast<const float4 *>(t_is + col);
        const float4 k1 = *reinterpret_cast<const float4 *>(t_c1 + col), k2 = *reinterpret_cast<const float4 *>(t_c2 + col);
        const int set = p.sets == 2 ? (tl & 1) : 0;
        if (dgrad) mbar_wait(&bars->w_empty[grp], (w_use & 1) ^ 1);  // W^T slot + staging slot free
        if (dgrad && gt == 0) {  // this step's W^T tile: one contiguous pre-swizzled image
          const uint32_t bytes = static_cast<uint32_t>(p.kpad16) * BK * 4;
          mbar_expect_tx(&bars->w_full[grp], bytes);
          tma_load_1d(s_w + grp * p.w_slot_bytes, p.WTp + static_cast<size_t>(c) * p.kpad16 * BK, bytes,
                      &bars->w_full[grp]);
        }
        ++w_use;
        if (tl >= p.sets)  // the set's previous tile has retired
          mbar_wait(&bars->dz_empty[(tl - p.sets) & 3], ((tl - p.sets) >> 2) & 1);
        unsigned char *slab = s_dz + (set * nb + c) * SLAB;
        unsigned char *stage = s_a1 + grp * SLAB;
        const bool relu = p.relu != 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + 16 * i;
          const float4 v = z[i];
          float4 o;
          o.x = fa.x * (((!relu || fmaf(fa.x, v.x, fb.x) > 0.f) ? d[i].x : 0.f) - k1.x - (v.x - mu.x) * is.x * k2.x);
          o.y = fa.y * (((!relu || fmaf(fa.y, v.y, fb.y) > 0.f) ? d[i].y : 0.f) - k1.y - (v.y - mu.y) * is.y * k2.y);
          o.z = fa.z * (((!relu || fmaf(fa.z, v.z, fb.z) > 0.f) ? d[i].z : 0.f) - k1.z - (v.z - mu.z) * is.z * k2.z);
          o.w = fa.w * (((!relu || fmaf(fa.w, v.w, fb.w) > 0.f) ? d[i].w : 0.f) - k1.w - (v.w - mu.w) * is.w * k2.w);
          if (!colok || tile * BM + r >= p.P) o = make_float4(0.f, 0.f, 0.f, 0.f);  // rows / channels past the end
          o = tf32x4(o);
          *reinterpret_cast<float4 *>(slab + swz_base32(r, chunk)) = o;
          if (dgrad) *reinterpret_cast<float4 *>(stage + r * 128 + ((chunk ^ (r & 7)) << 4)) = o;
        }
        fence_proxy_async();
        mbar_arrive(&bars->dz_full[set][c]);
      } else {
        // ---------------- H_prev slab h: columns [32h, 32h+32) of the layer below ----------------
        const int h = j - nb, col = h * BK + chunk * 4;
        const bool colok = col < p.lda;
        float4 v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int grow = tile * BM + row0 + 16 * i;
          v[i] = (colok && grow < p.P) ? ldg4(p.Aprev + static_cast<size_t>(grow) * p.lda + col)
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        const bool fold = p.a_in != nullptr && col < p.K;
        const float4 fa = *reinterpret_cast<const float4 *>(u_a + col), fb = *reinterpret_cast<const float4 *>(u_b + col);
        mbar_wait(&bars->h_empty[grp], (h_use & 1) ^ 1);
        ++h_use;
        unsigned char *slab = s_h + grp * SLAB;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + 16 * i;
          float4 y = v[i];
          if (fold) {
            y.x = fmaf(fa.x, y.x, fb.x), y.y = fmaf(fa.y, y.y, fb.y), y.z = fmaf(fa.z, y.z, fb.z), y.w = fmaf(fa.w, y.w, fb.w);
            if (p.relu_in) y.x = fmaxf(y.x, 0.f), y.y = fmaxf(y.y, 0.f), y.z = fmaxf(y.z, 0.f), y.w = fmaxf(y.w, 0.f);
            if (tile * BM + r >= p.P) y = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          *reinterpret_cast<float4 *>(slab + swz_base32(r, chunk)) = tf32x4(y);
        }
        fence_proxy_async();
        mbar_arrive(&bars->h_full[grp]);
      }
      j += G;
      while (j >= ipt) j -= ipt, ++tl;
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =====================================
    if (lane == 0) {
      const uint32_t idesc1 = make_idesc(BM, p.kpad16);                                   // dH_prev = dZ * W
      const uint32_t idesc2 = make_idesc(BM, BK) | kIdescAMajorMN | kIdescBMajorMN;       // dW += dZ^T * H
      int w_use[kGroups] = {0, 0, 0}, h_use[kGroups] = {0, 0, 0};
      int g = 0;  // loader group of the next item (items are dealt round-robin)
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int acc = tl & 1, acc_use = tl >> 1;
        const int set = p.sets == 2 ? (tl & 1) : 0, use = p.sets == 2 ? (tl >> 1) : tl;
        const uint32_t dz_base = smem_u32(s_dz + set * nb * SLAB);
        if (dgrad) {
          mbar_wait(&bars->acc_empty[acc], (acc_use & 1) ^ 1);
          tc_fence_after();
        }
        for (int c = 0; c < nb; ++c) {
          mbar_wait(&bars->dz_full[set][c], use & 1);
          if (dgrad) {
            mbar_wait(&bars->w_full[g], w_use[g] & 1);
            ++w_use[g];
            tc_fence_after();
            const uint64_t da = make_desc(smem_u32(s_a1 + g * SLAB));
            const uint64_t db = make_desc(smem_u32(s_w + g * p.w_slot_bytes));
#pragma unroll
            for (int k = 0; k < BK / 8; ++k)
              tc_mma_tf32(tm_acc + acc * kb32, da + 2 * k, db + 2 * k, idesc1, (c | k) != 0);
            tc_commit(&bars->w_empty[g]);
          }
          if (++g == G) g = 0;
        }
        tc_fence_after();
        if (dgrad) tc_commit(&bars->acc_full[acc]);
        for (int h = 0; h < kb; ++h) {
          mbar_wait(&bars->h_full[g], h_use[g] & 1);
          ++h_use[g];
          tc_fence_after();
          const uint32_t hb = smem_u32(s_h + g * SLAB);
#pragma unroll 4
          for (int ks = 0; ks < BM / 8; ++ks)  // 8 rows (one 1024-byte swizzle atom) per instruction
            tc_mma_tf32(tm_dw + h * BK, make_desc_mn32(dz_base + ks * 1024, SLAB, 512),
                        make_desc_mn32(hb + ks * 1024, SLAB, 512), idesc2, (tl | ks) != 0);
          tc_commit(&bars->h_empty[g]);
          if (++g == G) g = 0;
        }
        tc_commit(&bars->dz_empty[tl & 3]);
      }
      tc_commit(&bars->dw_full);
    }
  } else {
    // ====================================== epilogue ======================================
    const int wq = warp;  // TMEM lane quarter
    float csum[kMaxKB], cq[kMaxKB];
#pragma unroll
    for (int h = 0; h < kMaxKB; ++h) csum[h] = 0.f, cq[h] = 0.f;
    if (dgrad) {
      for (int tl = 0; tl < my_tiles; ++tl) {
        const int tile = blockIdx.x + tl * gridDim.x, set = tl & 1, use = tl >> 1;
        const int grow = tile * BM + wq * 32 + lane;
        const bool rowok = grow < p.P;
        mbar_wait(&bars->acc_full[set], use & 1);
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < kMaxKB; ++h) {
          if (h < kb) {
            const int n0 = h * BK;
            float zh[32];          // zhat of the layer below where its ReLU passes the gradient
            uint32_t on = 0;       // bit jj: dY_prev[row, n0 + jj] = dH_prev (else 0)
            if (stats) {  // the previous layer's pre-BN output for this row (L2 hit: the loaders just read it)
#pragma unroll
              for (int jj = 0; jj < 32; jj += 4) {
                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                if (rowok && n0 + jj < p.lda) t = ldg4(p.Aprev + static_cast<size_t>(grow) * p.lda + n0 + jj);
                zh[jj] = t.x, zh[jj + 1] = t.y, zh[jj + 2] = t.z, zh[jj + 3] = t.w;
              }
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                const float y = fmaf(u_a[n0 + jj], zh[jj], u_b[n0 + jj]);
                if (rowok && n0 + jj < p.K && (!p.relu_in || y > 0.f)) on |= 1u << jj;
                zh[jj] = (zh[jj] - u_mu[n0 + jj]) * u_is[n0 + jj];
              }
            }
            float v[32];
            tc_ld32(tm_acc + (static_cast<uint32_t>(wq * 32) << 16) + set * kb32 + n0, v);
            if (rowok) {
              float *dst = p.dHprev + static_cast<size_t>(grow) * p.lda + n0;
              if (n0 + 32 <= p.lda && (p.lda & 7) == 0) {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 8) st_global_v8(dst + jj, v + jj);
              } else {
#pragma unroll
                for (int jj = 0; jj < 32; jj += 4)
                  if (n0 + jj < p.lda) *reinterpret_cast<float4 *>(dst + jj) = make_float4(v[jj], v[jj + 1], v[jj + 2], v[jj + 3]);
              }
            }
            if (stats) {
#pragma unroll
              for (int jj = 0; jj < 32; ++jj) {
                v[jj] = (on >> jj) & 1u ? v[jj] : 0.f;
                zh[jj] *= v[jj];
              }
              csum[h] += warp_colsum32(v, lane);
              cq[h] += warp_colsum32(zh, lane);
            }
          }
        }
        tc_fence_before();
        mbar_arrive(&bars->acc_empty[set]);
      }
    }
    if (stats) {
#pragma unroll
      for (int h = 0; h < kMaxKB; ++h)
        if (h < kb) {
          s_stats[(wq * 2 + 0) * kb32 + h * 32 + lane] = csum[h];
          s_stats[(wq * 2 + 1) * kb32 + h * 32 + lane] = cq[h];
        }
    }
    // dW: TMEM lanes = output channels of this layer, columns = channels of the layer below
    mbar_wait(&bars->dw_full, 0);
    tc_fence_after();
    if (wq * 32 < nb32) {
      float *dst = p.dWpart + (static_cast<size_t>(blockIdx.x) * nb32 + wq * 32 + lane) * kb32;
#pragma unroll
      for (int h = 0; h < kMaxKB; ++h)
        if (h < kb) {
          float v[32];
          if (my_tiles > 0) {
            tc_ld32(tm_dw + (static_cast<uint32_t>(wq * 32) << 16) + h * BK, v);
          } else {
#pragma unroll
            for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
          }
#pragma unroll
          for (int jj = 0; jj < 32; jj += 8) st_global_v8(dst + h * 32 + jj, v + jj);
        }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (stats && tid < kb32) {  // combine the four epilogue warps, fixed order
    float s = 0.f, q = 0.f;
    if (my_tiles > 0 && dgrad)
      for (int w = 0; w < kEpiWarps; ++w) s += s_stats[(w * 2) * kb32 + tid], q += s_stats[(w * 2 + 1) * kb32 + tid];
    float *dst = p.stats_part + static_cast<size_t>(blockIdx.x) * 2 * kb32;
    dst[tid] = s;
    dst[kb32 + tid] = q;
  }
  if (warp == kMmaWarp) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// W [N, K] row-major -> W^T tile images [nb][kpad16][32]: tile c holds rows k = 0..kpad16-1 (input
// channel), columns n = 32c..32c+31 (output channel), i.e. the K-major B operand of dH_prev = dZ * W;
// 16-byte chunks XOR-swizzled by (row & 7), rounded to TF32, zero padded.
__global__ void pack_weights_t_kernel(const float *__restrict__ W, int N, int K, int ldw, int nb, int kpad16,
                                      float *__restrict__ Wp) {
  const size_t total = static_cast<size_t>(nb) * kpad16 * BK;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % 4);
    const int pc = static_cast<int>((i / 4) % 8);
    const int r = static_cast<int>((i / BK) % kpad16);
    const int c = static_cast<int>(i / (static_cast<size_t>(BK) * kpad16));
    const int n = c * BK + ((pc ^ (r & 7)) * 4) + e;
    const float v = (n < N && r < K) ? W[static_cast<size_t>(n) * ldw + r] : 0.f;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    Wp[i] = __uint_as_float(t);
  }
}

struct Geom {
  int nb, kb, kpad16, groups, sets, w_slot_bytes, tmem_cols, grid;
  size_t smem;
};

inline bool geom(int p, int n, int k, int lda, bool dgrad, Geom *g) {
  if (n <= 0 || k <= 0 || n > kMaxNB * BK || lda > kMaxKB * BK || n % 4 != 0 || lda % 4 != 0 || lda < k) return false;
  g->nb = (n + BK - 1) / BK;
  g->kb = (lda + BK - 1) / BK;
  g->kpad16 = (lda + 15) / 16 * 16;
  g->w_slot_bytes = (g->kpad16 * BK * 4 + 1023) / 1024 * 1024;
  const int cols = (dgrad ? 3 : 1) * g->kb * BK;
  g->tmem_cols = 32;
  while (g->tmem_cols < cols) g->tmem_cols *= 2;
  if (g->tmem_cols > 512) return false;
  g->sets = g->nb <= 2 ? 2 : 1;
  const size_t fixed = static_cast<size_t>(g->sets) * g->nb * SLAB +
                       sizeof(float) * (6 * g->nb * 32 + 4 * g->kb * 32 + kEpiWarps * 2 * g->kb * 32) +
                       sizeof(Barriers) + 1024;
  const size_t per_group = 2 * static_cast<size_t>(SLAB) + g->w_slot_bytes;
  g->groups = kGroups;
  while (g->groups > 1 && fixed + g->groups * per_group > 227 * 1024) --g->groups;
  g->smem = fixed + g->groups * per_group;
  // GEMM 2 addresses four 32-channel slabs from the start of a dZ set whatever nb is: the slabs
  // past nb are other live shared memory (their rows of dW are never read) but must be mapped
  const size_t reach = static_cast<size_t>((g->sets - 1) * g->nb + kMaxNB) * SLAB;
  if (g->smem < reach) g->smem = reach;
  if (g->smem > 227 * 1024) return false;
  const int m_tiles = (p + BM - 1) / BM;
  g->grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
  return true;
}

}  // namespace bwd
}  // namespace pn2

using namespace pn2;

// 1 when pn2_mlp_bwd_layer supports the shape (N <= 128 output channels, lda <= 160 input columns)
extern "C" int pn2_mlp_bwd_supported(int p, int n, int k, int lda) {
  bwd::Geom g;
  return p > 0 && bwd::geom(p, n, k, lda, true, &g) ? 1 : 0;
}
extern "C" int pn2_mlp_bwd_grid(int p) {
  const int m_tiles = (p + bwd::BM - 1) / bwd::BM;
  return m_tiles < kNumSMs ? m_tiles : kNumSMs;
}
extern "C" int pn2_mlp_bwd_nb(int n) { return (n + bwd::BK - 1) / bwd::BK; }
extern "C" int pn2_mlp_bwd_kb(int lda) { return (lda + bwd::BK - 1) / bwd::BK; }
extern "C" int pn2_mlp_bwd_kpad(int lda) { return (lda + 15) / 16 * 16; }

extern "C" int pn2_mlp_pack_weights_t(int n, int k, int lda, int ldw, const float *w, float *packed, void *stream) {
  if (n <= 0 || k <= 0 || ldw < k || lda < k) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(w, packed);
  const int nb = pn2_mlp_bwd_nb(n), kpad16 = pn2_mlp_bwd_kpad(lda);
  const size_t total = static_cast<size_t>(nb) * kpad16 * bwd::BK;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  bwd::pack_weights_t_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, n, k, ldw, nb, kpad16, packed);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_mlp_bwd_layer(int p, int n, int k, int lda, int ns, int relu, int relu_in, const float *dh,
                                 const float *dout, const int *argmax, const float *z, const float *a,
                                 const float *b, const float *mean, const float *invstd, const float *c1c2,
                                 const float *a_prev, const float *a_in, const float *b_in,
                                 const float *mean_in, const float *invstd_in, const float *wt_packed,
                                 float *dh_prev, float *dw_part, float *stats_part, void *stream) {
  if (p <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(z, a, b, mean, invstd, c1c2, a_prev, dw_part);
  if (dh == nullptr && (dout == nullptr || argmax == nullptr || ns <= 0 || p % ns != 0)) return PN2_ERR_INVALID_ARG;
  if ((wt_packed == nullptr) != (dh_prev == nullptr)) return PN2_ERR_INVALID_ARG;
  if (stats_part != nullptr && (wt_packed == nullptr || a_in == nullptr || mean_in == nullptr || invstd_in == nullptr))
    return PN2_ERR_INVALID_ARG;
  if (a_in != nullptr && b_in == nullptr) return PN2_ERR_INVALID_ARG;
  bwd::Geom g;
  if (!bwd::geom(p, n, k, lda, wt_packed != nullptr, &g)) return PN2_ERR_UNSUPPORTED;
  if (!aligned16(z) || !aligned16(a_prev) || !aligned16(dw_part) || (dh != nullptr && !aligned16(dh)) ||
      (dout != nullptr && (!aligned16(dout) || !aligned16(argmax))) || (dh_prev != nullptr && !aligned16(dh_prev)) ||
      (wt_packed != nullptr && !aligned16(wt_packed)))
    return PN2_ERR_UNSUPPORTED;
  bwd::Params prm;
  prm.dH = dh, prm.dout = dout, prm.argmax = argmax, prm.Z = z;
  prm.a = a, prm.b = b, prm.mean = mean, prm.invstd = invstd, prm.c1c2 = c1c2;
  prm.Aprev = a_prev, prm.a_in = a_in, prm.b_in = b_in, prm.mean_in = mean_in, prm.invstd_in = invstd_in;
  prm.WTp = wt_packed, prm.dHprev = dh_prev, prm.dWpart = dw_part, prm.stats_part = stats_part;
  prm.P = p, prm.N = n, prm.nb = g.nb, prm.K = k, prm.lda = lda, prm.kb = g.kb, prm.kpad16 = g.kpad16;
  prm.ns = ns > 0 ? ns : 1, prm.relu = relu, prm.relu_in = relu_in;
  prm.groups = g.groups, prm.sets = g.sets, prm.w_slot_bytes = g.w_slot_bytes, prm.tmem_cols = g.tmem_cols;
  cudaError_t e = cudaFuncSetAttribute(bwd::mlp_bwd_layer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(g.smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  bwd::mlp_bwd_layer_kernel<<<g.grid, bwd::kThreads, g.smem, static_cast<cudaStream_t>(stream)>>>(prm);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
