// Shared-MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// One 1x1-conv layer of a PointNet++ shared MLP in rows layout
// (/root/reference/pointnet2/pytorch_utils.py:11-36: Conv2d 1x1 -> BatchNorm2d -> ReLU,
//  as used by pointnet2_modules.py:65-67 and :414):
//
//     Z[P, N] = act(A)[P, K] * W[N, K]^T           act(x) = relu(a_in * x + b_in) | x
//
// fused with what surrounds it in the reference:
//   * PROLOGUE  the PREVIOUS layer's BatchNorm fold + ReLU is applied to the A tile after it
//               lands in shared memory (the normalised activation H is never read back from
//               HBM; optionally it is written out once, for the weight-gradient GEMM);
//   * EPILOGUE  per-channel sum / sum-of-squares of Z (this layer's BatchNorm batch
//               statistics) are reduced from the accumulator while it is stored, so Z is not
//               re-read for them.
//
// Structure (persistent CTAs, 288 threads; two per SM when BN <= 128, else one):
//   warps 0-3  loaders: cp.async 16-byte chunks of A straight into the 128B-swizzled K-major
//              UMMA tile (3 stages in flight), then transform the landed stage in place
//              (fold + ReLU + round-to-TF32), fence.proxy.async, arrive on the stage's
//              mbarrier; lane 0 of warp 0 also issues ONE 1-D TMA bulk copy per stage for the
//              pre-packed weight tile (pack_weights_kernel lays W out as ready tile images);
//   warp  8    one elected thread issues tcgen05.mma.kind::tf32 (M=128, N=BN, K=8) into a
//              double-buffered TMEM accumulator and commits to the stage / accumulator
//              mbarriers;
//   warps 4-7  epilogue: tcgen05.ld 32 lanes x 32 columns; the warp's 32 x 32 block of Z is staged in
//              shared memory (128-byte rows, 16-byte chunks XOR-swizzled: conflict-free STS.128) and
//              written by ONE TMA tensor store (cp.async.bulk.tensor.2d, SWIZZLE_128B; rows past P and
//              columns past N are clipped by the tensor map) -- a thread storing its own row straight
//              to global memory costs the LSU ~48 wavefronts per STG.256, which made the stores 58 % of
//              this kernel's L1 data-pipe traffic (ncu: l1tex data-pipe 75 % of peak, its limiter);
//              butterfly transpose-reduction of the column sums.
//
// Precision: TF32 operands (round-to-nearest), fp32 accumulation -- the arithmetic class of
// the reference's own convolutions (torch.backends.cudnn.allow_tf32 defaults to True).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pn2 {
namespace gemm {

constexpr int BM = 128;          // rows (positions) per tile = UMMA M
constexpr int BK = 32;           // fp32 per K block = one 128-byte swizzle row
constexpr int kMaxStages = 12;   // stage count is chosen per launch: as many as fit (see launcher)
constexpr int kLoadWarps = 4, kEpiWarps = 4;
constexpr int kThreads = 32 * (kLoadWarps + kEpiWarps + 2);   // + MMA warp + weight-tile producer warp
constexpr int kMaxBN = 256;
constexpr int A_STAGE_BYTES = BM * BK * 4;        // 16 KB
constexpr int kSmemBudget = 200 * 1024;           // A + B stages, one CTA per SM (BN > 128)
constexpr int kSmemBudgetHalf = 96 * 1024;        // two CTAs per SM (BN <= 128: 256 TMEM columns each)
constexpr int kOutStageBytes = 32 * 128;          // one epilogue warp's 32 rows x 32 fp32 block of Z
constexpr int kOutBytes = kEpiWarps * kOutStageBytes;

struct Params {
  const float *A;        // [P, lda]
  const float *a_in;     // [>= K] fold scale of the previous BatchNorm (nullptr: identity)
  const float *b_in;     // [>= K]
  float *H_out;          // optional [P, lda]: act(A) written once (fp32, not TF32-rounded)
  const float *Wp;       // packed weight tiles [n_tiles][k_blocks][BN][32] (swizzled, TF32)
  float *Z;              // [P, ldz] (nullptr: not stored)
  float *partials;       // [gridDim.x][2][n_tiles * BN] column sums / sums of squares, or nullptr
  int P, lda, K, k_blocks;
  int N, ldz, BN, n_tiles;
  int relu_in;
  int tma_store;              // Z leaves through the tensor map (else per-thread row stores)
  int tma_load;               // A tiles arrive by tensor loads (else cp.async chunks)
  // group extrema for the pooled last layer of an SA block (pool_ns = 16 or 32 rows per group, 0 = none):
  // max / min of Z over every group and the first row attaining them, [P / pool_ns, N] each; the sign of
  // the BatchNorm scale (known only after this kernel) decides which one max-pooling needs
  int pool_ns;
  float *zmax, *zmin;
  int *imax, *imin;
  int stages, b_stage_bytes;  // A pipeline depth; bytes reserved per weight slot (multiple of 1024)
  int b_slots;                // weight-tile ring: its own, shorter pipeline (the tiles come from L2, not HBM)
  int resident;               // b_slots == k_blocks: every weight tile of this CTA's N tile is loaded once
  int tmem_cols;              // 512 (one CTA per SM) or 256 (two): two accumulators of tmem_cols / 2
  // dropout between the previous layer's activation and this layer (nn.Dropout in the classifier,
  // pointnet2_semseg.py:156): act(x) is multiplied by drop[p, k] ? drop_scale : 0 on the A tile
  const unsigned char *drop;  // [P, K] keep mask (nullptr: none); requires a_in
  float drop_scale;
  int mask_off;               // byte offset of the A stages' dropout mask tiles ([stages][128 rows][32 bytes])
  // Data-gradient use (A = dZ of a layer, Z = dH_prev): BatchNorm-adjoint statistics of the layer BELOW from the
  // epilogue -- column sums of dY = relu'(a z + b) .* dH_prev and of dY .* zhat over this CTA's rows, written to
  // `partials` like the forward statistics (pn2_bwd_stats_finalize turns them into dbeta, dgamma, c1, c2).
  // zs: that layer's pre-BN output [P, ldzs]; zs_fold: its a | b | mean | invstd ([4][N]); nullptr: off.
  const float *zs, *zs_fold;
  int ldzs, zs_relu;
};

using namespace tc;

struct __align__(16) Barriers {
  uint64_t raw[kMaxStages];    // cp.async chunks of A landed            (128 arrivals)
  uint64_t full[kMaxStages];   // A transformed + weight tile landed     (128 arrivals + 1 expect_tx)
  uint64_t empty[kMaxStages];  // MMAs that read the stage have retired  (1 commit)
  uint64_t acc_full[2];    // accumulator complete                   (1 commit)
  uint64_t acc_empty[2];   // accumulator drained by the epilogue    (128 arrivals)
  uint64_t b_full[kMaxStages];   // weight tile landed in its slot         (1 arrival + tx bytes)
  uint64_t b_empty[kMaxStages];  // MMAs that read the slot have retired   (1 commit)
  uint32_t tmem_base;
};

// kDrop: dropout keep mask on the A tile.  kEpi: what the epilogue does beside storing Z -- 0 column statistics (or
// nothing), 1 group extrema of the pooled layer, 2 adjoint statistics of the layer below.  Compile-time so that each
// instantiation carries only its own epilogue (ncu put 17 % of the plain layer's stall samples on instruction fetch).
// kTma: A tiles by TMA tensor loads AND Z by TMA tensor stores (the common case; the cp.async / st.global paths
// are then not compiled in).
// kVar: 1 = a layer (fold + ReLU on the A tile, column statistics, no H side output), 2 = a plain product (no fold,
// no statistics: the lead products over the source points and the data gradients), 0 = read the flags at run time.
// kChunks: 32-column blocks of the N tile (BN / 32) as a compile-time constant for the hot widths (0 = run time, up
// to kMaxBN / 32): the epilogue is unrolled over them, and a 32-wide layer then carries one copy of it, not eight.
template <bool kDrop, int kEpi, bool kTma, int kVar, int kChunks>
__global__ void __launch_bounds__(kThreads, 2) gemm_tf32_fwd_kernel(Params p, const __grid_constant__ CUtensorMap tm_z,
                                                                     const __grid_constant__ CUtensorMap tm_a) {
  pdl_enter();
  extern __shared__ __align__(1024) unsigned char smem[];
  const int STAGES = p.stages, B_STAGE_BYTES = p.b_stage_bytes;
  // The weight tiles have their OWN ring (b_slots, usually 2): they are re-read from L2 for every row tile and
  // need little run-ahead, while the activation tiles stream from HBM and need all the depth shared memory
  // allows -- with one ring for both, a 128-wide layer had 2 stages of (16 + 16) KB and ONE activation tile in
  // flight per CTA.  `resident`: b_slots == k_blocks, every tile loaded once.
  const int b_slots = p.b_slots;
  unsigned char *smem_a = smem;                                  // STAGES x 16 KB
  unsigned char *smem_b = smem + STAGES * A_STAGE_BYTES;         // b_slots x b_stage_bytes
  unsigned char *smem_mask = smem + p.mask_off;                  // STAGES x 4 KB (dropout only)
  unsigned char *smem_out = smem_b + b_slots * B_STAGE_BYTES + (kDrop ? STAGES * BM * BK : 0);  // kEpiWarps x 4 KB
  Barriers *bars = reinterpret_cast<Barriers *>(smem_out + kOutBytes);
  float *s_stats = reinterpret_cast<float *>(bars + 1);          // [kEpiWarps][2][kMaxBN]
  float *s_fold = s_stats + kEpiWarps * 2 * kMaxBN;              // [2][k_blocks * BK]: a_in, b_in of the fold
  float *s_zs = s_fold + 2 * p.k_blocks * BK;                    // [4][kMaxBN]: a | b | mean | invstd of the layer below

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int n_tile = blockIdx.y;
  const int m_tiles = (p.P + BM - 1) / BM;
  const int my_tiles = blockIdx.x < m_tiles ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  const uint32_t b_bytes = static_cast<uint32_t>(p.BN) * BK * 4;

  if (tid == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&bars->raw[s], (kTma || p.tma_load) ? 1 : 32 * kLoadWarps);
      mbar_init(&bars->full[s], 32 * kLoadWarps);
      mbar_init(&bars->empty[s], 1);
    }
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&bars->b_full[s], 1);
      mbar_init(&bars->b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&bars->acc_full[a], 1);
      mbar_init(&bars->acc_empty[a], 32 * kEpiWarps);
    }
    mbar_fence_init();
  }
  if (warp == kLoadWarps + kEpiWarps) {  // the MMA warp owns the tensor memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&bars->tmem_base)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  for (int i = tid; i < p.k_blocks * BK; i += kThreads) {
    const bool ok = p.a_in != nullptr && i < p.K;
    s_fold[i] = ok ? p.a_in[i] : 1.f;
    s_fold[p.k_blocks * BK + i] = ok ? p.b_in[i] : 0.f;
  }
  if (kEpi == 2)
    for (int i = tid; i < 4 * p.BN; i += kThreads) {
      const int which = i / p.BN, col = n_tile * p.BN + (i - which * p.BN);
      s_zs[which * kMaxBN + (i - which * p.BN)] = col < p.N ? p.zs_fold[static_cast<size_t>(which) * p.N + col] : 0.f;
    }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  const int acc_cols = p.tmem_cols / 2;  // column stride between the two accumulators

  const int total_kb = my_tiles * p.k_blocks;  // K blocks this CTA streams, over all its tiles

  if (warp < kLoadWarps) {
    // =============================== loaders / transformers ===============================
    const int chunk = tid & 7;            // 16-byte chunk of the 128-byte K-block row
    const int row0 = tid >> 3;            // rows row0 + 16*i, i = 0..7
    // (stage, use parity, tile, k block) of the next K block to ISSUE and the next to CONSUME are
    // tracked incrementally (no divisions in the loop)
    int is_s = 0, is_par = 1, is_tile = blockIdx.x, is_kb = 0;
    auto issue = [&]() {                  // cp.async of the next K block into stage is_s
      const int s = is_s, tile = is_tile, kb = is_kb;
      mbar_wait(&bars->empty[s], is_par);
      if (kTma || (!kDrop && p.tma_load)) {
        // the whole 128 x 32 K block by ONE tensor load: same swizzled image as the cp.async path writes,
        // rows past P / columns past lda zero-filled by the TMA unit, no LSU wavefronts, no address math
        if (tid == 0) {
          mbar_expect_tx(&bars->raw[s], A_STAGE_BYTES);
          tma_load_box(&tm_a, smem_a + s * A_STAGE_BYTES, kb * BK, tile * BM, &bars->raw[s]);
        }
        if (++is_s == STAGES) is_s = 0, is_par ^= 1;
        if (++is_kb == p.k_blocks) is_kb = 0, is_tile += gridDim.x;
        return;
      }
      const int col = kb * BK + chunk * 4;
      const uint32_t dst0 = smem_u32(smem_a + s * A_STAGE_BYTES);
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const int r = row0 + 16 * i;
        const int grow = tile * BM + r;
        const bool ok = grow < p.P && col < p.lda;
        const float *src = ok ? p.A + static_cast<size_t>(grow) * p.lda + col : p.A;
        cp_async16(dst0 + r * 128 + ((chunk ^ (r & 7)) << 4), src, ok ? 16u : 0u);
        if (kDrop) {  // this thread's 4 mask bytes of the row ([128 rows][32 bytes] per stage)
          const bool mok = grow < p.P && col < p.K;
          cp_async4(smem_u32(smem_mask + s * BM * BK) + r * 32 + chunk * 4,
                    mok ? p.drop + static_cast<size_t>(grow) * p.K + col : p.drop, mok ? 4u : 0u);
        }
      }
      cp_async_arrive(&bars->raw[s]);
      if (++is_s == STAGES) is_s = 0, is_par ^= 1;
      if (++is_kb == p.k_blocks) is_kb = 0, is_tile += gridDim.x;
    };
    // Run AHEAD K blocks in front of the consumer, but refill only stages whose MMAs were issued
    // >= 3 iterations ago: the empty-barrier wait then never sits on the critical path (with
    // AHEAD = STAGES-1 every iteration would include a full loader -> MMA -> commit round trip).
    const int AHEAD = STAGES > 4 ? STAGES - 3 : STAGES - 1;
    int issued = 0;
    for (; issued < min(AHEAD, total_kb); ++issued) issue();
    int s = 0, par = 0, tile = blockIdx.x, kb = 0;
    for (int it = 0; it < total_kb; ++it) {
      if (issued < total_kb) issue(), ++issued;
      const int col = kb * BK + chunk * 4;
      mbar_wait(&bars->raw[s], par);
      // fold constants from shared memory (staged once per CTA: as per-stage __ldg they were this
      // loop's top stall, a global-load latency in front of every transform)
      const bool fold = kVar != 2 && (kVar == 1 || p.a_in != nullptr) && col < p.K;
      const float4 fa = *reinterpret_cast<const float4 *>(s_fold + col);
      const float4 fb = *reinterpret_cast<const float4 *>(s_fold + p.k_blocks * BK + col);
      unsigned char *base = smem_a + s * A_STAGE_BYTES;
      // The operand stays plain fp32 in shared memory: kind::tf32 reads the upper 19 bits of each
      // word (truncation, what the library TF32 GEMMs do with fp32 inputs).  Without a fold (a block's
      // first layer) the landed tile is therefore used as it is -- no pass over it at all; rows past
      // P were zero-filled by cp.async.
      if (fold) {
        const float2 a01 = make_float2(fa.x, fa.y), a23 = make_float2(fa.z, fa.w);
        const float2 b01 = make_float2(fb.x, fb.y), b23 = make_float2(fb.z, fb.w);
        const float lo = p.relu_in ? 0.f : -INFINITY;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int r = row0 + 16 * i;
          float4 *q = reinterpret_cast<float4 *>(base + r * 128 + ((chunk ^ (r & 7)) << 4));
          float4 v = *q;
          const float2 y01 = fma2(a01, make_float2(v.x, v.y), b01), y23 = fma2(a23, make_float2(v.z, v.w), b23);
          v = make_float4(fmaxf(y01.x, lo), fmaxf(y01.y, lo), fmaxf(y23.x, lo), fmaxf(y23.y, lo));
          if (kDrop) {
            const uint32_t m = *reinterpret_cast<const uint32_t *>(smem_mask + s * BM * BK + r * 32 + chunk * 4);
            v.x = (m & 0xffu) ? v.x * p.drop_scale : 0.f, v.y = (m & 0xff00u) ? v.y * p.drop_scale : 0.f;
            v.z = (m & 0xff0000u) ? v.z * p.drop_scale : 0.f, v.w = (m & 0xff000000u) ? v.w * p.drop_scale : 0.f;
          }
          const int grow = tile * BM + r;
          if (grow >= p.P) v = make_float4(0.f, 0.f, 0.f, 0.f);  // rows past the end stay zero: the
          else if (kVar == 0 && p.H_out != nullptr && n_tile == 0)  // epilogue statistics rely on it
            *reinterpret_cast<float4 *>(p.H_out + static_cast<size_t>(grow) * p.lda + col) = v;
          *q = v;
        }
      }
      fence_proxy_async();  // generic-proxy writes -> visible to the tensor core's async proxy
      mbar_arrive(&bars->full[s]);
      if (++s == STAGES) s = 0, par ^= 1;
      if (++kb == p.k_blocks) kb = 0, tile += gridDim.x;
    }
  } else if (warp == kLoadWarps + kEpiWarps) {
    // ===================================== MMA issuer =====================================
    {  // the whole warp runs the loop (warp-uniform control flow keeps the descriptors in uniform
       // registers); lane 0 issues
      const bool leader = lane == 0;
      const uint32_t idesc = make_idesc(BM, p.BN);
      const uint64_t k_hi = make_desc(0);
      int s = 0, par = 0, t = 0, kb = 0;
      int bs = 0, bpar = 0;   // weight ring slot / parity (resident: slot = K block, loaded once, parity stays 0)
      for (int it = 0; it < total_kb; ++it) {
        const int acc = t & 1;
        if (kb == 0) {
          mbar_wait(&bars->acc_empty[acc], ((t >> 1) & 1) ^ 1);
          tc_fence_after();
        }
        const int wslot = p.resident ? kb : bs;
        mbar_wait(&bars->b_full[wslot], p.resident ? 0 : bpar);
        mbar_wait(&bars->full[s], par);
        tc_fence_after();
        const uint64_t da = k_hi | static_cast<uint64_t>((smem_u32(smem_a + s * A_STAGE_BYTES) & 0x3ffff) >> 4);
        const uint64_t db = k_hi | static_cast<uint64_t>((smem_u32(smem_b + wslot * B_STAGE_BYTES) & 0x3ffff) >> 4);
        if (leader) {
          tc_mma_tf32(tmem_base + acc * acc_cols, da, db, idesc, kb != 0);
#pragma unroll
          for (int k = 1; k < BK / 8; ++k)  // 32 bytes (8 x tf32) further along K per instruction
            tc_mma_tf32(tmem_base + acc * acc_cols, da + 2 * k, db + 2 * k, idesc, 1);
          tc_commit(&bars->empty[s]);                       // stage reusable once these MMAs retire
          if (!p.resident) tc_commit(&bars->b_empty[wslot]);
          if (kb == p.k_blocks - 1) tc_commit(&bars->acc_full[acc]);
        }
        __syncwarp();
        if (++s == STAGES) s = 0, par ^= 1;
        if (++bs == b_slots) bs = 0, bpar ^= 1;
        if (++kb == p.k_blocks) kb = 0, ++t;
      }
    }
  } else if (warp == kLoadWarps + kEpiWarps + 1) {
    // ================================ weight-tile producer (one thread) ================================
    if (lane == 0) {
      const int total = p.resident ? min(total_kb, p.k_blocks) : total_kb;
      int bs = 0, bpar = 1, kb = 0;
      for (int q = 0; q < total; ++q) {
        if (!p.resident) mbar_wait(&bars->b_empty[bs], bpar);
        mbar_expect_tx(&bars->b_full[bs], b_bytes);
        tma_load_1d(smem_b + bs * B_STAGE_BYTES, p.Wp + (static_cast<size_t>(n_tile) * p.k_blocks + kb) * p.BN * BK,
                    b_bytes, &bars->b_full[bs]);
        if (++bs == b_slots) bs = 0, bpar ^= 1;
        if (++kb == p.k_blocks) kb = 0;
      }
    }
  } else {
    // ====================================== epilogue ======================================
    const int wq = warp & 3;  // TMEM lane quarter this warp may read
    constexpr int kCh = kChunks ? kChunks : kMaxBN / 32;
    const int chunks = kChunks ? kChunks : p.BN / 32 + (p.BN % 32 ? 1 : 0);
    float csum[kCh], csq[kCh];
#pragma unroll
    for (int c = 0; c < kCh; ++c) csum[c] = 0.f, csq[c] = 0.f;
    for (int t = 0; t < my_tiles; ++t) {
      const int tile = blockIdx.x + t * gridDim.x;
      const int acc = t & 1;
      mbar_wait(&bars->acc_full[acc], (t >> 1) & 1);
      tc_fence_after();
      const int grow = tile * BM + wq * 32 + lane;
#pragma unroll
      for (int c = 0; c < kCh; ++c) {
        if (c < chunks) {
          float v[32];
          tc_ld32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + acc * acc_cols + c * 32, v);
          const int n0 = n_tile * p.BN + c * 32;
          if (p.Z != nullptr && (kTma || p.tma_store)) {
            tma_store_block32(&tm_z, smem_out + wq * kOutStageBytes, v, lane, n0, tile * BM + wq * 32);
          } else if (p.Z != nullptr && grow < p.P) {
            float *dst = p.Z + static_cast<size_t>(grow) * p.ldz + n0;
            if (n0 + 32 <= p.N && (p.ldz & 7) == 0) {  // the thread's 128-byte row segment: 4 x STG.256
#pragma unroll
              for (int j = 0; j < 32; j += 8) st_global_v8(dst + j, v + j);
            } else
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              if (n0 + j + 3 < p.N) {
                *reinterpret_cast<float4 *>(dst + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
              } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                  if (n0 + j + e < p.N) dst[j + e] = v[j + e];
              }
          }
          if (kEpi == 1) {
            // pooled layer: column pass over the staged block (lane = column; the swizzle makes every row
            // read conflict-free) for the group extrema and this layer's BatchNorm sums.  (For the other
            // layers the shuffle butterfly below is cheaper than 32 more shared-memory wavefronts per
            // block: measured +93 us per step with the column pass everywhere.)
            const unsigned char *stg = smem_out + wq * kOutStageBytes + ((lane & 3) << 2);
            const int cj = lane >> 2;
            const int halves = 32 / p.pool_ns;  // 1 or 2 groups in the block
            float s0 = 0.f, q0 = 0.f;
            for (int hh = 0; hh < halves; ++hh) {
              float mx = -INFINITY, mn = INFINITY;
              int ix = 0, in = 0;
#pragma unroll 8
              for (int rr = 0; rr < p.pool_ns; ++rr) {
                const int r = hh * p.pool_ns + rr;
                const float x = *reinterpret_cast<const float *>(stg + r * 128 + ((cj ^ (r & 7)) << 4));
                s0 += x, q0 = fmaf(x, x, q0);
                if (x > mx) mx = x, ix = rr;
                if (x < mn) mn = x, in = rr;
              }
              const long long grp = (static_cast<long long>(tile) * BM + wq * 32) / p.pool_ns + hh;
              if (grp * p.pool_ns < p.P && n0 + lane < p.N) {
                const size_t o = static_cast<size_t>(grp) * p.N + n0 + lane;
                p.zmax[o] = mx, p.zmin[o] = mn, p.imax[o] = ix, p.imin[o] = in;
              }
            }
            csum[c] += s0, csq[c] += q0;
          } else if (kEpi == 2) {
            // adjoint statistics of the layer below: dY = dH_prev where that layer's ReLU passed, zhat from its z
            float dy[32], dq[32];
            const bool rok = grow < p.P;
            const float *zr = p.zs + static_cast<size_t>(rok ? grow : 0) * p.ldzs + n0;
#pragma unroll
            for (int j4 = 0; j4 < 8; ++j4) {
              const bool in = rok && n0 + 4 * j4 < p.N;
              const float4 zz = in ? __ldg(reinterpret_cast<const float4 *>(zr + 4 * j4)) : make_float4(0.f, 0.f, 0.f, 0.f);
              const float zv[4] = {zz.x, zz.y, zz.z, zz.w};
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                const int j = 4 * j4 + e, cc = c * 32 + j;
                const float ya = fmaf(s_zs[cc], zv[e], s_zs[kMaxBN + cc]);
                const float d = (in && (!p.zs_relu || ya > 0.f)) ? v[j] : 0.f;
                dy[j] = d;
                dq[j] = d * ((zv[e] - s_zs[2 * kMaxBN + cc]) * s_zs[3 * kMaxBN + cc]);
              }
            }
            csum[c] += warp_colsum32(dy, lane);
            csq[c] += warp_colsum32(dq, lane);
          } else if (kVar != 2 && (kVar == 1 || p.partials != nullptr)) {  // rows past P are zero (zero-filled A)
            float sq[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
            csum[c] += warp_colsum32(v, lane);
            csq[c] += warp_colsum32(sq, lane);
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&bars->acc_empty[acc]);
    }
    tma_store_drain(lane);
    if (kVar != 2 && p.partials != nullptr) {
#pragma unroll
      for (int c = 0; c < kCh; ++c)
        if (c < chunks) {
          s_stats[(wq * 2 + 0) * kMaxBN + c * 32 + lane] = csum[c];
          s_stats[(wq * 2 + 1) * kMaxBN + c * 32 + lane] = csq[c];
        }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (kVar != 2 && p.partials != nullptr && tid < p.BN) {  // combine the four epilogue warps, fixed order
    const int ntot = p.n_tiles * p.BN;
    float s = 0.f, q = 0.f;
    if (my_tiles > 0)
      for (int w = 0; w < kEpiWarps; ++w) s += s_stats[(w * 2) * kMaxBN + tid], q += s_stats[(w * 2 + 1) * kMaxBN + tid];
    float *dst = p.partials + static_cast<size_t>(blockIdx.x) * 2 * ntot;
    dst[n_tile * p.BN + tid] = s;
    dst[ntot + n_tile * p.BN + tid] = q;
  }
  if (warp == kLoadWarps + kEpiWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// W [N, K] row-major (a 1x1 conv weight) -> tile images [n_tiles][k_blocks][BN][32], each row's
// 16-byte chunks XOR-swizzled by (row & 7) exactly as the UMMA SWIZZLE_128B layout expects,
// rounded to TF32, zero padded.
// (sn, sk): strides of W along the output / reduction index -- (ldw, 1) for the forward product,
// (1, ldw) for the transposed weight of the data-gradient product dH_prev = dZ * W.
// (n0, k0): output / reduction indices below them read as zero weights (the coordinate columns of a layer whose
// coordinate term is evaluated elsewhere, mlp_lead.cu).
__global__ void pack_weights_kernel(const float *__restrict__ W, int N, int K, int sn, int sk, int BN, int n_tiles,
                                    int k_blocks, float *__restrict__ Wp, int n0 = 0, int k0 = 0) {
  pdl_enter();
  const size_t total = static_cast<size_t>(n_tiles) * k_blocks * BN * BK;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % 4);
    const int pc = static_cast<int>((i / 4) % 8);  // physical chunk
    const int r = static_cast<int>((i / BK) % BN);
    const size_t blk = i / (static_cast<size_t>(BK) * BN);
    const int kb = static_cast<int>(blk % k_blocks), nt = static_cast<int>(blk / k_blocks);
    const int n = nt * BN + r;
    const int k = kb * BK + ((pc ^ (r & 7)) * 4) + e;
    float v = (n >= n0 && n < N && k >= k0 && k < K) ? W[static_cast<size_t>(n) * sn + static_cast<size_t>(k) * sk] : 0.f;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    Wp[i] = __uint_as_float(t);
  }
}

// UMMA N per CTA: a multiple of 16 in [16, 256].  Few row tiles (the coarse levels: P = 2048 ... 65536)
// would leave most SMs idle with one wide N tile per row tile, so N is split further until the grid
// covers the machine (the A tile is then re-read from L2 by each N tile's CTA).
inline int pick_bn(int p, int n) {
  int bn = n > kMaxBN ? kMaxBN : (n + 15) / 16 * 16;
  // Wide layers: N tiles of <= 128 columns, two CTAs per SM, instead of one 256-wide tile per SM.  Measured
  // at 65536 x 128 -> 256: the 256-wide tile runs 10 us per row tile with 9 warps per SM, two 128-wide CTAs
  // 4.4 us per row tile each (FP0 shapes): warps to hide latency matter more than the A tile re-read from L2.
  static const bool wide_tiles = getenv("PN2_GEMM_WIDE_TILES") == nullptr || getenv("PN2_GEMM_WIDE_TILES")[0] != '0';
  if (n > 128 && wide_tiles) {
    const int nt = (n + 127) / 128;
    bn = ((n + nt - 1) / nt + 31) / 32 * 32;
  }
  const int m_tiles = (p + BM - 1) / BM;
  // (halves stay multiples of 32: the epilogue reads the accumulator out in 32-column blocks, and a block
  // that straddled the end of a 48- or 80-wide tile would overwrite the next tile's columns of Z)
  while (bn >= 64 && bn % 64 == 0 && m_tiles * ((n + bn - 1) / bn) < kNumSMs) bn /= 2;
  return bn;
}

}  // namespace gemm
}  // namespace pn2

using namespace pn2;

// geometry helpers for the caller (sizes of the packed weights / the partials buffer)
extern "C" int pn2_mlp_gemm_bn(int p, int n) { return gemm::pick_bn(p, n); }
extern "C" int pn2_mlp_gemm_n_tiles(int p, int n) { return (n + gemm::pick_bn(p, n) - 1) / gemm::pick_bn(p, n); }
extern "C" int pn2_mlp_gemm_k_blocks(int k) { return (k + gemm::BK - 1) / gemm::BK; }
extern "C" int pn2_mlp_gemm_grid(int p, int n) {
  const int m_tiles = (p + gemm::BM - 1) / gemm::BM;
  const int ctas_per_sm = gemm::pick_bn(p, n) <= 128 ? 2 : 1;
  const int per = ctas_per_sm * kNumSMs / pn2_mlp_gemm_n_tiles(p, n);
  return m_tiles < per ? m_tiles : (per < 1 ? 1 : per);
}

// k0 > 0: the first k0 reduction columns of W read as zeros (pn2_mlp_pack_weights_cols)
static int pack_fwd(int p, int n, int k, int k0, int ldw, const float *w, float *packed, void *stream) {
  if (p <= 0 || n <= 0 || k <= 0 || ldw < k || k0 < 0 || k0 > k) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(w, packed);
  const int bn = gemm::pick_bn(p, n), nt = pn2_mlp_gemm_n_tiles(p, n), kb = pn2_mlp_gemm_k_blocks(k);
  const size_t total = static_cast<size_t>(nt) * kb * bn * gemm::BK;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  pn2::launch(gemm::pack_weights_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), w, n, k, ldw, 1, bn, nt, kb, packed,
                                                                                  0, k0);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_mlp_pack_weights(int p, int n, int k, int ldw, const float *w, float *packed, void *stream) {
  return pack_fwd(p, n, k, 0, ldw, w, packed, stream);
}
extern "C" int pn2_mlp_pack_weights_cols(int p, int n, int k, int k0, int ldw, const float *w, float *packed,
                                         void *stream) {
  return pack_fwd(p, n, k, k0, ldw, w, packed, stream);
}

// Tile images of W^T for the data-gradient product dH_prev[p, n_out] = dZ[p, n] * W[n, k] run on
// pn2_mlp_gemm_fwd (a = dZ, reduction over the layer's n output channels): output column o < k takes
// W[:, o]; columns k <= o < n_out (the zero padding of the rows layout) get zero weights, so the
// product writes zeros there.  Same geometry as pn2_mlp_pack_weights(p, n_out, n).
static int pack_dgrad(int p, int n_out, int n, int k, int k0, int ldw, const float *w, float *packed, void *stream) {
  if (p <= 0 || n <= 0 || k <= 0 || ldw < k || n_out < k || k0 < 0 || k0 > k) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(w, packed);
  const int bn = gemm::pick_bn(p, n_out), nt = pn2_mlp_gemm_n_tiles(p, n_out), kb = pn2_mlp_gemm_k_blocks(n);
  const size_t total = static_cast<size_t>(nt) * kb * bn * gemm::BK;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  pn2::launch(gemm::pack_weights_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), w, k, n, 1, ldw, bn, nt, kb, packed,
                                                                                  k0, 0);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_mlp_pack_weights_dgrad(int p, int n_out, int n, int k, int ldw, const float *w, float *packed,
                                          void *stream) {
  return pack_dgrad(p, n_out, n, k, 0, ldw, w, packed, stream);
}
// ... with the first k0 OUTPUT columns written as zeros too (no gradient reaches the coordinate columns of a
// point table through this product: their term of the layer is differentiated in mlp_lead.cu)
extern "C" int pn2_mlp_pack_weights_dgrad_cols(int p, int n_out, int n, int k, int k0, int ldw, const float *w,
                                               float *packed, void *stream) {
  return pack_dgrad(p, n_out, n, k, k0, ldw, w, packed, stream);
}

extern "C" int pn2_mlp_gemm_pool_supported(int ns) { return (ns == 16 || ns == 32) ? 1 : 0; }

static int launch_gemm(int p, int k, int lda, int n, int ldz, const float *a, const float *a_in,
                       const float *b_in, int relu_in, float *h_out, const float *w_packed,
                       float *z, float *partials, const unsigned char *drop, float drop_scale,
                       int pool_ns, float *zmax, float *zmin, int *imax, int *imin, const float *zs,
                       const float *zs_fold, int ldzs, int zs_relu, void *stream) {
  if (p <= 0 || k <= 0 || n <= 0 || lda < 4 || lda % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(a, w_packed);
  if (!aligned16(a) || !aligned16(w_packed)) return PN2_ERR_UNSUPPORTED;
  if (z != nullptr && (!aligned16(z) || ldz % 4 != 0 || ldz < n)) return PN2_ERR_UNSUPPORTED;
  if (a_in != nullptr && (b_in == nullptr || !aligned16(a_in) || !aligned16(b_in))) return PN2_ERR_UNSUPPORTED;
  if (h_out != nullptr && !aligned16(h_out)) return PN2_ERR_UNSUPPORTED;
  gemm::Params prm;
  prm.A = a, prm.a_in = a_in, prm.b_in = b_in, prm.H_out = h_out, prm.Wp = w_packed, prm.Z = z;
  prm.partials = partials;
  prm.zs = zs, prm.zs_fold = zs_fold, prm.ldzs = ldzs, prm.zs_relu = zs_relu;
  if (zs != nullptr && (partials == nullptr || zs_fold == nullptr || pool_ns != 0 || ldzs < n || ldzs % 4 != 0 ||
                        !aligned16(zs)))
    return PN2_ERR_INVALID_ARG;
  prm.P = p, prm.lda = lda, prm.K = k, prm.k_blocks = pn2_mlp_gemm_k_blocks(k);
  prm.N = n, prm.ldz = ldz, prm.BN = gemm::pick_bn(p, n), prm.n_tiles = pn2_mlp_gemm_n_tiles(p, n);
  prm.relu_in = relu_in;
  // as many stages as fit: the narrow layers (BN = 32..64) are pure streaming and need >100 KB of
  // loads in flight per SM to cover HBM latency; the wide ones get 4 x 48 KB
  prm.b_stage_bytes = (prm.BN * gemm::BK * 4 + 1023) / 1024 * 1024;
  if (drop != nullptr && (a_in == nullptr || h_out != nullptr || k % 4 != 0)) return PN2_ERR_INVALID_ARG;
  prm.drop = drop, prm.drop_scale = drop_scale;
  const int mask_stage = drop != nullptr ? gemm::BM * gemm::BK : 0;   // [128 rows][32 mask bytes] per A stage
  // Narrow layers (BN <= 128) are pure streaming: two CTAs per SM (256 TMEM columns and <= 96 KB of
  // stages each) double the warps available to hide latency.
  const char *oenv = getenv("PN2_GEMM_ONE_PER_SM");   // timing experiment: N tiles of 128 as one CTA per SM
  const bool two_per_sm = prm.BN <= 128 && !(oenv != nullptr && oenv[0] == '1' && prm.BN == 128);
  prm.tmem_cols = two_per_sm ? 256 : 512;
  // the epilogue's staging blocks come out of the same budget (two CTAs of <= 113 KB share an SM)
  const int budget = (two_per_sm ? gemm::kSmemBudgetHalf : gemm::kSmemBudget) - gemm::kOutBytes;
  // Weight ring: all K blocks once (resident) when the whole set fits beside >= 3 activation stages and the
  // tiles are large (PN2_GEMM_RESIDENT = 0 / 1 forces it where it fits); else 2 slots (PN2_GEMM_BSLOTS).
  const int w_all = prm.k_blocks * prm.b_stage_bytes;
  const char *renv = getenv("PN2_GEMM_RESIDENT");
  const bool fits = budget - w_all >= 3 * (gemm::A_STAGE_BYTES + mask_stage) && prm.k_blocks <= gemm::kMaxStages;
  prm.resident = (fits && (renv != nullptr ? renv[0] == '1' : prm.b_stage_bytes >= gemm::A_STAGE_BYTES && !two_per_sm)) ? 1 : 0;
  const char *benv = getenv("PN2_GEMM_BSLOTS");
  prm.b_slots = prm.resident ? prm.k_blocks : (benv != nullptr ? atoi(benv) : 2);
  if (prm.b_slots < 1) prm.b_slots = 1;
  if (prm.b_slots > gemm::kMaxStages) prm.b_slots = gemm::kMaxStages;
  if (benv != nullptr && atoi(benv) == 0 && !prm.resident)   // experiment: one ring depth for both, as before the split
    prm.b_slots = budget / (gemm::A_STAGE_BYTES + mask_stage + prm.b_stage_bytes);
  prm.stages = (budget - prm.b_slots * prm.b_stage_bytes) / (gemm::A_STAGE_BYTES + mask_stage);
  if (benv != nullptr && atoi(benv) == 0 && !prm.resident) prm.stages = prm.b_slots;
  if (prm.stages > gemm::kMaxStages) prm.stages = gemm::kMaxStages;
  if (prm.stages < 2) return PN2_ERR_UNSUPPORTED;
  prm.mask_off = prm.stages * gemm::A_STAGE_BYTES + prm.b_slots * prm.b_stage_bytes;
  const size_t smem = static_cast<size_t>(prm.stages) * (gemm::A_STAGE_BYTES + mask_stage) +
                      static_cast<size_t>(prm.b_slots) * prm.b_stage_bytes + gemm::kOutBytes +
                      sizeof(gemm::Barriers) + sizeof(float) * gemm::kEpiWarps * 2 * gemm::kMaxBN +
                      sizeof(float) * 2 * prm.k_blocks * gemm::BK + (zs != nullptr ? sizeof(float) * 4 * gemm::kMaxBN : 0);
  CUtensorMap tm_z;
  memset(&tm_z, 0, sizeof(tm_z));
  static const bool no_tma = getenv("PN2_GEMM_NO_TMA_STORE") != nullptr;   // timing experiments only
  prm.tma_store = (z != nullptr && !no_tma && tc::make_tmap_rows32(&tm_z, z, p, n, ldz)) ? 1 : 0;
  CUtensorMap tm_a;
  memset(&tm_a, 0, sizeof(tm_a));
  static const bool no_tma_load = getenv("PN2_GEMM_NO_TMA_LOAD") != nullptr;  // timing experiments only
  prm.tma_load = (drop == nullptr && !no_tma_load && tc::make_tmap_rows32(&tm_a, a, p, lda, lda, gemm::BM)) ? 1 : 0;
  prm.pool_ns = pool_ns, prm.zmax = zmax, prm.zmin = zmin, prm.imax = imax, prm.imin = imin;
  if (pool_ns != 0) {
    if (!pn2_mlp_gemm_pool_supported(pool_ns) || p % pool_ns != 0) return PN2_ERR_UNSUPPORTED;
    PN2_CHECK_PTRS(zmax, zmin, imax, imin);
    if (!prm.tma_store) return PN2_ERR_UNSUPPORTED;  // the extrema are read from the staged output blocks
  }
  if (drop != nullptr && (pool_ns != 0 || zs != nullptr)) return PN2_ERR_UNSUPPORTED;
  const bool all_tma = prm.tma_load && (prm.tma_store || z == nullptr);
  const bool layer = a_in != nullptr && partials != nullptr && h_out == nullptr;   // fold + statistics
  const bool plain = a_in == nullptr && partials == nullptr && h_out == nullptr;   // bare product
  auto kern = gemm::gemm_tf32_fwd_kernel<false, 0, false, 0, 0>;
  const int ch = prm.BN % 32 == 0 ? prm.BN / 32 : 0;   // 32-column blocks of the N tile (0: partial block, run time)
#define PN2_GEMM_CH(E_, V_)                                                                     \
  (ch == 1 ? gemm::gemm_tf32_fwd_kernel<false, E_, true, V_, 1>                                  \
           : ch == 2 ? gemm::gemm_tf32_fwd_kernel<false, E_, true, V_, 2>                        \
                     : ch == 4 ? gemm::gemm_tf32_fwd_kernel<false, E_, true, V_, 4> : gemm::gemm_tf32_fwd_kernel<false, E_, true, V_, 0>)
  if (drop != nullptr)   // (the keep mask arrives by cp.async beside the A tile: no TMA variant)
    kern = (layer && ch == 1) ? gemm::gemm_tf32_fwd_kernel<true, 0, false, 1, 1> : gemm::gemm_tf32_fwd_kernel<true, 0, false, 0, 0>;
  else if (zs != nullptr) kern = gemm::gemm_tf32_fwd_kernel<false, 2, false, 0, 0>;
  else if (pool_ns != 0)
    kern = all_tma ? (layer ? PN2_GEMM_CH(1, 1) : gemm::gemm_tf32_fwd_kernel<false, 1, true, 0, 0>)
                   : gemm::gemm_tf32_fwd_kernel<false, 1, false, 0, 0>;
  else if (all_tma)
    kern = layer ? PN2_GEMM_CH(0, 1) : plain ? PN2_GEMM_CH(0, 2) : gemm::gemm_tf32_fwd_kernel<false, 0, true, 0, 0>;
#undef PN2_GEMM_CH
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(pn2_mlp_gemm_grid(p, n), prm.n_tiles);  // the partials buffer is sized by the same function
  pn2::launch(kern, grid, gemm::kThreads, smem, static_cast<cudaStream_t>(stream), prm, tm_z, tm_a);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_mlp_gemm_fwd(int p, int k, int lda, int n, int ldz, const float *a, const float *a_in,
                                const float *b_in, int relu_in, float *h_out, const float *w_packed,
                                float *z, float *partials, const unsigned char *drop, float drop_scale,
                                int pool_ns, float *zmax, float *zmin, int *imax, int *imin, void *stream) {
  return launch_gemm(p, k, lda, n, ldz, a, a_in, b_in, relu_in, h_out, w_packed, z, partials, drop, drop_scale, pool_ns,
                     zmax, zmin, imax, imin, nullptr, nullptr, 0, 0, stream);
}

// Data gradient of a wide layer, dh_prev[p, n_out] = dz[p, n] * W (tile images of pn2_mlp_pack_weights_dgrad), with the
// BatchNorm-adjoint statistics of the layer BELOW out of the epilogue: z_prev [p, n_out] is that layer's pre-BN output,
// fold_prev = its [a | b | mean | invstd] ([4][n_out]); partials: [pn2_mlp_gemm_grid(p, n_out)][2][n_tiles * bn] column
// sums of dY_prev and dY_prev .* zhat_prev, to be finalised by pn2_bwd_stats_finalize (ld = n_tiles * bn).  Removes the
// statistics pass over (dh_prev, z_prev) that pn2_relu_bn_bwd(_stats) would make.
extern "C" int pn2_mlp_dgrad_stats(int p, int n, int n_out, const float *dz, const float *w_packed, float *dh_prev,
                                   const float *z_prev, const float *fold_prev, int relu_prev, float *partials,
                                   void *stream) {
  if (z_prev == nullptr || fold_prev == nullptr || partials == nullptr || dh_prev == nullptr) return PN2_ERR_INVALID_ARG;
  return launch_gemm(p, n, n, n_out, n_out, dz, nullptr, nullptr, 0, nullptr, w_packed, dh_prev, partials, nullptr, 1.0f,
                     0, nullptr, nullptr, nullptr, nullptr, z_prev, fold_prev, n_out, relu_prev, stream);
}
