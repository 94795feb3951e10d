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
//   dW      = dZ^T * H_prev                   GEMM 2  (M = N, N = K, reduce over the rows)
//   s, q    = column sums of dY_prev, dY_prev * zhat_prev   (BN-adjoint constants of the layer below)
//
// dZ is formed on chip from (dH, Z) and never written; H_prev is recomputed from Z_prev (the
// forward pass does not store it); dW accumulates in tensor memory over all of a CTA's tiles and
// leaves the SM once.  HBM traffic per layer: read dH, Z, Z_prev, write dH_prev -- against 7
// reads/writes of [P, N] plus 2 of [P, K] for the unfused sequence.
//
// Statistics without an epilogue pass.  With M = relu mask of the layer below and
// H_prev = M .* (g1 * zhat_prev + g0)  (g1 = gamma, g0 = beta per column), GEMM 2 is run on the two
// operands M and M .* zhat_prev instead of H_prev:
//     S = dZ^T * M,   Q = dZ^T * (M .* zhat_prev)          (both [N, K], accumulated in TMEM)
//     dW = g1 .* Q + g0 .* S                                 (column scaling at the final read-out)
//     s[c] = sum_n W[n,c] S[n,c],   q[c] = sum_n W[n,c] Q[n,c]
// because dY_prev = M .* (dZ * W): the column sums the next adjoint needs are linear in S and Q.
// The per-tile epilogue is then only TMEM -> HBM of dH_prev.
//
// GEMM 2 reduces over ROWS, so its operands are read MN-major straight from row-major slabs
// ([128 rows][32 fp32], slab stride = leading byte offset): nothing is transposed in memory.
// MN-major TF32 operands must use the 32-byte-atom 128-byte swizzle, GEMM 1's K-major A operand the
// ordinary 16-byte-atom one, so a dZ slab is written twice by its loader: into the tile's dZ set
// (GEMM 2) and into the group's staging slot (GEMM 1, freed one step later).
//
// Structure (persistent CTAs, one per SM, 21 warps):
//   warps 0-3   epilogue: tcgen05.ld of dH_prev + row stores; at the end dW / statistics read-out
//   warp  4     one elected thread issues every tcgen05.mma (kind::tf32) and commit
//   warps 5-20  two loader groups of 256 threads taking the tile's slabs alternately, 4 float4 of a
//               slab per thread: 16-byte LDGs -> registers -> BN adjoint / fold + ReLU (packed fp32)
//               -> swizzled st.shared -> fence.proxy.async -> mbarrier arrive; W^T tiles by 1-D TMA.
//               One group's load latency overlaps the other's arithmetic.  (A single group with a
//               register prefetch of its next slab does not work: fence.proxy.async compiles to
//               MEMBAR.ALL.CTA, which waits for the prefetch loads -- measured 1.5 us per slab.)
//
// Precision: TF32 operands (round-to-nearest), fp32 accumulation -- the arithmetic class of the
// reference's own convolutions (torch.backends.cudnn.allow_tf32 defaults to True).
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

// Timing-experiment hooks of the backward layer (PN2_BWD_DBG: skip loads / stores, per-role wait cycles).  They are
// compiled in only with -DPN2_BWD_DEBUG: as run-time branches they were ~10 % of the kernel's instructions.
#ifdef PN2_BWD_DEBUG
#define PN2_DBG(p) ((p).dbg)
#else
#define PN2_DBG(p) 0
#endif

namespace pn2 {
namespace bwd {

using namespace tc;

constexpr int BM = 128;                 // rows per tile
constexpr int BK = 32;                  // channels per slab = one 128-byte swizzle row
constexpr int SLAB = BM * BK * 4;       // 16 KB
constexpr int kEpiWarps = 4, kGroups = 2, kGroupWarps = 8;
constexpr int kGroupThreads = 32 * kGroupWarps;  // 256: 4 float4 of a slab per thread
constexpr int kMmaWarp = kEpiWarps;     // warp 4
constexpr int kFirstLoadWarp = kEpiWarps + 1;
constexpr int kThreads = 32 * (kEpiWarps + 1 + kGroups * kGroupWarps);  // 672
constexpr int kMaxNB = 4;               // N <= 128
constexpr int kMaxKB = 5;               // lda <= 160 (<= 128 when statistics are wanted)
// Every loader group owns its H_prev slab(s), its dZ staging slab and its W^T tile slot: a slot then
// has a single, in-order producer, so its phase-parity waits can never alias (with a shared ring a
// group running two uses ahead of a lagging group would see the parity it expects and overwrite a
// live slab).

struct Params {
  // this layer, N output channels
  const float *dH;       // [P, N] upstream gradient wrt the layer's activated output (dense mode)
  const float *dout;     // [P / ns, N] gradient of the pooled output (pooled mode: dH == nullptr)
  const int *argmax;     // [P / ns, N] arg-max row of every (group, channel)
  const float *Z;        // [P, N] pre-BN output of this layer
  const float *a, *b, *mean, *invstd;  // [N] BN fold and statistics of this layer
  const float *c1c2;     // [2][N] dbeta / P, dgamma / P (0 in eval mode)
  const float *W;        // [N, K] this layer's weight, fp32 (statistics read-out)
  // the layer below, K channels
  const float *Aprev;    // [P, lda]: Z of the previous layer (a_in != nullptr) or the block's input rows
  const float *a_in, *b_in, *mean_in, *invstd_in;  // [K] fold / statistics of the previous layer
  const float *WTp;      // packed W^T tiles [nb][kpad16][32] (nullptr: dH_prev not wanted)
  float *dHprev;         // [P, lda]
  float *dWpart;         // [gridDim.x][nb * 32][kb * 32]: dW per CTA -- or, with a BN below, S = dZ^T M
  float *stats_part;     // [gridDim.x][nb * 32][kb * 32]: Q = dZ^T (M .* zhat_prev)  (nullptr: no BN below)
  int P, N, nb, K, lda, kb, kpad16;
  int ns, ns_shift, relu, relu_in;       // ns_shift >= 0 when ns is a power of two
  int sets, w_slot_bytes, tmem_cols;     // dZ sets (2, or 1 when N > 64)
  int out_bytes;                         // 16 KB of staging blocks when dH_prev leaves through the tensor map, else 0
  int dbg;                               // PN2_BWD_DBG timing experiments (results invalid when != 0)
  int *sched;                            // [2] {next tile, CTAs finished}, zero between launches (self-resetting)
  const unsigned char *drop_out, *drop_in;  // [P, N] / [P, K] dropout keep masks (nullptr: none)
  float drop_out_scale, drop_in_scale;
};

// PN2_BWD_DBG & 32: cycles each role spends in its waits, per CTA (read back with pn2_mlp_bwd_debug_read)
__device__ unsigned long long g_wait_cycles[kNumSMs][20];
__device__ __forceinline__ void timed_wait(uint64_t *bar, uint32_t parity, bool on, unsigned long long &acc) {
  if (!on) {
    mbar_wait(bar, parity);
    return;
  }
  const long long t0 = clock64();
  mbar_wait(bar, parity);
  acc += static_cast<unsigned long long>(clock64() - t0);
}

struct __align__(8) Barriers {
  uint64_t dz_full[2][kMaxNB];   // slab c of set s holds dZ           (256 arrivals: one loader group)
  uint64_t dz_empty[4];          // both products of tile t retired: barrier t & 3 (1 commit).  Four
                                 // barriers for <= two sets keep a loader at most one phase behind.
  uint64_t w_full[kGroups];      // W^T tile landed in the group's slot (1 arrival + tx bytes)
  uint64_t w_empty[kGroups];     // GEMM 1 step retired: W^T slot and dZ staging slot free (1 commit)
  uint64_t h_full[kGroups];      // H_prev slab(s) ready                (256 arrivals)
  uint64_t h_empty[kGroups];     // GEMM 2 on them retired              (1 commit)
  uint64_t acc_full[2];          // dH_prev accumulator complete        (1 commit)
  uint64_t acc_empty[2];         // drained by the epilogue             (128 arrivals)
  uint64_t dw_full;              // every MMA of the CTA retired        (1 commit)
  uint64_t tile_full[8];         // tile queue slot published           (1 arrival)
  int tile_id[8];                // row tile of the CTA's t-th tile, -1 = no more tiles
  int tiles_done;                // number of tiles this CTA processed (written by the MMA warp)
  uint32_t tmem_base;
};

__device__ __forceinline__ float4 ldg4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }
__device__ __forceinline__ float4 lds4(const float *p) { return *reinterpret_cast<const float4 *>(p); }
__device__ __forceinline__ float4 tf32x4(float4 v) {
  return make_float4(to_tf32(v.x), to_tf32(v.y), to_tf32(v.z), to_tf32(v.w));
}
__device__ __forceinline__ float4 zero4() { return make_float4(0.f, 0.f, 0.f, 0.f); }
// kDrop: dropout keep masks present (the classifier head's chain); compiled out otherwise -- the loader
// chain bounds this kernel and every instruction in it counts.
// kPooled: the pooled last layer of an SA block (gradient on the arg-max rows) -- compile-time, so that each
// instantiation carries one loader variant only.
// kFull: data gradient AND statistics of the layer below wanted (every layer but a block's first) -- the common
// case as compile-time constants; kFull = false reads the two flags from the parameters.
// NB / KB: the layer's 32-column slab counts (N / 32, K / 32 rounded up) as compile-time constants for the hot
// shapes (0 = read them from the parameters): the loaders' item bookkeeping and the slab loops then unroll.
template <bool kDrop, bool kPooled, bool kFull, int NB, int KB>
__global__ void __launch_bounds__(kThreads, 1) mlp_bwd_layer_kernel(Params p, const __grid_constant__ CUtensorMap tm_dh,
                                                                    const __grid_constant__ CUtensorMap tm_dw,
                                                                    const __grid_constant__ CUtensorMap tm_q) {
  pdl_enter();
  extern __shared__ __align__(1024) unsigned char smem[];
  const long long t_entry = (PN2_DBG(p) & 32) ? clock64() : 0;
  const int nb = NB ? NB : p.nb, kb = KB ? KB : p.kb;
  const int nb32 = nb * 32, kb32 = kb * 32;
  const bool dgrad = kFull || p.WTp != nullptr;
  const bool stats = kFull || p.stats_part != nullptr;
  const int hslabs = stats ? 2 : 1;                              // per group: H_prev, or M and M .* zhat
  unsigned char *s_dz = smem;                                    // [sets][nb] slabs, MN-major image (GEMM 2)
  unsigned char *s_h = s_dz + p.sets * nb * SLAB;                // [G][hslabs] slabs
  unsigned char *s_a1 = s_h + kGroups * hslabs * SLAB;           // [G] K-major dZ staging slabs (GEMM 1)
  unsigned char *s_w = s_a1 + kGroups * SLAB;                    // [G] W^T tiles
  unsigned char *s_out = s_w + kGroups * p.w_slot_bytes;         // [kEpiWarps] 4 KB blocks of dH_prev (TMA store)
  float *s_tab = reinterpret_cast<float *>(s_out + p.out_bytes);
  // dZ = a * dY + C * z + D  with  C = -a * invstd * c2,  D = -a * c1 - C * mean   (one FMA each)
  float *t_a = s_tab, *t_b = t_a + nb32, *t_C = t_b + nb32, *t_D = t_C + nb32, *t_x0 = t_D + nb32,
        *t_x1 = t_x0 + nb32;  // (t_x0 / t_x1 unused: keeps the table block's size)
  float *u_a = t_x1 + nb32, *u_b = u_a + kb32, *u_mu = u_b + kb32, *u_is = u_mu + kb32;
  float *s_stats = u_is + kb32;                                  // [kEpiWarps][2][kb32]
  Barriers *bars = reinterpret_cast<Barriers *>(s_stats + kEpiWarps * 2 * kb32);

  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = (p.P + BM - 1) / BM;
  // Row tiles are handed out DYNAMICALLY (one global atomic counter): the kernel usually shares the
  // GPU with the next batch's geometry kernels on another stream (FPS holds 32 SMs for ~0.7 ms), and
  // with a static round-robin the CTAs that cannot become resident delay their share of the tiles
  // until the first wave has finished -- up to 2x the kernel time.  The MMA warp fetches tile ids two
  // tiles ahead and publishes them in a shared-memory queue that the other roles read.
  auto tile_of = [&](int tl) -> int {
    mbar_wait(&bars->tile_full[tl & 7], (tl >> 3) & 1);
    return bars->tile_id[tl & 7];
  };

  if (tid == 0) {
    for (int s = 0; s < 2; ++s) {
      for (int c = 0; c < kMaxNB; ++c) mbar_init(&bars->dz_full[s][c], kGroupThreads);
      mbar_init(&bars->acc_full[s], 1);
      mbar_init(&bars->acc_empty[s], 32 * kEpiWarps);
    }
    for (int s = 0; s < 4; ++s) mbar_init(&bars->dz_empty[s], 1);
    for (int s = 0; s < kGroups; ++s) {
      mbar_init(&bars->w_full[s], 1), mbar_init(&bars->w_empty[s], 1);
      mbar_init(&bars->h_full[s], kGroupThreads), mbar_init(&bars->h_empty[s], 1);
    }
    mbar_init(&bars->dw_full, 1);
    for (int s = 0; s < 8; ++s) mbar_init(&bars->tile_full[s], 1);
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
    const bool ok = i < p.N;  // channels past N: all-zero constants give dZ = 0
    const float a = ok ? p.a[i] : 0.f;
    const float C = ok ? -a * p.invstd[i] * p.c1c2[p.N + i] : 0.f;
    t_a[i] = a, t_b[i] = ok ? p.b[i] : 0.f;
    t_C[i] = C, t_D[i] = ok ? -a * p.c1c2[i] - C * p.mean[i] : 0.f;
  }
  for (int i = tid; i < kb32; i += kThreads) {
    const bool ok = p.a_in != nullptr && i < p.K;
    u_a[i] = ok ? p.a_in[i] : 1.f, u_b[i] = ok ? p.b_in[i] : 0.f;
    u_mu[i] = ok && p.mean_in != nullptr ? p.mean_in[i] : 0.f;
    u_is[i] = ok && p.invstd_in != nullptr ? p.invstd_in[i] : 0.f;
  }
  // (zhat_prev = invstd_in * z + e with e = -mean_in * invstd_in is formed from u_mu / u_is by the loaders)
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;
  if ((PN2_DBG(p) & 32) && tid == 0) g_wait_cycles[blockIdx.x][16] = clock64() - t_entry;  // set-up
  // tensor memory: [acc 0 | acc 1] (kb32 columns each, only when dgrad), then dW [kb32] -- or,
  // with statistics, [S_h | Q_h] pairs of 32 columns per slab h of the layer below
  const uint32_t tm_acc = tmem_base, tm_dw_addr = tmem_base + (dgrad ? 2 * kb32 : 0);

  if (warp >= kFirstLoadWarp) {
    // ======================================= loaders =======================================
    const int lw = warp - kFirstLoadWarp, grp = lw / kGroupWarps;
    const int gt = (lw % kGroupWarps) * 32 + lane;  // 0..255 within the group
    const int chunk = gt & 7, row0 = (gt >> 3) * 4;  // 16-byte chunk of the slab row; rows row0 .. row0 + 3
    const int ipt = nb + kb;                          // items (slabs) per tile: dZ slabs, then H slabs
    uint32_t off16[4], off32[4];                      // byte offsets of the thread's 4 chunks in a slab
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      off16[i] = static_cast<uint32_t>((row0 + i) * 128 + ((chunk ^ ((row0 + i) & 7)) << 4));  // SWIZZLE_128B
      off32[i] = swz_base32(row0 + i, chunk);                                                    // .._BASE32B
    }
    int w_use = 0, h_use = 0;                        // uses of this group's own slots so far
    const bool tm = (PN2_DBG(p) & 32) != 0;
    unsigned long long c_w = 0, c_dz = 0, c_h = 0, c_ld = 0, c_tot = clock64();

    // global loads of item (tl, j) into a register buffer: b[0..3] = dH (or dOut) / Z_prev rows,
    // b[4..7] = Z rows (dZ items only).  Rows past P load as zeros.
    auto issue = [&](int tile, int j, float4 (&b)[8], uint32_t (&mk)[4]) {
      if (PN2_DBG(p) & 16) {  // timing experiment: no global loads
#pragma unroll
        for (int i = 0; i < 8; ++i) b[i] = zero4();
        return;
      }
      const int r0 = tile * BM + row0;
      const bool full = (tile + 1) * BM <= p.P;
      if (kDrop) {  // dropout keep-mask bytes of this thread's 4 columns (all-ones where there is no dropout)
        const unsigned char *dm = j < nb ? p.drop_out : p.drop_in;
        const int ld = j < nb ? p.N : p.K;
        const int col = (j < nb ? j : j - nb) * BK + chunk * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          mk[i] = (kDrop && dm != nullptr && col < ld && r0 + i < p.P)
                      ? __ldg(reinterpret_cast<const uint32_t *>(dm + static_cast<size_t>(r0 + i) * ld + col))
                      : 0x01010101u;
      }
      if (j < nb) {
        const int col = j * BK + chunk * 4;
        const bool colok = col < p.N;
        const float *zp = p.Z + static_cast<size_t>(r0) * p.N + col;
        if (!kPooled) {
          const float *dp = p.dH + static_cast<size_t>(r0) * p.N + col;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const bool ok = colok && (full || r0 + i < p.P);
            b[4 + i] = ok ? ldg4(zp + static_cast<size_t>(i) * p.N) : zero4();
            b[i] = ok ? ldg4(dp + static_cast<size_t>(i) * p.N) : zero4();
          }
        } else {  // pooled: the thread's 4 consecutive rows lie in ONE group (ns % 4 == 0)
#pragma unroll
          for (int i = 0; i < 4; ++i)
            b[4 + i] = (colok && (full || r0 + i < p.P)) ? ldg4(zp + static_cast<size_t>(i) * p.N) : zero4();
          const int g = p.ns_shift >= 0 ? (r0 >> p.ns_shift) : r0 / p.ns;
          const bool ok = colok && r0 < p.P;
          b[0] = ok ? ldg4(p.dout + static_cast<size_t>(g) * p.N + col) : zero4();
          b[1] = ok ? ldg4(reinterpret_cast<const float *>(p.argmax) + static_cast<size_t>(g) * p.N + col) : zero4();
        }
      } else {
        const int col = (j - nb) * BK + chunk * 4;
        const bool colok = col < p.lda;
        const float *ap = p.Aprev + static_cast<size_t>(r0) * p.lda + col;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          b[i] = (colok && (full || r0 + i < p.P)) ? ldg4(ap + static_cast<size_t>(i) * p.lda) : zero4();
      }
    };

    // Operands go to shared memory as plain fp32: kind::tf32 reads the upper 19 bits of each word
    // (truncation), as the library TF32 GEMMs do with fp32 inputs.  The arithmetic is Blackwell's
    // packed fp32 (FFMA2): the loaders' issue slots, not HBM, bound this kernel otherwise.
    auto process = [&](int tl, int tile, int j, float4 (&b)[8], const uint32_t (&mk)[4]) {
      const int r0 = tile * BM + row0;
      const bool full = (tile + 1) * BM <= p.P;
      if (j < nb) {
        // ---------------- dZ slab c: columns [32c, 32c+32) of this layer ----------------
        const int c = j, col = c * BK + chunk * 4;
        if (kPooled) {  // pooled layer: the gradient lives on the arg-max row of every (group, channel)
          const int g = p.ns_shift >= 0 ? (r0 >> p.ns_shift) : r0 / p.ns;
          const int s0 = r0 - g * p.ns;
          const float4 dd = b[0];
          const int4 am = make_int4(__float_as_int(b[1].x), __float_as_int(b[1].y), __float_as_int(b[1].z),
                                    __float_as_int(b[1].w));
#pragma unroll
          for (int i = 0; i < 4; ++i)
            b[i] = make_float4(am.x == s0 + i ? dd.x : 0.f, am.y == s0 + i ? dd.y : 0.f, am.z == s0 + i ? dd.z : 0.f,
                               am.w == s0 + i ? dd.w : 0.f);
        }
        const int set = p.sets == 2 ? (tl & 1) : 0;
        if (dgrad) {
          timed_wait(&bars->w_empty[grp], (w_use & 1) ^ 1, tm, c_w);  // W^T slot + staging slot free
          if (gt == 0) {  // this step's W^T tile: one contiguous pre-swizzled image
            const uint32_t bytes = static_cast<uint32_t>(p.kpad16) * BK * 4;
            mbar_expect_tx(&bars->w_full[grp], bytes);
            tma_load_1d(s_w + grp * p.w_slot_bytes, p.WTp + static_cast<size_t>(c) * p.kpad16 * BK, bytes,
                        &bars->w_full[grp]);
          }
        }
        ++w_use;
        if (tl >= p.sets)  // the set's previous tile has retired
          timed_wait(&bars->dz_empty[(tl - p.sets) & 3], ((tl - p.sets) >> 2) & 1, tm, c_dz);
        if (tm) {  // time until the item's global loads have landed
          const long long t0 = clock64();
          asm volatile("" ::"f"(b[0].x), "f"(b[4].x), "f"(b[3].w), "f"(b[7].w));
          c_ld += static_cast<unsigned long long>(clock64() - t0);
        }
        unsigned char *slab = s_dz + (set * nb + c) * SLAB;
        unsigned char *stage = s_a1 + grp * SLAB;
        const float4 fa = lds4(t_a + col), fb = lds4(t_b + col), fC = lds4(t_C + col), fD = lds4(t_D + col);
        const float2 a01 = make_float2(fa.x, fa.y), a23 = make_float2(fa.z, fa.w);
        const float2 b01 = make_float2(fb.x, fb.y), b23 = make_float2(fb.z, fb.w);
        const float2 C01 = make_float2(fC.x, fC.y), C23 = make_float2(fC.z, fC.w);
        const float2 D01 = make_float2(fD.x, fD.y), D23 = make_float2(fD.z, fD.w);
        const float thr = p.relu ? 0.f : -INFINITY;  // no ReLU: every gradient passes
        if (kDrop && p.drop_out != nullptr) {  // adjoint of the dropout on this layer's output
          const float sc = p.drop_out_scale;
#pragma unroll
          for (int i = 0; i < 4; ++i)
            b[i] = make_float4((mk[i] & 0xffu) ? b[i].x * sc : 0.f, (mk[i] & 0xff00u) ? b[i].y * sc : 0.f,
                               (mk[i] & 0xff0000u) ? b[i].z * sc : 0.f, (mk[i] & 0xff000000u) ? b[i].w * sc : 0.f);
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 v = b[4 + i], d = b[i];
          const float2 v01 = make_float2(v.x, v.y), v23 = make_float2(v.z, v.w);
          const float2 y01 = fma2(a01, v01, b01), y23 = fma2(a23, v23, b23);
          const float2 t01 = fma2(C01, v01, D01), t23 = fma2(C23, v23, D23);
          const float2 e01 = make_float2(y01.x > thr ? d.x : 0.f, y01.y > thr ? d.y : 0.f);
          const float2 e23 = make_float2(y23.x > thr ? d.z : 0.f, y23.y > thr ? d.w : 0.f);
          const float2 o01 = fma2(a01, e01, t01), o23 = fma2(a23, e23, t23);
          float4 o = make_float4(o01.x, o01.y, o23.x, o23.y);
          if (!full && r0 + i >= p.P) o = zero4();  // rows past the end (channels past N: zero constants)
          *reinterpret_cast<float4 *>(slab + off32[i]) = o;
          if (dgrad) *reinterpret_cast<float4 *>(stage + off16[i]) = o;
        }
        fence_proxy_async();
        mbar_arrive(&bars->dz_full[set][c]);
      } else {
        // ---------------- slab h of the layer below: columns [32h, 32h+32) ----------------
        const int col = (j - nb) * BK + chunk * 4;
        const bool fold = p.a_in != nullptr && col < p.K;
        const float4 fa = lds4(u_a + col), fb = lds4(u_b + col);
        const float2 a01 = make_float2(fa.x, fa.y), a23 = make_float2(fa.z, fa.w);
        const float2 b01 = make_float2(fb.x, fb.y), b23 = make_float2(fb.z, fb.w);
        timed_wait(&bars->h_empty[grp], (h_use & 1) ^ 1, tm, c_h);
        ++h_use;
        if (tm) {
          const long long t0 = clock64();
          asm volatile("" ::"f"(b[0].x), "f"(b[3].w));
          c_ld += static_cast<unsigned long long>(clock64() - t0);
        }
        unsigned char *slab = s_h + grp * hslabs * SLAB;
        if (!stats) {  // H_prev = relu(a_in * Z_prev + b_in), or the block's input rows as they are
          const float lo = p.relu_in ? 0.f : -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            float4 y = b[i];
            if (fold) {
              const float2 y01 = fma2(a01, make_float2(y.x, y.y), b01), y23 = fma2(a23, make_float2(y.z, y.w), b23);
              y = make_float4(fmaxf(y01.x, lo), fmaxf(y01.y, lo), fmaxf(y23.x, lo), fmaxf(y23.y, lo));
              if (kDrop && p.drop_in != nullptr) {
                const float sc = p.drop_in_scale;
                y = make_float4((mk[i] & 0xffu) ? y.x * sc : 0.f, (mk[i] & 0xff00u) ? y.y * sc : 0.f,
                                (mk[i] & 0xff0000u) ? y.z * sc : 0.f, (mk[i] & 0xff000000u) ? y.w * sc : 0.f);
              }
              if (!full && r0 + i >= p.P) y = zero4();
            }
            *reinterpret_cast<float4 *>(slab + off32[i]) = y;
          }
        } else {       // M = ReLU mask of the layer below, and M .* zhat_prev
          const float4 mu = lds4(u_mu + col), is = lds4(u_is + col);
          const float2 s01 = make_float2(is.x, is.y), s23 = make_float2(is.z, is.w);
          const float2 e01 = make_float2(-mu.x * is.x, -mu.y * is.y), e23 = make_float2(-mu.z * is.z, -mu.w * is.w);
          const float thr = p.relu_in ? 0.f : -INFINITY;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const float4 z = b[i];
            const float2 z01 = make_float2(z.x, z.y), z23 = make_float2(z.z, z.w);
            const float2 y01 = fma2(a01, z01, b01), y23 = fma2(a23, z23, b23);
            const float2 h01 = fma2(s01, z01, e01), h23 = fma2(s23, z23, e23);   // zhat
            float4 m = make_float4(y01.x > thr ? 1.f : 0.f, y01.y > thr ? 1.f : 0.f, y23.x > thr ? 1.f : 0.f,
                                   y23.y > thr ? 1.f : 0.f);
            if (kDrop && p.drop_in != nullptr) {  // M carries the dropout of the layer below
              const float keep = p.drop_in_scale;
              m = make_float4((mk[i] & 0xffu) ? m.x * keep : 0.f, (mk[i] & 0xff00u) ? m.y * keep : 0.f,
                              (mk[i] & 0xff0000u) ? m.z * keep : 0.f, (mk[i] & 0xff000000u) ? m.w * keep : 0.f);
            }
            if (!fold || (!full && r0 + i >= p.P)) m = zero4();
            const float2 q01 = mul2(make_float2(m.x, m.y), h01), q23 = mul2(make_float2(m.z, m.w), h23);
            *reinterpret_cast<float4 *>(slab + off32[i]) = m;
            *reinterpret_cast<float4 *>(slab + SLAB + off32[i]) = make_float4(q01.x, q01.y, q23.x, q23.y);
          }
        }
        fence_proxy_async();
        mbar_arrive(&bars->h_full[grp]);
      }
    };

    // this group's items: every kGroups-th slab of the CTA's tile sequence
    float4 buf[8];
    uint32_t mk[4];
    int tl = 0, j = grp;
    while (j >= ipt) j -= ipt, ++tl;
    int cur_tl = -1, tile = -1;
    while (true) {
      if (tl != cur_tl) cur_tl = tl, tile = tile_of(tl);
      if (tile < 0) break;
      issue(tile, j, buf, mk);
      process(tl, tile, j, buf, mk);
      j += kGroups;
      while (j >= ipt) j -= ipt, ++tl;
    }
    if (tm && gt == 0) {
      unsigned long long *o = g_wait_cycles[blockIdx.x] + grp * 5;
      o[0] = clock64() - c_tot, o[1] = c_w, o[2] = c_dz, o[3] = c_h, o[4] = c_ld;
    }
  } else if (warp == kMmaWarp) {
    // ===================================== MMA issuer =====================================
    {
      // One lane issues ~100 MMAs per tile, but the WHOLE warp runs this loop: with warp-uniform control
      // flow the descriptors live in uniform registers and each MMA costs a couple of uniform adds.
      // (Under `if (lane == 0)` every operand is re-materialised into uniform registers per
      // instruction -- ~30 SASS instructions per MMA, which made this thread the kernel's bottleneck.)
      const bool leader = lane == 0;
      const uint32_t idesc1 = make_idesc(BM, p.kpad16);                                          // dH_prev = dZ * W
      const uint32_t idesc2 = make_idesc(BM, hslabs * BK) | kIdescAMajorMN | kIdescBMajorMN;   // dW (S|Q) += dZ^T * H
      const uint64_t mn_hi = make_desc_mn32(0, SLAB, 512);   // MN-major template: LBO = slab stride
      const uint64_t k_hi = make_desc(0);                    // K-major template
      auto addr = [](uint32_t a) { return static_cast<uint64_t>((a & 0x3ffff) >> 4); };
      uint32_t w_par = 0, h_par = 0;  // bit g: phase parity of group g's W^T / H_prev slot
      const bool tm = (PN2_DBG(p) & 32) != 0;
      unsigned long long c_acc = 0, c_dzf = 0, c_wf = 0, c_hf = 0;
      int g = 0;  // loader group of the next item (items are dealt round-robin)
      // tile scheduler: ids of the tiles tl and tl + 1 are known; tl + 2 is fetched while tile tl's MMAs are issued
      auto fetch = [&](bool want) -> int {  // one lane asks, every lane gets the answer (uniform control flow)
        int id = -1;
        if (want && leader) {
          id = p.sched != nullptr ? atomicAdd(p.sched, 1) : -1;
          if (id >= m_tiles) id = -1;
        }
        return __shfl_sync(0xffffffffu, id, 0);
      };
      auto publish = [&](int tl, int id) {
        if (leader) {
          bars->tile_id[tl & 7] = id;
          mbar_arrive(&bars->tile_full[tl & 7]);
        }
      };
      int id_cur = fetch(true), id_next = fetch(id_cur >= 0);
      publish(0, id_cur), publish(1, id_next);
      int tl = 0;
      for (; id_cur >= 0; ++tl) {
        const int id_after = fetch(id_next >= 0);   // lands while this tile's MMAs are issued
        const int acc = tl & 1, acc_use = tl >> 1;
        const int set = p.sets == 2 ? (tl & 1) : 0, use = p.sets == 2 ? (tl >> 1) : tl;
        const uint64_t dz_desc = mn_hi | addr(smem_u32(s_dz + set * nb * SLAB));
        if (dgrad) {
          timed_wait(&bars->acc_empty[acc], (acc_use & 1) ^ 1, tm, c_acc);
          tc_fence_after();
        }
        for (int c = 0; c < nb; ++c) {
          timed_wait(&bars->dz_full[set][c], use & 1, tm, c_dzf);
          if (dgrad) {
            timed_wait(&bars->w_full[g], (w_par >> g) & 1, tm, c_wf);
            w_par ^= 1u << g;
            tc_fence_after();
            const uint64_t da = k_hi | addr(smem_u32(s_a1 + g * SLAB));
            const uint64_t db = k_hi | addr(smem_u32(s_w + g * p.w_slot_bytes));
            const uint32_t d = tm_acc + acc * kb32;
            if (leader) {
              tc_mma_tf32(d, da, db, idesc1, c != 0);
#pragma unroll
              for (int k = 1; k < BK / 8; ++k) tc_mma_tf32(d, da + 2 * k, db + 2 * k, idesc1, 1);
              tc_commit(&bars->w_empty[g]);
            }
            __syncwarp();
          }
          g ^= 1;
        }
        tc_fence_after();
        if (dgrad && leader) tc_commit(&bars->acc_full[acc]);
        for (int h = 0; h < kb; ++h) {
          timed_wait(&bars->h_full[g], (h_par >> g) & 1, tm, c_hf);
          h_par ^= 1u << g;
          tc_fence_after();
          // stats mode: the group's [M | M .* zhat] slabs are adjacent, one N = 64 instruction covers both
          const uint64_t hb = mn_hi | addr(smem_u32(s_h + g * hslabs * SLAB));
          const uint32_t d = tm_dw_addr + h * hslabs * BK;
          if (leader) {
            tc_mma_tf32(d, dz_desc, hb, idesc2, tl != 0);
#pragma unroll
            for (int ks = 1; ks < BM / 8; ++ks)  // 8 rows (two 4-row swizzle atoms, 1024 bytes) per instruction
              tc_mma_tf32(d, dz_desc + 64 * ks, hb + 64 * ks, idesc2, 1);
            tc_commit(&bars->h_empty[g]);
          }
          __syncwarp();
          g ^= 1;
        }
        if (leader) tc_commit(&bars->dz_empty[tl & 3]);
        publish(tl + 2, id_after);
        id_cur = id_next, id_next = id_after;
      }
      if (leader) {
        bars->tiles_done = tl;
        tc_commit(&bars->dw_full);
        // every CTA draws exactly one id past the end; the last one to do so re-arms the counters
        if (p.sched != nullptr && atomicAdd(p.sched + 1, 1) == static_cast<int>(gridDim.x) - 1) {
          p.sched[0] = 0, p.sched[1] = 0;
          __threadfence();
        }
      }
      if (tm && leader) {
        unsigned long long *o = g_wait_cycles[blockIdx.x] + 10;
        o[0] = c_acc, o[1] = c_dzf, o[2] = c_wf, o[3] = c_hf;
      }
    }
  } else {
    // ====================================== epilogue ======================================
    const int wq = warp;  // TMEM lane quarter
    if (dgrad) {
      for (int tl = 0;; ++tl) {
        const int tile = tile_of(tl), acc = tl & 1, use = tl >> 1;
        if (tile < 0) break;
        const int grow = tile * BM + wq * 32 + lane;
        const bool rowok = grow < p.P;
        {
          unsigned long long c_e = 0;
          timed_wait(&bars->acc_full[acc], use & 1, (PN2_DBG(p) & 32) != 0, c_e);
          if ((PN2_DBG(p) & 32) && tid == 0) g_wait_cycles[blockIdx.x][14] = (tl == 0 ? 0 : g_wait_cycles[blockIdx.x][14]) + c_e;
        }
        tc_fence_after();
#pragma unroll
        for (int h = 0; h < kMaxKB; ++h) {
          if (h < kb) {
            const int n0 = h * BK;
            float v[32];
            tc_ld32(tm_acc + (static_cast<uint32_t>(wq * 32) << 16) + acc * kb32 + n0, v);
            if (p.out_bytes && !(PN2_DBG(p) & 4)) {  // staged block -> one TMA tensor store (see tc_common.cuh)
              tma_store_block32(&tm_dh, s_out + wq * 4096, v, lane, n0, tile * BM + wq * 32);
            } else if (rowok && !(PN2_DBG(p) & 4)) {
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
          }
        }
        tc_fence_before();
        mbar_arrive(&bars->acc_empty[acc]);
      }
      tma_store_drain(lane);
    }
    // read-out.  TMEM lanes = output channels n of this layer, columns = channels c of the layer below.
    mbar_wait(&bars->dw_full, 0);
    tc_fence_after();
    const long long t_ro = (PN2_DBG(p) & 32) ? clock64() : 0;
    if ((PN2_DBG(p) & 32) && tid == 0) g_wait_cycles[blockIdx.x][17] = t_ro - t_entry;  // entry -> all MMAs retired
    // With a BatchNorm below, the two accumulators S = dZ^T M and Q = dZ^T (M .* zhat) are written out
    // as they are: dW = g1 .* Q + g0 .* S and the statistics sum_n W .* S, sum_n W .* Q are linear in
    // them, so pn2_mlp_bwd_finalize forms them ONCE from the sums over all CTAs (done per CTA here,
    // with W re-read from L2 by every CTA at the same moment, this tail took 47 us at N = K = 128).
    const int n = wq * 32 + lane;
    const bool mine = wq * 32 < nb32;
    const bool active = mine && bars->tiles_done > 0;  // written before the dw_full commit
    const int row0 = blockIdx.x * nb32 + wq * 32;
    auto put = [&](float *base, const CUtensorMap *tm, const float (&v)[32], int h) {
      if (!mine) return;
      if (p.out_bytes) {
        tma_store_block32(tm, s_out + wq * 4096, v, lane, h * 32, row0);
      } else {
        float *dst = base + (static_cast<size_t>(row0) + lane) * kb32 + h * 32;
#pragma unroll
        for (int jj = 0; jj < 32; jj += 8) st_global_v8(dst + jj, v + jj);
      }
    };
#pragma unroll
    for (int h = 0; h < kMaxKB; ++h) {
      if (h < kb) {
        float v[32];
#pragma unroll
        for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
        if (active) tc_ld32(tm_dw_addr + (static_cast<uint32_t>(wq * 32) << 16) + h * hslabs * BK, v);
        put(p.dWpart, &tm_dw, v, h);
        if (stats) {
#pragma unroll
          for (int jj = 0; jj < 32; ++jj) v[jj] = 0.f;
          if (active) tc_ld32(tm_dw_addr + (static_cast<uint32_t>(wq * 32) << 16) + h * hslabs * BK + BK, v);
          put(p.stats_part, &tm_q, v, h);
        }
      }
    }
    tma_store_drain(lane);
    (void)n;
    if ((PN2_DBG(p) & 32) && tid == 0) g_wait_cycles[blockIdx.x][15] = clock64() - t_ro;  // read-out
  }

  tc_fence_before();
  __syncthreads();
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
  pdl_enter();
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

// Everything that follows a fused backward layer with a BatchNorm below, in ONE launch: the per-CTA
// images S_p = dZ^T M and Q_p = dZ^T (M .* zhat_prev) are summed over the CTAs (fixed order), then
//   dW[n, c]   = g1[c] * Q[n, c] + g0[c] * S[n, c]      g1 = a_in / invstd_in, g0 = a_in * mean_in + b_in
//   dbeta[c]   = sum_n W[n, c] * S[n, c]                 (gradient wrt the layer below's BN output, summed)
//   dgamma[c]  = sum_n W[n, c] * Q[n, c]
//   c1 = dbeta / P, c2 = dgamma / P                      (0 in eval mode)
// Grid (column slab of 32, group of 8 rows n), 512 threads = 8 image slices x 8 rows x 8 float4 column
// groups: a thread adds the images i = slice, slice + 8, ... of its (n, c4) -- independent 16-byte loads,
// a warp covers four 128-byte row segments -- the slices are combined through shared memory in fixed
// order, then the CTA's 8 rows; the CTA that finishes last adds the row groups' partial column sums (fixed
// order, fp64) into dbeta / dgamma / c1 / c2.
constexpr int kFinRows = 8, kFinSlices = 8, kFinThreads = 64 * kFinSlices;
__global__ void __launch_bounds__(kFinThreads)
bwd_layer_finalize_sq_kernel(int parts, int N, int K, int nb32, int kb32, const float *__restrict__ s_part,
                             const float *__restrict__ q_part, const float *__restrict__ W,
                             const float *__restrict__ a_in, const float *__restrict__ b_in,
                             const float *__restrict__ mean_in, const float *__restrict__ invstd_in, int P,
                             int training, float *__restrict__ dW, float *__restrict__ dgamma,
                             float *__restrict__ dbeta, float *__restrict__ c1, float *__restrict__ c2,
                             float *__restrict__ scratch, unsigned *__restrict__ counter) {
  pdl_enter();
  __shared__ float4 s_red[kFinSlices][64][2];
  __shared__ float s_col[kFinRows][2][32];
  __shared__ bool s_last;
  const int tid = threadIdx.x, c4l = tid & 7, r = (tid >> 3) & 7, slice = tid >> 6;
  const int n = blockIdx.y * kFinRows + r, c0 = blockIdx.x * 32 + c4l * 4;
  const size_t image = static_cast<size_t>(nb32) * kb32, off = static_cast<size_t>(n) * kb32 + c0;
  float4 s = zero4(), q = zero4();
#pragma unroll 4
  for (int i = slice; i < parts; i += kFinSlices) {
    const float4 a = ldg4(s_part + i * image + off), b = ldg4(q_part + i * image + off);
    s.x += a.x, s.y += a.y, s.z += a.z, s.w += a.w;
    q.x += b.x, q.y += b.y, q.z += b.z, q.w += b.w;
  }
  s_red[slice][tid & 63][0] = s, s_red[slice][tid & 63][1] = q;
  __syncthreads();
  if (slice == 0) {
#pragma unroll
    for (int k = 1; k < kFinSlices; ++k) {
      const float4 a = s_red[k][tid][0], b = s_red[k][tid][1];
      s.x += a.x, s.y += a.y, s.z += a.z, s.w += a.w;
      q.x += b.x, q.y += b.y, q.z += b.z, q.w += b.w;
    }
    const float sv[4] = {s.x, s.y, s.z, s.w}, qv[4] = {q.x, q.y, q.z, q.w};
    float o[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int c = c0 + e;
      const bool ok = c < K;
      const float is = ok ? __ldg(invstd_in + c) : 0.f, a = ok ? __ldg(a_in + c) : 0.f;
      const float g1 = is > 0.f ? a / is : 0.f;
      const float g0 = ok ? fmaf(a, __ldg(mean_in + c), __ldg(b_in + c)) : 0.f;
      o[e] = fmaf(g1, qv[e], g0 * sv[e]);
      const float w = (ok && n < N) ? __ldg(W + static_cast<size_t>(n) * K + c) : 0.f;
      s_col[r][0][c4l * 4 + e] = w * sv[e];
      s_col[r][1][c4l * 4 + e] = w * qv[e];
    }
    *reinterpret_cast<float4 *>(dW + off) = make_float4(o[0], o[1], o[2], o[3]);
  }
  __syncthreads();
  if (tid < 64) {  // the CTA's 8 rows, fixed order -> scratch [gridDim.y][2][kb32]
    const int which = tid >> 5, cl = tid & 31;
    float t = 0.f;
#pragma unroll
    for (int k = 0; k < kFinRows; ++k) t += s_col[k][which][cl];
    scratch[(static_cast<size_t>(blockIdx.y) * 2 + which) * kb32 + blockIdx.x * 32 + cl] = t;
    __threadfence();
  }
  __syncthreads();
  if (tid == 0) s_last = atomicAdd(counter, 1u) == gridDim.x * gridDim.y - 1;
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  for (int c = tid; c < K; c += kFinThreads) {
    double sb = 0.0, sg = 0.0;
    for (unsigned g = 0; g < gridDim.y; ++g) {
      sb += static_cast<double>(__ldcg(scratch + (static_cast<size_t>(g) * 2 + 0) * kb32 + c));
      sg += static_cast<double>(__ldcg(scratch + (static_cast<size_t>(g) * 2 + 1) * kb32 + c));
    }
    dbeta[c] = static_cast<float>(sb);
    dgamma[c] = static_cast<float>(sg);
    const float inv_p = 1.0f / static_cast<float>(P);
    c1[c] = training ? static_cast<float>(sb) * inv_p : 0.f;
    c2[c] = training ? static_cast<float>(sg) * inv_p : 0.f;
  }
  if (tid == 0) *counter = 0u;  // re-armed for the next launch on this stream
}

struct Geom {
  int nb, kb, kpad16, sets, w_slot_bytes, tmem_cols, grid, out_bytes;
  size_t smem;
};

inline bool geom(int p, int n, int k, int lda, bool dgrad, bool stats, Geom *g) {
  if (n <= 0 || k <= 0 || n > kMaxNB * BK || lda > kMaxKB * BK || n % 4 != 0 || lda % 4 != 0 || lda < k) return false;
  g->nb = (n + BK - 1) / BK;
  g->kb = (lda + BK - 1) / BK;
  g->kpad16 = (lda + 15) / 16 * 16;
  g->w_slot_bytes = (g->kpad16 * BK * 4 + 1023) / 1024 * 1024;
  const int cols = ((dgrad ? 2 : 0) + (stats ? 2 : 1)) * g->kb * BK;
  g->tmem_cols = 32;
  while (g->tmem_cols < cols) g->tmem_cols *= 2;
  if (g->tmem_cols > 512) return false;
  g->sets = g->nb <= 2 ? 2 : 1;
  const size_t fixed = static_cast<size_t>(g->sets) * g->nb * SLAB +
                       sizeof(float) * (6 * g->nb * 32 + 4 * g->kb * 32 + kEpiWarps * 2 * g->kb * 32) +
                       sizeof(Barriers) + 1024;
  const size_t per_group = (stats ? 3 : 2) * static_cast<size_t>(SLAB) + g->w_slot_bytes;
  g->smem = fixed + kGroups * per_group;
  g->out_bytes = 0;
  if (dgrad && g->smem + kEpiWarps * 4096 <= 227 * 1024) g->out_bytes = kEpiWarps * 4096, g->smem += g->out_bytes;
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

// 1 when pn2_mlp_bwd_layer supports the shape: N <= 128 output channels; lda <= 160 input columns,
// <= 128 when the statistics of a BatchNorm below are wanted (with_stats)
extern "C" int pn2_mlp_bwd_supported(int p, int n, int k, int lda, int with_stats) {
  bwd::Geom g;
  return p > 0 && bwd::geom(p, n, k, lda, true, with_stats != 0, &g) ? 1 : 0;
}
extern "C" __attribute__((visibility("default"))) int pn2_mlp_bwd_debug_read(unsigned long long *out) {
  return static_cast<int>(cudaMemcpyFromSymbol(out, bwd::g_wait_cycles, sizeof(unsigned long long) * kNumSMs * 20));
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
  pn2::launch(bwd::pack_weights_t_kernel, blocks, 256, 0, static_cast<cudaStream_t>(stream), w, n, k, ldw, nb, kpad16, packed);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_mlp_bwd_layer(int p, int n, int k, int lda, int ns, int relu, int relu_in, const float *dh,
                                 const float *dout, const int *argmax, const float *z, const float *a,
                                 const float *b, const float *mean, const float *invstd, const float *c1c2,
                                 const float *w, const float *a_prev, const float *a_in, const float *b_in,
                                 const float *mean_in, const float *invstd_in, const float *wt_packed,
                                 float *dh_prev, float *dw_part, float *stats_part, int *sched,
                                 const unsigned char *drop_out, float drop_out_scale, const unsigned char *drop_in,
                                 float drop_in_scale, void *stream) {
  if (p <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(z, a, b, mean, invstd, c1c2, a_prev, dw_part, sched);
  if (dh == nullptr && (dout == nullptr || argmax == nullptr || ns <= 0 || p % ns != 0)) return PN2_ERR_INVALID_ARG;
  if (dh == nullptr && ns % 4 != 0) return PN2_ERR_UNSUPPORTED;  // a thread's 4 consecutive rows share one group
  if ((wt_packed == nullptr) != (dh_prev == nullptr)) return PN2_ERR_INVALID_ARG;
  if (stats_part != nullptr && (wt_packed == nullptr || a_in == nullptr || mean_in == nullptr ||
                                invstd_in == nullptr || w == nullptr))
    return PN2_ERR_INVALID_ARG;
  if (a_in != nullptr && b_in == nullptr) return PN2_ERR_INVALID_ARG;
  bwd::Geom g;
  if (!bwd::geom(p, n, k, lda, wt_packed != nullptr, stats_part != nullptr, &g)) return PN2_ERR_UNSUPPORTED;
  if (!aligned16(z) || !aligned16(a_prev) || !aligned16(dw_part) || (stats_part != nullptr && !aligned16(stats_part)) || (dh != nullptr && !aligned16(dh)) ||
      (dout != nullptr && (!aligned16(dout) || !aligned16(argmax))) || (dh_prev != nullptr && !aligned16(dh_prev)) ||
      (wt_packed != nullptr && !aligned16(wt_packed)))
    return PN2_ERR_UNSUPPORTED;
  bwd::Params prm;
  prm.dH = dh, prm.dout = dout, prm.argmax = argmax, prm.Z = z;
  prm.a = a, prm.b = b, prm.mean = mean, prm.invstd = invstd, prm.c1c2 = c1c2, prm.W = w;
  prm.Aprev = a_prev, prm.a_in = a_in, prm.b_in = b_in, prm.mean_in = mean_in, prm.invstd_in = invstd_in;
  prm.WTp = wt_packed, prm.dHprev = dh_prev, prm.dWpart = dw_part, prm.stats_part = stats_part;
  prm.P = p, prm.N = n, prm.nb = g.nb, prm.K = k, prm.lda = lda, prm.kb = g.kb, prm.kpad16 = g.kpad16;
  prm.ns = ns > 0 ? ns : 1, prm.relu = relu, prm.relu_in = relu_in;
  prm.sched = sched;
  if ((drop_out != nullptr && dh == nullptr) || (drop_in != nullptr && a_in == nullptr)) return PN2_ERR_INVALID_ARG;
  prm.drop_out = drop_out, prm.drop_in = drop_in, prm.drop_out_scale = drop_out_scale, prm.drop_in_scale = drop_in_scale;
  prm.dbg = getenv("PN2_BWD_DBG") != nullptr ? atoi(getenv("PN2_BWD_DBG")) : 0;
  prm.ns_shift = -1;
  for (int sh = 0; sh < 31; ++sh)
    if ((1 << sh) == prm.ns) prm.ns_shift = sh;
  prm.sets = g.sets, prm.w_slot_bytes = g.w_slot_bytes, prm.tmem_cols = g.tmem_cols;
  CUtensorMap tm_dh;
  memset(&tm_dh, 0, sizeof(tm_dh));
  static const bool no_tma = getenv("PN2_BWD_NO_TMA_STORE") != nullptr;   // timing experiments only
  CUtensorMap tm_dw, tm_q;
  memset(&tm_dw, 0, sizeof(tm_dw)), memset(&tm_q, 0, sizeof(tm_q));
  const int part_rows = g.grid * g.nb * 32, part_cols = g.kb * 32;  // per-CTA images of dW (or S and Q)
  prm.out_bytes = (g.out_bytes != 0 && !no_tma && tc::make_tmap_rows32(&tm_dh, dh_prev, p, lda, lda) &&
                   tc::make_tmap_rows32(&tm_dw, dw_part, part_rows, part_cols, part_cols) &&
                   (stats_part == nullptr || tc::make_tmap_rows32(&tm_q, stats_part, part_rows, part_cols, part_cols)))
                      ? g.out_bytes : 0;
  if (prm.out_bytes == 0) g.smem -= g.out_bytes;
  const bool pooled = prm.dH == nullptr;
  if (pooled && (drop_out != nullptr || drop_in != nullptr)) return PN2_ERR_UNSUPPORTED;
  const bool full = prm.WTp != nullptr && prm.stats_part != nullptr;
  const bool drop = drop_out != nullptr || drop_in != nullptr;
  auto kern = bwd::mlp_bwd_layer_kernel<false, false, false, 0, 0>;
#define PN2_BWD_PICK(NB_, KB_)                                                                         \
  (drop ? (full ? bwd::mlp_bwd_layer_kernel<true, false, true, NB_, KB_> : bwd::mlp_bwd_layer_kernel<true, false, false, 0, 0>) \
        : pooled ? (full ? bwd::mlp_bwd_layer_kernel<false, true, true, NB_, KB_> : bwd::mlp_bwd_layer_kernel<false, true, false, 0, 0>) \
                 : (full ? bwd::mlp_bwd_layer_kernel<false, false, true, NB_, KB_> : bwd::mlp_bwd_layer_kernel<false, false, false, 0, 0>))
  // shapes of the reference's fine SA blocks: SA2's pooled layer (64 -> 128), SA1 (32 -> 32 -> 64)
  // (measured per shape in one call, us, specialised vs not: 64 x 32 pooled 140 vs 146, 32 x 32 106 vs 114, 128 x 64 pooled 87 vs 90;
  // 128 x 128 got SLOWER unrolled -- 155 vs 145, with dropout 182 vs 166 -- and 64 x 64 did not move: they stay generic)
  if (prm.nb == 4 && prm.kb == 2) kern = PN2_BWD_PICK(4, 2);
  else if (prm.nb == 2 && prm.kb == 1) kern = PN2_BWD_PICK(2, 1);
  else if (prm.nb == 1 && prm.kb == 1) kern = PN2_BWD_PICK(1, 1);
  else kern = PN2_BWD_PICK(0, 0);
#undef PN2_BWD_PICK
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(g.smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  pn2::launch(kern, g.grid, bwd::kThreads, g.smem, static_cast<cudaStream_t>(stream), prm, tm_dh, tm_dw, tm_q);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_mlp_bwd_finalize_scratch_floats(int n, int lda) {
  return pn2_mlp_bwd_nb(n) * (32 / bwd::kFinRows) * 2 * pn2_mlp_bwd_kb(lda) * 32;
}

extern "C" int pn2_mlp_bwd_finalize(int parts, int n, int k, int lda, const float *s_part, const float *q_part,
                                    const float *w, const float *a_in, const float *b_in, const float *mean_in,
                                    const float *invstd_in, int p, int training, float *dw, float *dgamma,
                                    float *dbeta, float *c1c2, float *scratch, unsigned *counter, void *stream) {
  if (parts <= 0 || n <= 0 || k <= 0 || lda < k || p <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(s_part, q_part, w, a_in, b_in, mean_in, invstd_in, dw, dgamma, dbeta, c1c2, scratch, counter);
  if (!aligned16(s_part) || !aligned16(q_part) || !aligned16(dw)) return PN2_ERR_UNSUPPORTED;
  const int nb32 = pn2_mlp_bwd_nb(n) * 32, kb32 = pn2_mlp_bwd_kb(lda) * 32;
  dim3 grid(kb32 / 32, nb32 / bwd::kFinRows);
  pn2::launch(bwd::bwd_layer_finalize_sq_kernel, grid, bwd::kFinThreads, 0, static_cast<cudaStream_t>(stream), parts, n, k, nb32, kb32, s_part, q_part, w, a_in, b_in, mean_in, invstd_in, p, training, dw, dgamma, dbeta,
      c1c2, c1c2 + k, scratch, counter);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
