// Whole set-abstraction block in ONE kernel for inference (tcgen05 + TMEM), sm_100a:
//
//     out[g, :] = max over the nsample rows of group g of  relu(BN3(W3 relu(BN2(W2 relu(BN1(W1 x))))))
//     x = [xyz(idx) - centre | features(idx)]                      (BatchNorm folded: y = a * z + b)
//
// i.e. QueryAndGroup -> SharedMLP (three Conv2d 1x1 + BatchNorm2d + ReLU) -> max_pool2d over nsample of
// PointnetSAModule(MSG).forward in eval mode (/root/reference/pointnet2/pointnet2_modules.py:60-75,
// pointnet2_utils.py:317-359, pytorch_utils.py:11-36) -- the "1x1 conv + folded BN + ReLU + max over
// nsample" fused grouped MLP.  The layer-by-layer path writes and re-reads every [P, C] activation
// (P = B * npoint * nsample rows: 1.1 GB for SSG SA1 at B = 32); here nothing but the gathered input rows
// (from the L2-resident point table) and the pooled output touches memory: the activations go
// TMEM -> registers (fold + ReLU) -> shared memory (next layer's A operand) -> tensor core.
//
// Per CTA (persistent over 128-row tiles = 128 / nsample whole groups), 13 warps:
//   warps 8-11 gather: thread r builds row r of the tile's A0 operand -- 16-byte loads of table[b, idx[p], :],
//              centre subtracted from the first three columns -- in the 128B-swizzled K-major UMMA layout
//   warp  12   one elected thread issues tcgen05.mma.kind::tf32 for the three layers (weights resident in
//              shared memory as pre-packed tile images, loaded once by 1-D TMA bulk copies)
//   warps 0-7  epilogues (two warpgroups: one per tile in flight, or alternating 32-column chunks of the
//              single tile): layers 1, 2: tcgen05.ld -> relu(a z + b) -> next A operand; layer 3:
//              relu(a z + b) and a shuffle butterfly max over the 32 (or 2 x 16) lanes of the warp, which are
//              exactly the rows of one (two) group(s); one 128-byte store per group and 32 channels
// Two tiles are in flight (two sets of operand tiles and accumulators) where shared / tensor memory
// allow: the tensor core and the gather of the next pair overlap the epilogues, which bound the kernel.
//
// Arithmetic is that of the layer-by-layer tensor-core path (same fp32 fold expression, activations read
// by the tensor core as TF32 by truncation, weights rounded to nearest, fp32 accumulation over K in the same
// order), so the two produce IDENTICAL bits (tests/test_chain_gpu.py).
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pn2 {
namespace chain {

using namespace tc;

constexpr int BM = 128;
constexpr int SLAB = BM * 128;               // [128 rows][32 fp32] = 16 KB
constexpr int kEpiWarps = 8, kGatherWarps = 4;   // two epilogue warpgroups (TMEM lane quarter = warp % 4)
constexpr int kThreads = 32 * (kEpiWarps + kGatherWarps + 1);
constexpr int kSmemMax = 220 * 1024;

struct Params {
  const float *table;    // [B, N, ct]
  const float *new_xyz;  // [B, npoint, 3]
  const int *idx;        // [B, npoint, ns]
  const float *w[3];     // packed weight tiles [k_blocks][C_l][32] (pn2_mlp_pack_weights_bn)
  const float *fold;     // a1 | b1 | a2 | b2 | a3 | b3
  float *out;            // [B * npoint, C3]
  int P, G, n, npoint, ns, ct;
  int c[4];              // c[0] = ct rounded up to 8 (K of layer 1), c[1..3] = layer widths
  int kb[3];             // K blocks (32 columns) of layers 1..3
  int slots;             // tiles in flight: 1 or 2
  int tmem_cols, d_off[3], slot_cols;   // accumulator columns of layer l at slot * slot_cols + d_off[l]
  int a_off[3];          // byte offsets of the A0 / A1 / A2 operand tiles inside a slot's block
  int slot_bytes, w_off[3], w_bytes[3], fold_off, bars_off;
};

struct __align__(8) Barriers {
  uint64_t w_full;           // weights landed                          (1 arrival + tx bytes)
  uint64_t a0_full[2];       // gathered rows of the slot's tile ready  (128 arrivals)
  uint64_t a0_empty[2];      // layer-1 MMAs retired                    (1 commit)
  uint64_t d_full[2][3];     // accumulator of layer l complete         (1 commit)
  uint64_t a_full[2][2];     // A operand of layer l + 1 written        (128 arrivals)
  uint64_t tile_done[2];     // the slot's accumulators drained         (128 arrivals)
  uint32_t tmem_base;
};

// max over the lanes of a group for every column: kHalf = false: 32 lanes, lane l ends with column l in v[0];
// kHalf = true: two groups of 16 lanes, lane l ends with columns 2 (l & 15), 2 (l & 15) + 1 in v[0], v[1]
template <bool kHalf>
__device__ __forceinline__ void group_colmax(float (&v)[32], int lane) {
  if (!kHalf) {
    const bool up = (lane & 16) != 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      const float send = up ? v[i] : v[i + 16], keep = up ? v[i + 16] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, 16));
    }
  }
#pragma unroll
  for (int o = 8; o >= 1; o >>= 1) {
    const int d = kHalf ? 2 * o : o;   // distance of the column pair inside v
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < d; ++i) {
      const float send = up ? v[i] : v[i + d], keep = up ? v[i + d] : v[i];
      v[i] = fmaxf(keep, __shfl_xor_sync(0xffffffffu, send, o));
    }
  }
}

__global__ void __launch_bounds__(kThreads, 1) sa_chain_eval_kernel(Params p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  Barriers *bars = reinterpret_cast<Barriers *>(smem + p.bars_off);
  const float *s_fold = reinterpret_cast<const float *>(smem + p.fold_off);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int m_tiles = (p.P + BM - 1) / BM;
  const int my_tiles = blockIdx.x < m_tiles ? (m_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto slot_base = [&](int s) { return smem + static_cast<size_t>(s) * p.slot_bytes; };

  if (tid == 0) {
    mbar_init(&bars->w_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&bars->a0_full[s], 32 * kGatherWarps);
      mbar_init(&bars->a0_empty[s], 1);
      for (int l = 0; l < 3; ++l) mbar_init(&bars->d_full[s][l], 1);
      // two tiles in flight: one epilogue warpgroup per slot (128 arrivals); one tile: both warpgroups share
      // the slot, alternating 32-column chunks (256 arrivals)
      const int epi = p.slots == 2 ? 128 : 256;
      for (int l = 0; l < 2; ++l) mbar_init(&bars->a_full[s][l], epi);
      mbar_init(&bars->tile_done[s], epi);
    }
    mbar_fence_init();
    if (my_tiles > 0) {
      mbar_expect_tx(&bars->w_full, static_cast<uint32_t>(p.w_bytes[0] + p.w_bytes[1] + p.w_bytes[2]));
      for (int l = 0; l < 3; ++l)
        tma_load_1d(smem + p.w_off[l], p.w[l], static_cast<uint32_t>(p.w_bytes[l]), &bars->w_full);
    }
  }
  if (warp == kEpiWarps + kGatherWarps) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&bars->tmem_base)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  {  // fold constants: a1 | b1 | a2 | b2 | a3 | b3
    float *dst = reinterpret_cast<float *>(smem + p.fold_off);
    const int total = 2 * (p.c[1] + p.c[2] + p.c[3]);
    for (int i = tid; i < total; i += kThreads) dst[i] = __ldg(p.fold + i);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp >= kEpiWarps && warp < kEpiWarps + kGatherWarps) {
    // ========================================= gather =========================================
    // Two dependent global-memory latencies per row (index -> point row) bound this role: the indices and
    // centres of the NEXT pair of tiles are fetched while the current pair is gathered, and when a row fits
    // one K block (ct <= 32) the rows of both tiles are in flight together.
    const int r = tid - 32 * kEpiWarps;                 // row of the tile
    const int np_ns = p.npoint * p.ns, ct4 = p.ct / 4;
    struct Row { const float4 *src; float cx, cy, cz; bool ok; };
    auto fetch = [&](int t) -> Row {                    // t = index into this CTA's tile sequence
      Row q;
      q.src = nullptr, q.cx = q.cy = q.cz = 0.f, q.ok = false;
      if (t >= my_tiles) return q;
      const int tile = blockIdx.x + t * gridDim.x;
      const long long row = static_cast<long long>(tile) * BM + r;
      if (row >= p.P) return q;
      const long long cloud = row / np_ns;
      const int k = __ldg(p.idx + row);
      q.src = reinterpret_cast<const float4 *>(p.table + (cloud * p.n + k) * p.ct);
      const float *ctr = p.new_xyz + (row / p.ns) * 3;
      q.cx = __ldg(ctr), q.cy = __ldg(ctr + 1), q.cz = __ldg(ctr + 2);
      q.ok = true;
      return q;
    };
    Row nxt[2] = {fetch(0), fetch(1)};
    for (int it = 0, t0 = 0; t0 < my_tiles; t0 += p.slots, ++it) {
      const int nt = min(p.slots, my_tiles - t0);
      Row cur[2] = {nxt[0], nxt[1]};
      if (p.slots == 2) nxt[0] = fetch(t0 + 2), nxt[1] = fetch(t0 + 3);
      else nxt[0] = fetch(t0 + 1);
      if (p.kb[0] == 1) {
        float4 v[2][8];
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
          for (int j = 0; j < 8; ++j)
            v[s][j] = (s < nt && cur[s].ok && j < ct4) ? __ldg(cur[s].src + j) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int s = 0; s < 2; ++s) {
          if (s < nt) {
            mbar_wait(&bars->a0_empty[s], (it & 1) ^ 1);    // layer-1 MMAs of the slot's previous tile retired
            if (cur[s].ok) v[s][0].x -= cur[s].cx, v[s][0].y -= cur[s].cy, v[s][0].z -= cur[s].cz;
            unsigned char *slab = slot_base(s) + p.a_off[0] + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if (j * 4 < p.c[0]) *reinterpret_cast<float4 *>(slab + ((j ^ (r & 7)) << 4)) = v[s][j];
            fence_proxy_async();
            mbar_arrive(&bars->a0_full[s]);
          }
        }
      } else {
        for (int s = 0; s < nt; ++s) {
          mbar_wait(&bars->a0_empty[s], (it & 1) ^ 1);
          unsigned char *a0 = slot_base(s) + p.a_off[0];
          const Row q = cur[s];
          for (int kb = 0; kb < p.kb[0]; ++kb) {
            float4 v[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              const int c4 = kb * 8 + j;
              v[j] = (q.ok && c4 < ct4) ? __ldg(q.src + c4) : make_float4(0.f, 0.f, 0.f, 0.f);
            }
            if (kb == 0 && q.ok) v[0].x -= q.cx, v[0].y -= q.cy, v[0].z -= q.cz;
            unsigned char *slab = a0 + kb * SLAB + r * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j)
              if ((kb * 8 + j) * 4 < p.c[0]) *reinterpret_cast<float4 *>(slab + ((j ^ (r & 7)) << 4)) = v[j];
          }
          fence_proxy_async();
          mbar_arrive(&bars->a0_full[s]);
        }
      }
    }
  } else if (warp == kEpiWarps + kGatherWarps) {
    // ======================================= MMA issuer =======================================
    const bool leader = lane == 0;
    const uint64_t k_hi = make_desc(0);
    auto addr = [](uint32_t a) { return static_cast<uint64_t>((a & 0x3ffff) >> 4); };
    if (my_tiles > 0) mbar_wait(&bars->w_full, 0);
    for (int it = 0, t0 = 0; t0 < my_tiles; t0 += p.slots, ++it) {
      const int nt = min(p.slots, my_tiles - t0);
      for (int l = 0; l < 3; ++l) {
        const uint32_t idesc = make_idesc(BM, p.c[l + 1]);
        const int K = p.c[l];                            // columns of the A operand actually multiplied
        for (int s = 0; s < nt; ++s) {
          if (l == 0) {
            mbar_wait(&bars->tile_done[s], (it & 1) ^ 1);   // the slot's accumulators were drained
            mbar_wait(&bars->a0_full[s], it & 1);
          } else {
            mbar_wait(&bars->a_full[s][l - 1], it & 1);
          }
          tc_fence_after();
          const uint32_t d = tmem_base + s * p.slot_cols + p.d_off[l];
          const uint32_t a_base = smem_u32(slot_base(s) + p.a_off[l]);
          const uint32_t w_base = smem_u32(smem + p.w_off[l]);
          const uint32_t w_blk = static_cast<uint32_t>(p.c[l + 1]) * 128u;
          if (leader) {
            for (int kb = 0; kb < p.kb[l]; ++kb) {
              const uint64_t da = k_hi | addr(a_base + kb * SLAB), db = k_hi | addr(w_base + kb * w_blk);
              const int steps = min(4, (K - kb * 32 + 7) / 8);
              for (int j = 0; j < steps; ++j) tc_mma_tf32(d, da + 2 * j, db + 2 * j, idesc, (kb | j) != 0);
            }
            if (l == 0) tc_commit(&bars->a0_empty[s]);
            tc_commit(&bars->d_full[s][l]);
          }
          __syncwarp();
        }
      }
    }
  } else {
    // ======================================== epilogues ========================================
    const int wg = warp >> 2, wq = warp & 3;              // warpgroup; TMEM lane quarter = rows 32 wq .. 32 wq + 31
    const int r = wq * 32 + lane;
    const uint32_t lane_off = static_cast<uint32_t>(wq * 32) << 16;
    const bool share = p.slots == 1;                      // both warpgroups work on slot 0, alternating chunks
    for (int it = 0, t0 = 0; t0 < my_tiles; t0 += p.slots, ++it) {
      const int nt = min(p.slots, my_tiles - t0);
      const int s = share ? 0 : wg;
      if (s >= nt) continue;
      for (int l = 0; l < 3; ++l) {
        const int C = p.c[l + 1];
        const float *fa = s_fold + 2 * (l > 0 ? p.c[1] : 0) + 2 * (l > 1 ? p.c[2] : 0), *fb = fa + C;
        mbar_wait(&bars->d_full[s][l], it & 1);
        tc_fence_after();
        const uint32_t d = tmem_base + lane_off + s * p.slot_cols + p.d_off[l];
        if (l < 2) {
          unsigned char *a_next = slot_base(s) + p.a_off[l + 1] + r * 128;
          for (int c0 = share ? 32 * wg : 0; c0 < C; c0 += share ? 64 : 32) {
            float v[32];
            tc_ld32(d + c0, v);
            unsigned char *slab = a_next + (c0 >> 5) * SLAB;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (c0 + 4 * j < C) {
                const float4 a4 = *reinterpret_cast<const float4 *>(fa + c0 + 4 * j);
                const float4 b4 = *reinterpret_cast<const float4 *>(fb + c0 + 4 * j);
                float4 y;
                y.x = fmaxf(fmaf(a4.x, v[4 * j], b4.x), 0.f), y.y = fmaxf(fmaf(a4.y, v[4 * j + 1], b4.y), 0.f);
                y.z = fmaxf(fmaf(a4.z, v[4 * j + 2], b4.z), 0.f), y.w = fmaxf(fmaf(a4.w, v[4 * j + 3], b4.w), 0.f);
                *reinterpret_cast<float4 *>(slab + ((j ^ (r & 7)) << 4)) = y;
              }
            }
          }
          fence_proxy_async();
          tc_fence_before();
          mbar_arrive(&bars->a_full[s][l]);
        } else {
          const int tile = blockIdx.x + (t0 + s) * gridDim.x;
          const long long g0 = (static_cast<long long>(tile) * BM + wq * 32) / p.ns;   // first group of this warp
          for (int c0 = share ? 32 * wg : 0; c0 < C; c0 += share ? 64 : 32) {
            float v[32];
            tc_ld32(d + c0, v);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const bool in = c0 + j < C;
              v[j] = fmaxf(fmaf(in ? fa[c0 + j] : 0.f, v[j], in ? fb[c0 + j] : 0.f), 0.f);
            }
            if (p.ns == 32) {
              group_colmax<false>(v, lane);
              if (g0 < p.G && c0 + lane < C) p.out[g0 * C + c0 + lane] = v[0];
            } else {   // ns == 16: lanes 0-15 and 16-31 are two groups
              group_colmax<true>(v, lane);
              const long long g = g0 + (lane >> 4);
              const int col = c0 + 2 * (lane & 15);
              if (g < p.G && col < C) {
                p.out[g * C + col] = v[0];
                if (col + 1 < C) p.out[g * C + col + 1] = v[1];
              }
            }
          }
          tc_fence_before();
          mbar_arrive(&bars->tile_done[s]);
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + kGatherWarps) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// W [N, K] row-major -> tile images [k_blocks][N][32] (one N tile of the whole width), 16-byte chunks
// XOR-swizzled by (row & 7), rounded to TF32, zero padded: the B operand of a chain layer.
__global__ void pack_weights_bn_kernel(const float *__restrict__ W, int N, int K, int ldw, int k_blocks,
                                       float *__restrict__ Wp) {
  const size_t total = static_cast<size_t>(k_blocks) * N * 32;
  for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<size_t>(gridDim.x) * blockDim.x) {
    const int e = static_cast<int>(i % 4), pc = static_cast<int>((i / 4) % 8);
    const int r = static_cast<int>((i / 32) % N), kb = static_cast<int>(i / (static_cast<size_t>(32) * N));
    const int k = kb * 32 + ((pc ^ (r & 7)) * 4) + e;
    const float v = k < K ? W[static_cast<size_t>(r) * ldw + k] : 0.f;
    uint32_t t;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(t) : "f"(v));
    Wp[i] = __uint_as_float(t);
  }
}

inline bool geom(int ns, int ct, int c1, int c2, int c3, Params *p) {
  if ((ns != 16 && ns != 32) || ct < 4 || ct % 4 != 0 || ct > 160) return false;
  const int c[4] = {(ct + 7) / 8 * 8, c1, c2, c3};
  for (int l = 1; l <= 3; ++l)
    if (c[l] < 16 || c[l] % 16 != 0 || c[l] > (l < 3 ? 128 : 256)) return false;
  if (c1 % 8 != 0 || c2 % 8 != 0) return false;
  size_t off = 0;
  int cols = 0;
  for (int l = 0; l < 3; ++l) {
    p->c[l] = c[l];
    p->kb[l] = (c[l] + 31) / 32;
    p->a_off[l] = static_cast<int>(off);
    off += static_cast<size_t>(p->kb[l]) * SLAB;
    p->d_off[l] = cols;
    cols += (c[l + 1] + 31) / 32 * 32;
  }
  p->c[3] = c[3];
  p->slot_bytes = static_cast<int>(off);
  p->slot_cols = cols;
  size_t wbytes = 0;
  for (int l = 0; l < 3; ++l) {
    p->w_bytes[l] = p->kb[l] * c[l + 1] * 128;
    wbytes += (p->w_bytes[l] + 1023) / 1024 * 1024;
  }
  const size_t tail = sizeof(float) * 2 * (c1 + c2 + c3) + 16 + sizeof(Barriers);
  p->slots = (2 * cols <= 512 && 2 * off + wbytes + tail + 1024 <= kSmemMax) ? 2 : 1;
  if (p->slots * cols > 512 || p->slots * off + wbytes + tail + 1024 > kSmemMax) return false;
  p->tmem_cols = 32;
  while (p->tmem_cols < p->slots * cols) p->tmem_cols *= 2;
  size_t o = static_cast<size_t>(p->slots) * off;
  for (int l = 0; l < 3; ++l) {
    p->w_off[l] = static_cast<int>(o);
    o += (p->w_bytes[l] + 1023) / 1024 * 1024;
  }
  p->fold_off = static_cast<int>(o);
  o += sizeof(float) * 2 * (c1 + c2 + c3);
  o = (o + 15) / 16 * 16;
  p->bars_off = static_cast<int>(o);
  return true;
}

}  // namespace chain
}  // namespace pn2

using namespace pn2;

// 1 when pn2_sa_chain_eval supports the block: nsample 16 or 32; input row width ct <= 160 (multiple of 4);
// layer widths multiples of 16, c1, c2 <= 128, c3 <= 256; operand tiles + weights within shared memory
extern "C" int pn2_sa_chain_supported(int ns, int ct, int c1, int c2, int c3) {
  chain::Params p;
  return chain::geom(ns, ct, c1, c2, c3, &p) ? 1 : 0;
}
extern "C" int pn2_sa_chain_weight_floats(int n, int k) { return (k + 31) / 32 * n * 32; }

extern "C" int pn2_mlp_pack_weights_bn(int n, int k, int ldw, const float *w, float *packed, void *stream) {
  if (n <= 0 || k <= 0 || ldw < k) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(w, packed);
  const int kb = (k + 31) / 32;
  const size_t total = static_cast<size_t>(kb) * n * 32;
  int blocks = static_cast<int>((total + 255) / 256);
  if (blocks > 4 * kNumSMs) blocks = 4 * kNumSMs;
  chain::pack_weights_bn_kernel<<<blocks, 256, 0, static_cast<cudaStream_t>(stream)>>>(w, n, k, ldw, kb, packed);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_sa_chain_eval(int b, int n, int npoint, int ns, int ct, const float *table, const float *new_xyz,
                                 const int *idx, int c1, int c2, int c3, const float *w1p, const float *w2p,
                                 const float *w3p, const float *fold, float *out, void *stream) {
  if (b <= 0 || n <= 0 || npoint <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(table, new_xyz, idx, w1p, w2p, w3p, fold, out);
  chain::Params p;
  if (!chain::geom(ns, ct, c1, c2, c3, &p)) return PN2_ERR_UNSUPPORTED;
  if (!aligned16(table) || !aligned16(w1p) || !aligned16(w2p) || !aligned16(w3p)) return PN2_ERR_UNSUPPORTED;
  const long long rows = static_cast<long long>(b) * npoint * ns;
  if (rows > 0x7fffffffLL) return PN2_ERR_UNSUPPORTED;
  p.table = table, p.new_xyz = new_xyz, p.idx = idx, p.fold = fold, p.out = out;
  p.w[0] = w1p, p.w[1] = w2p, p.w[2] = w3p;
  p.P = static_cast<int>(rows), p.G = b * npoint, p.n = n, p.npoint = npoint, p.ns = ns, p.ct = ct;
  const size_t smem = static_cast<size_t>(p.bars_off) + sizeof(chain::Barriers);
  cudaError_t e = cudaFuncSetAttribute(chain::sa_chain_eval_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  const int m_tiles = (p.P + chain::BM - 1) / chain::BM;
  const int grid = m_tiles < kNumSMs ? m_tiles : kNumSMs;
  chain::sa_chain_eval_kernel<<<grid, chain::kThreads, smem, static_cast<cudaStream_t>(stream)>>>(p);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
