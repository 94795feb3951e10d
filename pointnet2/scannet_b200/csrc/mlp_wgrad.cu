// Weight gradient of one shared-MLP layer on the 5th-generation tensor cores (tcgen05 + TMEM):
//
//     dW[N, K] = X[P, N]^T * Y[P, K]          X = dZ (gradient wrt the layer's pre-BN output),
//                                             Y = H_prev (the layer's activated input rows)
//
// the `dZ^T H` product of autograd's Conv2d(1x1) adjoint for the layers whose width does not fit the
// one-pass fused backward kernel (mlp_bwd.cu: N <= 128, lda <= 160): the coarse levels of the
// reference's models -- SA3 / SA4 [.., 256] / [.., 512], FP [.., 256, 256]
// (/root/reference/pointnet2/pointnet2_semseg.py:32-34, 112-114).  Their P is small (<= 65536 rows, the
// tensors are L2-resident), so the reduction over rows is split across CTAs.
//
// Both operands are read MN-major: the product reduces over ROWS, and a row-major [32 rows][32 fp32]
// box IS the MN-major UMMA operand layout when its 32-byte chunks are XOR-swizzled by (row & 3)
// (SWIZZLE_128B_BASE32B, the only layout the tensor core accepts for MN-major TF32 data).  The TMA unit
// writes exactly that image (CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B), so the operands go
// HBM/L2 -> shared memory -> tensor core without touching a register: one thread issues the box loads,
// one thread issues the MMAs, four warps read the accumulator out at the end.
//
// Grid (row split s, 128-row tile of dW, column tile of dW); per-CTA partial images
// part[s][n_tiles * 128][k_tiles * KT], summed in fixed order by pn2_sum_partials (deterministic).
// Precision: the tensor core reads the upper 19 bits of each fp32 word (TF32 by truncation), fp32
// accumulation -- the arithmetic class of the reference's own convolutions.
#include <stdlib.h>
#include <string.h>

#include "tc_common.cuh"

namespace pn2 {
namespace wgrad {

using namespace tc;

constexpr int RB = 32;            // rows per pipeline stage (4 MMAs of K = 8 rows)
constexpr int NT = 128;           // rows of dW per CTA = UMMA M
constexpr int kMaxKT = 256;       // columns of dW per CTA = UMMA N
constexpr int SLAB = RB * 128;    // 4 KB: [32 rows][32 fp32]
constexpr int kXSlabs = NT / 32;  // 4
constexpr int kMaxStages = 8;
constexpr int kEpiWarps = 4;
constexpr int kThreads = 32 * (kEpiWarps + 2);  // + TMA producer warp + MMA warp
constexpr int kSmemBudget = 100 * 1024;   // half an SM: see geom()

struct Params {
  const float *X, *Y;   // [P, ldx], [P, ldy] (fallback loader only; the TMA path uses the tensor maps)
  float *part;          // [splits][n_rows][k_cols]
  int P, N, ldx, ldy;
  int units, per_cta;   // 32-row units in all / per row split
  int KT, yslabs;       // dW columns per CTA, = 32 * yslabs
  int n_rows, k_cols;   // padded extents of a partial image
  int stages, tmem_cols;
  int tma;              // operands by TMA box loads (else: the epilogue warps copy them, LDG -> swizzled STS)
};

struct __align__(8) Barriers {
  uint64_t full[kMaxStages];   // stage landed      (TMA: 1 arrival + tx bytes; fallback: 128 arrivals)
  uint64_t empty[kMaxStages];  // its MMAs retired  (1 commit)
  uint64_t acc_full;           // every MMA retired (1 commit)
  uint32_t tmem_base;
};

__global__ void __launch_bounds__(kThreads, 1) wgrad_tf32_kernel(Params p, const __grid_constant__ CUtensorMap tm_x,
                                                                 const __grid_constant__ CUtensorMap tm_y) {
  pdl_enter();
  extern __shared__ __align__(1024) unsigned char smem[];
  const int stage_bytes = (kXSlabs + p.yslabs) * SLAB;
  Barriers *bars = reinterpret_cast<Barriers *>(smem + p.stages * stage_bytes);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int split = blockIdx.x, n_tile = blockIdx.y, k_tile = blockIdx.z;
  const int u0 = split * p.per_cta;
  const int u1 = min(p.units, u0 + p.per_cta);
  const int my_units = max(0, u1 - u0);

  if (tid == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(&bars->full[s], p.tma ? 1 : 32 * kEpiWarps);
      mbar_init(&bars->empty[s], 1);
    }
    mbar_init(&bars->acc_full, 1);
    mbar_fence_init();
  }
  if (warp == kEpiWarps + 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(&bars->tmem_base)),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = bars->tmem_base;

  if (warp == kEpiWarps) {
    // ================================ TMA producer (one thread) ================================
    if (p.tma && lane == 0) {
      int s = 0, par = 1;
      for (int u = u0; u < u1; ++u) {
        mbar_wait(&bars->empty[s], par);
        unsigned char *st = smem + s * stage_bytes;
        mbar_expect_tx(&bars->full[s], static_cast<uint32_t>(stage_bytes));
        // rows past P and columns past N / ldy are zero-filled by the TMA unit (and still counted)
#pragma unroll
        for (int j = 0; j < kXSlabs; ++j) tma_load_box(&tm_x, st + j * SLAB, n_tile * NT + 32 * j, u * RB, &bars->full[s]);
        for (int h = 0; h < p.yslabs; ++h)
          tma_load_box(&tm_y, st + (kXSlabs + h) * SLAB, k_tile * p.KT + 32 * h, u * RB, &bars->full[s]);
        if (++s == p.stages) s = 0, par ^= 1;
      }
    }
  } else if (warp == kEpiWarps + 1) {
    // ======================================= MMA issuer =======================================
    const bool leader = lane == 0;
    const uint32_t idesc = make_idesc(NT, p.KT) | kIdescAMajorMN | kIdescBMajorMN;
    const uint64_t mn_hi = make_desc_mn32(0, SLAB, 512);  // LBO: next 32-wide M/N block; SBO: next 4 rows
    auto addr = [](uint32_t a) { return static_cast<uint64_t>((a & 0x3ffff) >> 4); };
    int s = 0, par = 0;
    for (int u = 0; u < my_units; ++u) {
      mbar_wait(&bars->full[s], par);
      tc_fence_after();
      const uint64_t da = mn_hi | addr(smem_u32(smem + s * stage_bytes));
      const uint64_t db = mn_hi | addr(smem_u32(smem + s * stage_bytes + kXSlabs * SLAB));
      if (leader) {
#pragma unroll
        for (int ks = 0; ks < RB / 8; ++ks)  // 8 rows (two 4-row swizzle atoms, 1024 bytes) per instruction
          tc_mma_tf32(tmem_base, da + 64 * ks, db + 64 * ks, idesc, (u | ks) != 0);
        tc_commit(&bars->empty[s]);
      }
      __syncwarp();
      if (++s == p.stages) s = 0, par ^= 1;
    }
    if (leader) tc_commit(&bars->acc_full);
  } else {
    // =============================== epilogue (and fallback loaders) ===============================
    const int wq = warp;
    if (!p.tma) {
      // Every 16-byte chunk of the stage: slab q (X: columns of this CTA's dW rows, Y: of its dW columns),
      // row r, chunk c -> swizzled position.  128 threads, 8 chunks per row.
      int s = 0, par = 1;
      const int slabs = kXSlabs + p.yslabs;
      for (int u = u0; u < u1; ++u) {
        mbar_wait(&bars->empty[s], par);
        unsigned char *st = smem + s * stage_bytes;
        for (int i = tid; i < slabs * RB * 8; i += 32 * kEpiWarps) {
          const int c = i & 7, r = (i >> 3) & (RB - 1), q = i >> 8;
          const int grow = u * RB + r;
          float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
          if (q < kXSlabs) {
            const int col = n_tile * NT + 32 * q + 4 * c;
            if (grow < p.P && col < p.N) v = __ldg(reinterpret_cast<const float4 *>(p.X + static_cast<size_t>(grow) * p.ldx + col));
          } else {
            const int col = k_tile * p.KT + 32 * (q - kXSlabs) + 4 * c;
            if (grow < p.P && col < p.ldy) v = __ldg(reinterpret_cast<const float4 *>(p.Y + static_cast<size_t>(grow) * p.ldy + col));
          }
          *reinterpret_cast<float4 *>(st + q * SLAB + swz_base32(r, c)) = v;
        }
        fence_proxy_async();
        mbar_arrive(&bars->full[s]);
        if (++s == p.stages) s = 0, par ^= 1;
      }
    }
    mbar_wait(&bars->acc_full, 0);
    tc_fence_after();
    // TMEM lanes = rows n of dW, columns = its columns k
    float *dst = p.part + (static_cast<size_t>(split) * p.n_rows + n_tile * NT + wq * 32 + lane) * p.k_cols + k_tile * p.KT;
    for (int c = 0; c < p.yslabs; ++c) {
      float v[32];
      if (my_units > 0) {
        tc_ld32(tmem_base + (static_cast<uint32_t>(wq * 32) << 16) + c * 32, v);
      } else {
#pragma unroll
        for (int j = 0; j < 32; ++j) v[j] = 0.f;
      }
#pragma unroll
      for (int j = 0; j < 32; j += 8) st_global_v8(dst + c * 32 + j, v + j);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == kEpiWarps + 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base),
                 "r"(static_cast<uint32_t>(p.tmem_cols))
                 : "memory");
  }
}

// fp32 row-major [rows, cols] -> boxes of 32 rows x 32 columns written as the MN-major UMMA image
// (32-byte chunks XOR-swizzled by (row & 3)); reads outside [rows, cols] are zero-filled.
inline bool make_tmap_mn32(CUtensorMap *tm, const float *base, int rows, int cols, int ld) {
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(RB)}, estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B_ATOM_32B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Geom {
  int n_tiles, k_tiles, KT, yslabs, units, splits, per_cta, stages, tmem_cols;
  size_t smem;
};

inline bool geom(int p, int n, int ldy, Geom *g) {
  if (p <= 0 || n <= 0 || ldy <= 0 || n % 4 != 0 || ldy % 4 != 0) return false;
  g->n_tiles = (n + NT - 1) / NT;
  const int kslabs = (ldy + 31) / 32;
  g->k_tiles = (kslabs * 32 + kMaxKT - 1) / kMaxKT;
  g->yslabs = (kslabs + g->k_tiles - 1) / g->k_tiles;
  g->KT = g->yslabs * 32;
  g->units = (p + RB - 1) / RB;
  // row splits: cover the machine, but at least 4 units (128 rows) per CTA
  int splits = kNumSMs / (g->n_tiles * g->k_tiles);
  if (splits < 1) splits = 1;
  const int max_splits = (g->units + 3) / 4;
  if (splits > max_splits) splits = max_splits;
  g->per_cta = (g->units + splits - 1) / splits;
  g->splits = (g->units + g->per_cta - 1) / g->per_cta;
  g->tmem_cols = 32;
  while (g->tmem_cols < g->KT) g->tmem_cols *= 2;
  const int stage_bytes = (kXSlabs + g->yslabs) * SLAB;
  // The pipeline's shared-memory budget (PN2_WGRAD_SMEM_KB overrides): the kernel runs on the weight-gradient branch
  // BESIDE the critical path, and a CTA that takes a whole SM's shared memory keeps the main stream's CTAs out until
  // it retires (ordered trace: a 41 us hole in front of the first coarse-level kernel).  Measured in one call:
  // 3.442 / 3.446 ms per step with 200 KB, 3.434 / 3.439 ms with 100 KB, 3.441 ms with 64 KB.
  static const int budget = getenv("PN2_WGRAD_SMEM_KB") != nullptr ? atoi(getenv("PN2_WGRAD_SMEM_KB")) * 1024 : kSmemBudget;
  g->stages = budget / stage_bytes;
  if (g->stages < 2) g->stages = 2;
  if (g->stages > kMaxStages) g->stages = kMaxStages;
  if (g->stages > g->per_cta) g->stages = g->per_cta < 2 ? 2 : g->per_cta;
  g->smem = static_cast<size_t>(g->stages) * stage_bytes + sizeof(Barriers);
  return true;
}

}  // namespace wgrad
}  // namespace pn2

using namespace pn2;

// geometry of the partial images: part[splits][rows][cols] floats; dW = the leading [n, k] block of their sum
extern "C" int pn2_mlp_wgrad_splits(int p, int n, int ldy) {
  wgrad::Geom g;
  return wgrad::geom(p, n, ldy, &g) ? g.splits : 0;
}
extern "C" int pn2_mlp_wgrad_rows(int n) { return (n + wgrad::NT - 1) / wgrad::NT * wgrad::NT; }
extern "C" int pn2_mlp_wgrad_cols(int ldy) {
  wgrad::Geom g;
  return wgrad::geom(128, 128, ldy, &g) ? g.k_tiles * g.KT : 0;
}

extern "C" int pn2_mlp_wgrad(int p, int n, int ldx, int ldy, const float *x, const float *y, float *part,
                             void *stream) {
  wgrad::Geom g;
  if (ldx < n || !wgrad::geom(p, n, ldy, &g) || ldx % 4 != 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(x, y, part);
  if (!aligned16(x) || !aligned16(y) || !aligned16(part)) return PN2_ERR_UNSUPPORTED;
  wgrad::Params prm;
  prm.X = x, prm.Y = y, prm.part = part;
  prm.P = p, prm.N = n, prm.ldx = ldx, prm.ldy = ldy;
  prm.units = g.units, prm.per_cta = g.per_cta, prm.KT = g.KT, prm.yslabs = g.yslabs;
  prm.n_rows = g.n_tiles * wgrad::NT, prm.k_cols = g.k_tiles * g.KT;
  prm.stages = g.stages, prm.tmem_cols = g.tmem_cols;
  CUtensorMap tm_x, tm_y;
  memset(&tm_x, 0, sizeof(tm_x)), memset(&tm_y, 0, sizeof(tm_y));
  const bool no_tma = getenv("PN2_WGRAD_NO_TMA") != nullptr;  // the register-path loaders (tests / timing)
  prm.tma = (!no_tma && wgrad::make_tmap_mn32(&tm_x, x, p, n, ldx) && wgrad::make_tmap_mn32(&tm_y, y, p, ldy, ldy)) ? 1 : 0;
  cudaError_t e = cudaFuncSetAttribute(wgrad::wgrad_tf32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       static_cast<int>(g.smem));
  if (e != cudaSuccess) return static_cast<int>(e);
  dim3 grid(g.splits, g.n_tiles, g.k_tiles);
  pn2::launch(wgrad::wgrad_tf32_kernel, grid, wgrad::kThreads, g.smem, static_cast<cudaStream_t>(stream), prm, tm_x, tm_y);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
