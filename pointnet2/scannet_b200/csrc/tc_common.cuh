// tcgen05 / TMEM / mbarrier PTX wrappers and UMMA descriptor builders shared by the tensor-core
// kernels (mlp_gemm.cu: forward layer, mlp_bwd.cu: fused backward layer).  sm_100a only.
#pragma once
#include <cuda.h>

#include "pn2_common.cuh"

namespace pn2 {
namespace tc {

// ------------------------------------------------------------------ small PTX wrappers
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void *src, uint32_t src_bytes) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
// arrive on `bar` once all cp.async issued so far by this thread have landed
__device__ __forceinline__ void cp_async_arrive(uint64_t *bar) {
  asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t *bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                            uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
      "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
// 256-bit store (Blackwell STG.256): a full 32-byte sector per thread
__device__ __forceinline__ void st_global_v8(float *p, const float *v) {
  asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]),
               "f"(v[3]), "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ float to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

// cuTensorMapEncodeTiled through the runtime's driver entry point (the library links only libcudart)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
inline EncodeTiledFn encode_tiled() {
  static EncodeTiledFn fn = [] {
    void *f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}
// fp32 row-major [rows, cols] with `ld` floats between rows, boxes of box_rows x 32 columns (one
// 128-byte swizzle row per box row), SWIZZLE_128B.  Reads outside [rows, cols] are zero-filled, writes
// outside are dropped.
inline bool make_tmap_rows32(CUtensorMap *tm, const float *base, int rows, int cols, int ld, int box_rows = 32) {
  EncodeTiledFn enc = encode_tiled();
  if (enc == nullptr) return false;
  const cuuint64_t dims[2] = {static_cast<cuuint64_t>(cols), static_cast<cuuint64_t>(rows)};
  const cuuint64_t strides[1] = {static_cast<cuuint64_t>(ld) * 4};
  const cuuint32_t box[2] = {32, static_cast<cuuint32_t>(box_rows)}, estr[2] = {1, 1};
  return enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float *>(base), dims, strides, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// TMA tensor load of one box (its rows x 32 columns at (col0, row0)) into 1024-byte aligned shared
// memory; completes `bytes` = box bytes on `bar` (zero-filled parts included).  One thread calls.
__device__ __forceinline__ void tma_load_box(const CUtensorMap *tm, void *dst_smem, int col0, int row0, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst_smem)),
      "l"(reinterpret_cast<uint64_t>(tm)), "r"(smem_u32(bar)), "r"(col0), "r"(row0)
      : "memory");
}

// ---- TMA tensor store of one warp's 32 x 32 fp32 block (row = lane) ---------------------------------
// A thread that stores its own 128-byte row segment straight to global memory costs the LSU ~48
// wavefronts per STG.256 (32 lanes, 32 different lines); staged in shared memory with the 16-byte
// chunks XOR-swizzled by (row & 7) the same block is 8 conflict-free STS.128 per thread and ONE
// cp.async.bulk.tensor.2d (SWIZZLE_128B tensor map, box 32 x 32) that clips rows / columns past the
// tensor's extent.  `stg`: the warp's own 4 KB, 1024-byte aligned block.  All 32 lanes call.
__device__ __forceinline__ void tma_store_block32(const CUtensorMap *tm, unsigned char *stg, const float (&v)[32],
                                                  int lane, int col0, int row0) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // previous block read out
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j)
    *reinterpret_cast<float4 *>(stg + lane * 128 + ((j ^ (lane & 7)) << 4)) =
        make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
  fence_proxy_async();
  __syncwarp();
  if (lane == 0) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"(
                     reinterpret_cast<uint64_t>(tm)),
                 "r"(smem_u32(stg)), "r"(col0), "r"(row0)
                 : "memory");
    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
  }
}
// before the CTA exits (its shared memory is released) / before the data is consumed
__device__ __forceinline__ void tma_store_drain(int lane) {
  if (lane == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): rows of
// 128 bytes, 8-row swizzle atoms 1024 bytes apart; start address / offsets in 16-byte units.
__device__ __forceinline__ uint64_t make_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);       // start address, bits [0,14)
  d |= static_cast<uint64_t>(1) << 16;                          // leading byte offset (unused for SW128 K-major)
  d |= static_cast<uint64_t>(1024 >> 4) << 32;                  // stride byte offset: next 8-row atom
  d |= static_cast<uint64_t>(1) << 46;                          // descriptor version (Blackwell)
  d |= static_cast<uint64_t>(2) << 61;                          // layout type: SWIZZLE_128B
  return d;
}
// cute::UMMA::InstrDescriptor for kind::tf32, fp32 accumulate, both operands K-major
__host__ __device__ inline uint32_t make_idesc(int m, int n) {
  return (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(n >> 3) << 17) |
         (static_cast<uint32_t>(m >> 4) << 24);
}

// sum over the 32 lanes of v[c] for every c: after the butterfly lane l holds column l
__device__ __forceinline__ float warp_colsum32(float (&v)[32], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; ++i) {
      const float send = up ? v[i] : v[i + o];
      const float keep = up ? v[i + o] : v[i];
      v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
  }
  return v[0];
}


// MN-major descriptor for 32-bit (TF32) operands.  The tensor core transposes MN-major TF32 data at
// 32-byte granularity, so the ONLY shared-memory layout it accepts for them is "128-byte swizzle
// with 32-byte atomicity" (cute::UMMA::LayoutType::SWIZZLE_128B_BASE32B = 1; with the ordinary
// 16-byte-atom SWIZZLE_128B the instruction silently produces zeros -- measured).  Canonical
// layout in 16-byte units ((8,n),(4,k)):((1,LBO),(8,SBO)): 32 consecutive fp32 along M/N per
// 128-byte row, 4 K rows per 512-byte atom whose 32-byte chunks are XOR-ed with (row & 3)
// (Swizzle<2,5,2> on byte addresses), `sbo_bytes` between 4-row K groups, `lbo_bytes` between
// 32-wide M/N blocks.
__device__ __forceinline__ uint64_t make_desc_mn32(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= static_cast<uint64_t>((smem_addr & 0x3ffff) >> 4);
  d |= static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16;
  d |= static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32;
  d |= static_cast<uint64_t>(1) << 46;
  d |= static_cast<uint64_t>(1) << 61;  // SWIZZLE_128B_BASE32B
  return d;
}
// byte offset of the 16-byte chunk `chunk` (0..7) of row r in a [rows][128 B] slab with that swizzle
__device__ __forceinline__ uint32_t swz_base32(int r, int chunk) {
  return static_cast<uint32_t>(r * 128 + ((((chunk >> 1) ^ (r & 3)) << 5) | ((chunk & 1) << 4)));
}
// Blackwell packed fp32: two fused multiply-adds / multiplies per instruction (FFMA2 / FMUL2)
__device__ __forceinline__ float2 fma2(float2 a, float2 b, float2 c) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mov.b64 rc, {%6,%7}; "
      "fma.rn.f32x2 rd, ra, rb, rc; mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y), "f"(c.x), "f"(c.y));
  return r;
}
__device__ __forceinline__ float2 mul2(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rd; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; mul.rn.f32x2 rd, ra, rb; "
      "mov.b64 {%0,%1}, rd;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}

constexpr uint32_t kIdescAMajorMN = 1u << 15, kIdescBMajorMN = 1u << 16;

}  // namespace tc
}  // namespace pn2
