// Shared device/host helpers for libpn2_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pn2_b200.h"

#if defined(__CUDA_ARCH__) && (__CUDA_ARCH__ < 1000)
#error "libpn2_b200 is written for sm_100a (B200) only"
#endif

namespace pn2 {

constexpr int kNumSMs = 148;  // B200: 2 dies x 74 SMs

// Squared distance exactly as the reference evaluates it with FMA contraction
// off: ((dx*dx) + (dy*dy)) + (dz*dz), one rounding per operation
// (sampling_gpu.cu:103-104, ball_query_gpu.cu:31-32, interpolate_gpu.cu:33).
// The intrinsics are never contracted, whatever -fmad says.
__device__ __forceinline__ float sqdist_rn(float ax, float ay, float az,
                                           float bx, float by, float bz) {
  const float dx = __fsub_rn(ax, bx);
  const float dy = __fsub_rn(ay, by);
  const float dz = __fsub_rn(az, bz);
  return __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)),
                   __fmul_rn(dz, dz));
}

// ---- two distances per instruction: Blackwell packed fp32 (FADD2 / FFMA2) ------------
// Same IEEE operations as sqdist_rn on each half, so the results are bit-identical.  The
// squares are written as fma(d, d, +0): ptxas fuses a packed mul.rn + add.rn pair into one
// FFMA2 even with --fmad=false (which would skip a rounding), but it cannot fuse an fma
// into an add, and fma(d, d, +0) rounds exactly like mul(d, d).  build.py checks the SASS.
__device__ __forceinline__ float2 sub2_rn(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; sub.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 add2_rn(float2 a, float2 b) {
  float2 r;
  asm("{.reg .b64 ra, rb, rc; mov.b64 ra, {%2,%3}; mov.b64 rb, {%4,%5}; add.rn.f32x2 rc, ra, rb; "
      "mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(b.x), "f"(b.y));
  return r;
}
__device__ __forceinline__ float2 sq2_rn(float2 a) {
  float2 r;
  asm("{.reg .b64 ra, rz, rc; mov.b64 ra, {%2,%3}; mov.b64 rz, {%4,%4}; fma.rn.f32x2 rc, ra, ra, rz; "
      "mov.b64 {%0,%1}, rc;}"
      : "=f"(r.x), "=f"(r.y)
      : "f"(a.x), "f"(a.y), "f"(0.0f));
  return r;
}
// (ax.x,ay.x,az.x)-(bx.x,..) and (ax.y,..)-(bx.y,..): ((dx*dx)+(dy*dy))+(dz*dz) per half
__device__ __forceinline__ float2 sqdist2_rn(float2 ax, float2 ay, float2 az, float2 bx,
                                             float2 by, float2 bz) {
  const float2 dx = sub2_rn(ax, bx), dy = sub2_rn(ay, by), dz = sub2_rn(az, bz);
  return add2_rn(add2_rn(sq2_rn(dx), sq2_rn(dy)), sq2_rn(dz));
}

__device__ __forceinline__ float sqnorm_rn(float x, float y, float z) {
  return __fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)),
                   __fmul_rn(z, z));
}

// ---- spatial hash shared by the grid-based ball query and 3-NN ------------------------------
// Cell of a coordinate for cell size 1 / inv: floor(x * inv), a monotone function of x, clamped so
// that the integer arithmetic of the hash cannot overflow.  Two coordinates closer than the cell
// size land in the same or in adjacent cells (the callers keep a 1e-3 relative margin for the fp32
// rounding of x * inv).
__device__ __forceinline__ int grid_cell(float x, float inv) {
  return max(-(1 << 20), min(1 << 20, __float2int_rd(x * inv)));
}
__device__ __forceinline__ unsigned grid_hash(int ix, int iy, int iz, unsigned mask) {
  return ((static_cast<unsigned>(ix) * 73856093u) ^ (static_cast<unsigned>(iy) * 19349663u) ^
          (static_cast<unsigned>(iz) * 83492791u)) & mask;
}
// Counting sort of a cloud's points into hash buckets, in shared memory (order inside a bucket is
// arbitrary).  cnt[T] must be zero on entry and is used as the fill cursor; start[T + 1] receives
// the bucket offsets; sorted[n] the points as (x, y, z, bits(index)).  All threads of the CTA call.
template <int kThreads>
__device__ __forceinline__ void grid_build(const float *__restrict__ pts, int n, float inv, int T, int *cnt,
                                           int *start, float4 *sorted, int *scratch) {
  const int tid = threadIdx.x;
  for (int k = tid; k < n; k += kThreads) {
    const float x = __ldg(pts + 3 * k), y = __ldg(pts + 3 * k + 1), z = __ldg(pts + 3 * k + 2);
    atomicAdd(&cnt[grid_hash(grid_cell(x, inv), grid_cell(y, inv), grid_cell(z, inv), T - 1)], 1);
  }
  __syncthreads();
  // exclusive scan of cnt[0..T): each thread owns T / kThreads consecutive buckets
  const int per = (T + kThreads - 1) / kThreads;
  int local = 0;
  for (int i = 0; i < per; ++i) {
    const int b = tid * per + i;
    if (b < T) local += cnt[b];
  }
  int incl = local;
  const int lane = tid & 31, warp = tid >> 5;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (lane == 31) scratch[warp] = incl;
  __syncthreads();
  if (warp == 0) {
    int w = lane < kThreads / 32 ? scratch[lane] : 0;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int v = __shfl_up_sync(0xffffffffu, w, o);
      if (lane >= o) w += v;
    }
    scratch[lane] = w;  // inclusive over warps
  }
  __syncthreads();
  int run = incl - local + (warp > 0 ? scratch[warp - 1] : 0);
  for (int i = 0; i < per; ++i) {
    const int b = tid * per + i;
    if (b < T) {
      const int c = cnt[b];
      start[b] = run;
      cnt[b] = 0;
      run += c;
    }
  }
  if (tid == kThreads - 1) start[T] = n;
  __syncthreads();
  for (int k = tid; k < n; k += kThreads) {
    const float x = __ldg(pts + 3 * k), y = __ldg(pts + 3 * k + 1), z = __ldg(pts + 3 * k + 2);
    const unsigned b = grid_hash(grid_cell(x, inv), grid_cell(y, inv), grid_cell(z, inv), T - 1);
    sorted[start[b] + atomicAdd(&cnt[b], 1)] = make_float4(x, y, z, __int_as_float(k));
  }
  __syncthreads();
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared::cta) -----------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t phase) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t"
      "}" ::"r"(smem_u32(bar)),
      "r"(phase)
      : "memory");
}
// non-blocking: has the phase with this parity completed?
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t phase) {
  uint32_t ok;
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t"
      "}"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(phase)
      : "memory");
  return ok != 0;
}
// bytes must be a multiple of 16; src and dst 16-byte aligned.
__device__ __forceinline__ void tma_load_1d(void *dst_smem, const void *src_gmem,
                                            uint32_t bytes, uint64_t *bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes "
      "[%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
      "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- streaming global accesses --------------------------------------------
__device__ __forceinline__ void st_cs_f4(float *p, float4 v) {
  asm volatile("st.global.cs.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
               "f"(v.y), "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float4 ld_cs_f4(const float *p) {
  float4 v;
  asm volatile("ld.global.cs.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ int4 ld_nc_i4(const int *p) {
  int4 v;
  asm volatile("ld.global.nc.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}

// ---- programmatic dependent launch --------------------------------------------------------
// A training step is ~300 dependent launches, most of them a few microseconds long; what a kernel boundary
// costs inside the replayed graph is the drain of the previous grid's last CTAs plus the next grid's ramp.
// Kernels launched through pn2::launch can carry the programmatic-stream-serialization attribute: their CTAs may be
// scheduled while the previous kernel of the stream is still draining.  Every such kernel calls pdl_enter()
// FIRST: griddepcontrol.wait blocks until the previous grid has completed and its memory operations are visible
// (so nothing is read or written early), then griddepcontrol.launch_dependents lets the NEXT kernel's CTAs be
// scheduled as soon as every CTA of this grid has started.  (The trigger sits after the wait on purpose: a chain
// of kernels cannot run more than one grid ahead.)  Without the attribute both instructions are no-ops.
__device__ __forceinline__ void pdl_enter() {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

// ---- host side --------------------------------------------------------------
void count_launch(int n = 1);
bool pdl_enabled();   // PN2_PDL=1 (lib.cu; measured slower on the replayed step graph, so off by default)

template <class... KArgs, class... Args>
inline cudaError_t launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                          Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid, cfg.blockDim = block, cfg.dynamicSmemBytes = smem, cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<Args &&>(args)...);
}

// The reference's opt_n_threads (_ext_src/include/cuda_utils.h:15-19), same
// double-precision log quotient so that the block size -- which fixes the FPS
// tie-break -- is the one the reference would have launched.
int ref_block_threads(int work);

inline bool aligned16(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}
inline bool aligned4(const void *p) {
  return (reinterpret_cast<uintptr_t>(p) & 3u) == 0;
}

#define PN2_CHECK_PTRS(...)                                  \
  do {                                                       \
    const void *ptrs__[] = {__VA_ARGS__};                    \
    for (const void *q__ : ptrs__) {                         \
      if (q__ == nullptr) return PN2_ERR_INVALID_ARG;        \
      if (!pn2::aligned4(q__)) return PN2_ERR_ALIGNMENT;     \
    }                                                        \
  } while (0)

#define PN2_RETURN_LAUNCH_STATUS()          \
  do {                                      \
    cudaError_t e__ = cudaGetLastError();   \
    return static_cast<int>(e__);           \
  } while (0)

}  // namespace pn2
