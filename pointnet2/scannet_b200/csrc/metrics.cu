// Evaluation metrics on the device for sm_100a (the solver loop's per-iteration host round trip).
//
// The reference scores every training / validation iteration on the host
// (/root/reference/lib/solver.py:327-347 -> scripts/eval.py:39-123): four .cpu().numpy() copies, then
// filter_points -- a Python loop that builds hash(str(x) + str(y) + str(z)) for each of the B * N points
// (:42) -- np.unique, a voxelisation at 2 cm (lib/pc_util.py:39-51) and per-class np.sum / intersect1d /
// union1d passes.  Everything those functions return is a function of two confusion matrices:
//   points : one count per DISTINCT coordinate (its first occurrence in the batch), target x prediction
//   voxels : one count per occupied 2 cm voxel, labelled by one representative point
// so the device path is: exact-duplicate removal and voxel representatives by lock-free hash tables
// (open addressing, the slot keeps the LOWEST point index of its key via atomicMin), two 2-D histograms,
// and a handful of 21-element operations.  No host synchronisation.
//
// Integer / byte work bounded by L2 atomics and launch latency; nothing here is GEMM-shaped.
#include "pn2_common.cuh"

namespace pn2 {

constexpr int kMetThreads = 256;

__device__ __forceinline__ unsigned mix32(unsigned h) {
  h ^= h >> 16, h *= 0x7feb352du, h ^= h >> 15, h *= 0x846ca68bu, h ^= h >> 16;
  return h;
}

template <int W>
__device__ __forceinline__ unsigned key_hash(const unsigned *__restrict__ k) {
  unsigned h = 0x9e3779b9u;
#pragma unroll
  for (int w = 0; w < W; ++w) h = mix32(h ^ k[w]) + 0x9e3779b9u * (w + 1);
  return h;
}

template <int W>
__device__ __forceinline__ bool key_eq(const unsigned *__restrict__ keys, int i, const unsigned *__restrict__ mine) {
  bool eq = true;
#pragma unroll
  for (int w = 0; w < W; ++w) eq = eq && __ldg(keys + static_cast<size_t>(i) * W + w) == mine[w];
  return eq;
}

// table[slot] = lowest index i (among the active points) whose key lives in that slot; -1 = empty.
// A slot's key is the key of whichever point it currently names; atomicMin only ever replaces a point by
// another one with the same key, so the key of an occupied slot never changes.
template <int W>
__global__ void __launch_bounds__(kMetThreads)
first_occurrence_insert_kernel(int p, const unsigned *__restrict__ keys, const int *__restrict__ active_rep,
                               int *__restrict__ table, unsigned mask) {
  const int i = blockIdx.x * kMetThreads + threadIdx.x;
  if (i >= p || (active_rep != nullptr && __ldg(active_rep + i) != i)) return;
  unsigned mine[W];
#pragma unroll
  for (int w = 0; w < W; ++w) mine[w] = __ldg(keys + static_cast<size_t>(i) * W + w);
  unsigned h = key_hash<W>(mine) & mask;
  for (;;) {
    int cur = *reinterpret_cast<volatile int *>(table + h);
    if (cur == -1) {
      cur = atomicCAS(table + h, -1, i);
      if (cur == -1) return;
    }
    if (key_eq<W>(keys, cur, mine)) {
      atomicMin(table + h, i);
      return;
    }
    h = (h + 1) & mask;
  }
}

// rep[i] = the lowest active index with i's key (i itself for a first occurrence); -1 for inactive points
template <int W>
__global__ void __launch_bounds__(kMetThreads)
first_occurrence_lookup_kernel(int p, const unsigned *__restrict__ keys, const int *__restrict__ active_rep,
                               const int *__restrict__ table, unsigned mask, int *__restrict__ rep) {
  const int i = blockIdx.x * kMetThreads + threadIdx.x;
  if (i >= p) return;
  if (active_rep != nullptr && __ldg(active_rep + i) != i) {
    rep[i] = -1;
    return;
  }
  unsigned mine[W];
#pragma unroll
  for (int w = 0; w < W; ++w) mine[w] = __ldg(keys + static_cast<size_t>(i) * W + w);
  unsigned h = key_hash<W>(mine) & mask;
  for (;;) {
    const int cur = __ldg(table + h);
    if (cur == -1 || key_eq<W>(keys, cur, mine)) {   // (-1 cannot happen after the insert pass)
      rep[i] = cur;
      return;
    }
    h = (h + 1) & mask;
  }
}

// order-preserving map float -> unsigned, so that atomicMin / atomicMax work on coordinates
__device__ __forceinline__ unsigned f2ord(float f) {
  const unsigned u = __float_as_uint(f);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ord2f(unsigned o) {
  return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}

// bounds[0..2] = min, bounds[3..5] = max of the coordinates, as ordered unsigned (init: 0xffffffff / 0)
__global__ void __launch_bounds__(kMetThreads)
coord_bounds_kernel(int p, const float *__restrict__ coords, unsigned *__restrict__ bounds) {
  __shared__ unsigned s_b[6];
  if (threadIdx.x < 6) s_b[threadIdx.x] = threadIdx.x < 3 ? 0xffffffffu : 0u;
  __syncthreads();
  unsigned lo[3] = {0xffffffffu, 0xffffffffu, 0xffffffffu}, hi[3] = {0u, 0u, 0u};
  for (int i = blockIdx.x * kMetThreads + threadIdx.x; i < p; i += gridDim.x * kMetThreads) {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const unsigned o = f2ord(__ldg(coords + 3 * static_cast<size_t>(i) + d));
      lo[d] = min(lo[d], o), hi[d] = max(hi[d], o);
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    lo[d] = __reduce_min_sync(0xffffffffu, lo[d]);
    hi[d] = __reduce_max_sync(0xffffffffu, hi[d]);
  }
  if ((threadIdx.x & 31) == 0) {
#pragma unroll
    for (int d = 0; d < 3; ++d) atomicMin(&s_b[d], lo[d]), atomicMax(&s_b[3 + d], hi[d]);
  }
  __syncthreads();
  if (threadIdx.x < 3) atomicMin(bounds + threadIdx.x, s_b[threadIdx.x]);
  else if (threadIdx.x < 6) atomicMax(bounds + threadIdx.x, s_b[threadIdx.x]);
}

// lib/pc_util.py:40-44 in the float32 arithmetic NumPy uses there (float32 coordinates, `res` a weak Python
// scalar):  nvox = ceil((max - min) / res); v = ceil((p - min) / res); key = (v0 + v1 * nvox0) + (v2 * nvox0) * nvox1
__global__ void __launch_bounds__(kMetThreads)
voxel_key_kernel(int p, const float *__restrict__ coords, const unsigned *__restrict__ bounds, float res,
                 unsigned *__restrict__ vkey) {
  const int i = blockIdx.x * kMetThreads + threadIdx.x;
  if (i >= p) return;
  float mn[3], nv[2], v[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) mn[d] = ord2f(__ldg(bounds + d));
#pragma unroll
  for (int d = 0; d < 2; ++d) nv[d] = ceilf(__fdiv_rn(__fsub_rn(ord2f(__ldg(bounds + 3 + d)), mn[d]), res));
#pragma unroll
  for (int d = 0; d < 3; ++d)
    v[d] = ceilf(__fdiv_rn(__fsub_rn(__ldg(coords + 3 * static_cast<size_t>(i) + d), mn[d]), res));
  const float key = __fadd_rn(__fadd_rn(v[0], __fmul_rn(v[1], nv[0])), __fmul_rn(__fmul_rn(v[2], nv[0]), nv[1]));
  vkey[i] = __float_as_uint(__fadd_rn(key, 0.0f));   // -0 and +0 are one voxel (np.unique compares values)
}

// conf[t * nc + q] += 1 for every representative point i (rep[i] == i) with target t and prediction q
__global__ void __launch_bounds__(kMetThreads)
confusion_kernel(int p, const int *__restrict__ rep, const long long *__restrict__ preds,
                 const long long *__restrict__ targets, int nc, int *__restrict__ conf, int *__restrict__ bad) {
  extern __shared__ int s_conf[];
  for (int k = threadIdx.x; k < nc * nc; k += kMetThreads) s_conf[k] = 0;
  __syncthreads();
  for (int i = blockIdx.x * kMetThreads + threadIdx.x; i < p; i += gridDim.x * kMetThreads) {
    if (__ldg(rep + i) != i) continue;
    const long long t = targets[i], q = preds[i];
    if (t < 0 || t >= nc || q < 0 || q >= nc) {
      atomicAdd(bad, 1);
      continue;
    }
    atomicAdd(&s_conf[static_cast<int>(t) * nc + static_cast<int>(q)], 1);
  }
  __syncthreads();
  for (int k = threadIdx.x; k < nc * nc; k += kMetThreads)
    if (s_conf[k]) atomicAdd(conf + k, s_conf[k]);
}

}  // namespace pn2

using namespace pn2;

static unsigned table_mask(int p) {
  unsigned n = 1024;
  while (n < 2u * static_cast<unsigned>(p)) n <<= 1;
  return n - 1;
}

extern "C" int pn2_metrics_table_ints(int p) { return p > 0 ? static_cast<int>(table_mask(p) + 1) : 0; }

// rep[i] = index of the first point with the same (bitwise) coordinates; table: pn2_metrics_table_ints(p) ints
extern "C" int pn2_first_occurrence_xyz(int p, const float *coords, int *table, int *rep, void *stream) {
  if (p <= 0 || p > (1 << 29)) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(coords, table, rep);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned mask = table_mask(p);
  const int blocks = (p + kMetThreads - 1) / kMetThreads;
  cudaMemsetAsync(table, 0xff, sizeof(int) * (static_cast<size_t>(mask) + 1), s);
  const unsigned *keys = reinterpret_cast<const unsigned *>(coords);
  first_occurrence_insert_kernel<3><<<blocks, kMetThreads, 0, s>>>(p, keys, nullptr, table, mask);
  first_occurrence_lookup_kernel<3><<<blocks, kMetThreads, 0, s>>>(p, keys, nullptr, table, mask, rep);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}

// vrep[i] = the first point (among those with rep[i] == i) in i's voxel of edge `res`, -1 for filtered points.
// bounds: 6 unsigned, vkey: p unsigned (scratch).
extern "C" int pn2_first_occurrence_voxel(int p, const float *coords, const int *rep, float res, unsigned *bounds,
                                          unsigned *vkey, int *table, int *vrep, void *stream) {
  if (p <= 0 || p > (1 << 29) || !(res > 0.f)) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(coords, rep, bounds, vkey, table, vrep);
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  const unsigned mask = table_mask(p);
  const int blocks = (p + kMetThreads - 1) / kMetThreads;
  cudaMemsetAsync(bounds, 0xff, 3 * sizeof(unsigned), s);
  cudaMemsetAsync(bounds + 3, 0x00, 3 * sizeof(unsigned), s);
  cudaMemsetAsync(table, 0xff, sizeof(int) * (static_cast<size_t>(mask) + 1), s);
  coord_bounds_kernel<<<min(blocks, 4 * kNumSMs), kMetThreads, 0, s>>>(p, coords, bounds);
  voxel_key_kernel<<<blocks, kMetThreads, 0, s>>>(p, coords, bounds, res, vkey);
  first_occurrence_insert_kernel<1><<<blocks, kMetThreads, 0, s>>>(p, vkey, rep, table, mask);
  first_occurrence_lookup_kernel<1><<<blocks, kMetThreads, 0, s>>>(p, vkey, rep, table, mask, vrep);
  count_launch(4);
  PN2_RETURN_LAUNCH_STATUS();
}

// conf [nc, nc] int32 (caller zero-fills) += target x prediction counts over the points with rep[i] == i;
// bad[0] += number of representatives with a label outside [0, nc)
extern "C" int pn2_confusion(int p, const int *rep, const long long *preds, const long long *targets, int nc,
                             int *conf, int *bad, void *stream) {
  if (p <= 0 || nc <= 0 || nc > 64) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(rep, preds, targets, conf, bad);
  if ((reinterpret_cast<uintptr_t>(preds) & 7u) != 0 || (reinterpret_cast<uintptr_t>(targets) & 7u) != 0)
    return PN2_ERR_ALIGNMENT;
  const int blocks = min((p + kMetThreads - 1) / kMetThreads, 2 * kNumSMs);
  confusion_kernel<<<blocks, kMetThreads, sizeof(int) * nc * nc, static_cast<cudaStream_t>(stream)>>>(
      p, rep, preds, targets, nc, conf, bad);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}
