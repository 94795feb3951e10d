// Chunk generation on the GPU for sm_100a: what feeds the SA/FP hot path.
//
// Replaces the NumPy code of /root/reference/lib/dataset.py that cuts training chunks and
// whole-scene tiles out of a preprocessed scene, with the scene resident in HBM:
//   * box_select_kernel      ScannetDataset.generate_chunks, one try (:220-238): the rows inside the
//                            crop box in scene order (an ordered stream compaction), the number of
//                            annotated rows and the number of occupied voxels of the 31 x 31 x 62 grid;
//                            ScannetDatasetWholeScene.__getitem__ (:319-322): the rows of every tile
//   * chunk_centre_kernel    _augment's chunk mean (:101)
//   * chunk_assemble_kernel  chunk[choices] (:248, :331), __getitem__'s column assembly (:66-80,
//                            :300-311), _augment's arithmetic (:102-141, 146-200), the label and
//                            sample-weight lookups (:90-91, :338-339) and collate_*'s casts (:379-381)
//
// The reference's scenes are float64 arrays whose columns hold float32-representable values
// (collect_scannet_scenes.py:64-79); they are stored here as float32 rows [n, ld] and every
// comparison / product is evaluated in double exactly as NumPy evaluates it, so selections, labels,
// weights and the un-augmented columns are bit-identical; the augmented coordinates differ from
// NumPy's only where its BLAS matmul / sequential mean round differently in the 53rd bit (i.e.
// almost never after the float32 cast: tests/test_chunker_gpu.py states the bound).
// The random draws stay on the host (np.random, the reference's order) and arrive as arguments.
//
// HBM-bound integer/byte work; nothing here is GEMM-shaped.  One CTA per box: a scene is at most a
// few MB and stays in L2 across the boxes of one launch.
#include "pn2_common.cuh"

namespace pn2 {

constexpr int kSelThreads = 1024;
constexpr int kVoxelWords = 2048;  // keys 0 .. 65535 (vx, vy <= 32, vz <= 63: 32*1922 + 32*62 + 63 = 63551)

struct CropSpec {          // generate_chunks' box, formed on the device from the centre row
  int centre_row;          // -1: boxes come from the lo / hi arrays
  double zmin, zmax;       // coordmin[2], coordmax[2] of the scene (:216-217, 224-225)
};

// Rows of `scene` inside [lo, hi] (inclusive, compared in double like NumPy's float32-vs-float64
// broadcast) in ascending row order -> sel[box, 0 .. count).  kStats: also count the rows with
// label > 0 and the distinct voxel keys of the rows inside [vmin - 0.01, vmax + 0.01].
template <bool kStats>
__global__ void __launch_bounds__(kSelThreads)
box_select_kernel(int n, int ld, const float *__restrict__ scene, CropSpec crop,
                  const double *__restrict__ lo_arr, const double *__restrict__ hi_arr, int label_col,
                  int *__restrict__ sel, int *__restrict__ stats, unsigned char *__restrict__ mask) {
  __shared__ int s_warp[kSelThreads / 32];
  __shared__ int s_base;
  __shared__ int s_lab, s_odd;
  __shared__ unsigned s_vox[kStats ? kVoxelWords : 1];
  const int box = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double lo[3], hi[3], vmin[3], vmax[3];
  if (crop.centre_row >= 0) {
    const float *c = scene + static_cast<size_t>(crop.centre_row) * ld;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double half = d == 2 ? 1.5 : 0.75;
      vmin[d] = __dsub_rn(static_cast<double>(__ldg(c + d)), half);   // curcenter - [0.75, 0.75, 1.5]
      vmax[d] = __dadd_rn(static_cast<double>(__ldg(c + d)), half);
    }
    vmin[2] = crop.zmin;
    vmax[2] = crop.zmax;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = __dsub_rn(vmin[d], 0.2);
      hi[d] = __dadd_rn(vmax[d], 0.2);
    }
  } else {
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      lo[d] = lo_arr[3 * box + d];
      hi[d] = hi_arr[3 * box + d];
      vmin[d] = lo[d];
      vmax[d] = hi[d];
    }
  }
  double ilo[3], ihi[3], span[3];
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    ilo[d] = __dsub_rn(vmin[d], 0.01);
    ihi[d] = __dadd_rn(vmax[d], 0.01);
    span[d] = __dsub_rn(vmax[d], vmin[d]);
  }
  if (tid == 0) s_base = 0, s_lab = 0, s_odd = 0;
  if (kStats)
    for (int i = tid; i < kVoxelWords; i += kSelThreads) s_vox[i] = 0u;
  __syncthreads();
  int *out = sel + static_cast<size_t>(box) * n;
  int n_lab = 0, n_odd = 0;
  for (int base = 0; base < n; base += kSelThreads) {
    const int k = base + tid;
    bool in = false;
    double p[3];
    if (k < n) {
      const float *r = scene + static_cast<size_t>(k) * ld;
      in = true;
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        p[d] = static_cast<double>(__ldg(r + d));
        in = in && p[d] >= lo[d] && p[d] <= hi[d];
      }
      if (mask != nullptr) mask[static_cast<size_t>(box) * n + k] = in ? 1 : 0;
      if (kStats && in) {
        n_lab += static_cast<int>(__ldg(r + label_col)) > 0 ? 1 : 0;
        bool inner = true;
#pragma unroll
        for (int d = 0; d < 3; ++d) inner = inner && p[d] >= ilo[d] && p[d] <= ihi[d];
        if (inner) {
          // ceil((p - curmin) / (curmax - curmin) * [31, 31, 62]);  key = vx*31*62 + vy*62 + vz   (:237-238)
          const double vx = ceil(__dmul_rn(__ddiv_rn(__dsub_rn(p[0], vmin[0]), span[0]), 31.0));
          const double vy = ceil(__dmul_rn(__ddiv_rn(__dsub_rn(p[1], vmin[1]), span[1]), 31.0));
          const double vz = ceil(__dmul_rn(__ddiv_rn(__dsub_rn(p[2], vmin[2]), span[2]), 62.0));
          const double key = vx * 31.0 * 62.0 + vy * 62.0 + vz;  // small integers: exact in any order
          if (key >= 0.0 && key < 32.0 * kVoxelWords) {
            const unsigned q = static_cast<unsigned>(key);
            atomicOr(&s_vox[q >> 5], 1u << (q & 31));
          } else {
            ++n_odd;  // degenerate box (zero extent): the caller rejects the result
          }
        }
      }
    }
    const unsigned b = __ballot_sync(0xffffffffu, in);
    if (lane == 0) s_warp[warp] = __popc(b);
    __syncthreads();
    if (warp == 0) {
      const int c = s_warp[lane], start = s_base;
      int incl = c;
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const int v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
      }
      s_warp[lane] = start + incl - c;  // first slot of each warp's hits
      __syncwarp();                     // every lane has read s_base
      if (lane == 31) s_base = start + incl;
    }
    __syncthreads();
    if (in) out[s_warp[warp] + __popc(b & ((1u << lane) - 1u))] = k;
    __syncthreads();  // s_warp is rewritten by the next tile
  }
  if (kStats) {
    if (n_lab) atomicAdd(&s_lab, n_lab);
    if (n_odd) atomicAdd(&s_odd, n_odd);
    __syncthreads();
    int bits = 0;
    for (int i = tid; i < kVoxelWords; i += kSelThreads) bits += __popc(s_vox[i]);
    for (int o = 16; o > 0; o >>= 1) bits += __shfl_xor_sync(0xffffffffu, bits, o);
    __syncthreads();
    if (lane == 0) s_warp[warp] = bits;
    __syncthreads();
    if (tid == 0) {
      int t = 0;
      for (int w = 0; w < kSelThreads / 32; ++w) t += s_warp[w];
      stats[4 * box + 0] = s_base;
      stats[4 * box + 1] = s_lab;
      stats[4 * box + 2] = t;
      stats[4 * box + 3] = s_odd;
    }
  } else if (tid == 0) {
    stats[box] = s_base;
  }
}

__device__ __forceinline__ int source_row(const int *__restrict__ sel, const int *__restrict__ choice, int i) {
  const int c = __ldg(choice + i);
  return sel != nullptr ? __ldg(sel + c) : c;
}

// centre[d] = mean over the chunk's rows of the d-th coordinate, in double (fixed summation tree).
__device__ __forceinline__ void chunk_centre(int npoints, int ld, const float *__restrict__ scene,
                                             const int *__restrict__ sel, const int *__restrict__ choice,
                                             double *__restrict__ centre) {
  __shared__ double s_sum[kSelThreads / 32][3];
  double acc[3] = {0.0, 0.0, 0.0};
  for (int i = threadIdx.x; i < npoints; i += kSelThreads) {
    const float *r = scene + static_cast<size_t>(source_row(sel, choice, i)) * ld;
#pragma unroll
    for (int d = 0; d < 3; ++d) acc[d] += static_cast<double>(__ldg(r + d));
  }
#pragma unroll
  for (int d = 0; d < 3; ++d)
    for (int o = 16; o > 0; o >>= 1) acc[d] += __shfl_xor_sync(0xffffffffu, acc[d], o);
  if ((threadIdx.x & 31) == 0)
    for (int d = 0; d < 3; ++d) s_sum[threadIdx.x >> 5][d] = acc[d];
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int w = 0; w < kSelThreads / 32; ++w) t += s_sum[w][threadIdx.x];
    centre[threadIdx.x] = t / static_cast<double>(npoints);
  }
}

__global__ void __launch_bounds__(kSelThreads)
chunk_centre_kernel(int npoints, int ld, const float *__restrict__ scene, const int *__restrict__ sel,
                    const int *__restrict__ choice, double *__restrict__ centre) {
  chunk_centre(npoints, ld, scene, sel, choice, centre);
}

struct Augment {   // _augment's draws (lib/dataset.py:104-137): ops applied in the order t, r, s
  int ops;         // bit 0 translate, bit 1 rotate, bit 2 scale
  double t[3], R[9], f;
};

constexpr int kAsmThreads = 256;
constexpr int kAsmRows = 256;  // rows per CTA

// out[i, :] = [xyz' | (features) | rgb / 255 | normal | (features)] of scene row sel[choice[i]] as float32,
// labels[i] = (int) label column, weights[i] = labelweights[label] * (wmask ? wmask[choice[i]] : 1).
__device__ __forceinline__ void
chunk_assemble(int npoints, int ld, const float *__restrict__ scene, const int *__restrict__ sel,
               const int *__restrict__ choice, const float *__restrict__ feature, int nfeat,
               int feat_first, int use_color, int use_normal, int label_col,
               const float *__restrict__ labelweights, int num_classes,
               const unsigned char *__restrict__ wmask, const double *__restrict__ centre, const Augment &aug,
               int ldo, float *__restrict__ out, long long *__restrict__ labels,
               float *__restrict__ weights) {
  __shared__ int s_src[kAsmRows];
  const int row0 = blockIdx.x * kAsmRows;
  const int rows = min(kAsmRows, npoints - row0);
  const int i = row0 + threadIdx.x;
  if (threadIdx.x < rows) {
    const int src = source_row(sel, choice, i);
    s_src[threadIdx.x] = src;
    const float *r = scene + static_cast<size_t>(src) * ld;
    float *o = out + static_cast<size_t>(i) * ldo;
    double c[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) c[d] = static_cast<double>(__ldg(r + d));
    if (aug.ops != 0) {
      double m[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        m[d] = centre[d];
        c[d] = __dsub_rn(c[d], m[d]);
      }
      if (aug.ops & 1)
        for (int d = 0; d < 3; ++d) c[d] = __dadd_rn(c[d], aug.t[d]);
      if (aug.ops & 2) {
        double q[3];
#pragma unroll
        for (int d = 0; d < 3; ++d)
          q[d] = __dadd_rn(__dadd_rn(__dmul_rn(aug.R[3 * d], c[0]), __dmul_rn(aug.R[3 * d + 1], c[1])),
                           __dmul_rn(aug.R[3 * d + 2], c[2]));
#pragma unroll
        for (int d = 0; d < 3; ++d) c[d] = q[d];
      }
      if (aug.ops & 4)
        for (int d = 0; d < 3; ++d) c[d] = __dmul_rn(c[d], aug.f);
#pragma unroll
      for (int d = 0; d < 3; ++d) c[d] = __dadd_rn(c[d], m[d]);
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) o[d] = static_cast<float>(c[d]);
    int col = 3 + (feat_first ? nfeat : 0);
    if (use_color) {
#pragma unroll
      for (int d = 0; d < 3; ++d)
        o[col + d] = static_cast<float>(__ddiv_rn(static_cast<double>(__ldg(r + 3 + d)), 255.0));
      col += 3;
    }
    if (use_normal) {
#pragma unroll
      for (int d = 0; d < 3; ++d) o[col + d] = __ldg(r + 6 + d);
    }
    const int lab = static_cast<int>(__ldg(r + label_col));  // astype(np.int32): truncation
    labels[i] = lab;
    float w = lab >= 0 && lab < num_classes ? __ldg(labelweights + lab) : 0.f;
    if (wmask != nullptr && !wmask[__ldg(choice + i)]) w = 0.f;  // `mask = mask[choice]` (:334)
    weights[i] = w;
  }
  if (nfeat == 0) return;
  __syncthreads();
  const int col0 = feat_first ? 3 : 3 + 3 * use_color + 3 * use_normal;
  for (int e = threadIdx.x; e < rows * nfeat; e += kAsmThreads) {  // consecutive threads: consecutive columns
    const int rr = e / nfeat, cc = e - rr * nfeat;
    out[static_cast<size_t>(row0 + rr) * ldo + col0 + cc] =
        __ldg(feature + static_cast<size_t>(s_src[rr]) * nfeat + cc);
  }
}

__global__ void __launch_bounds__(kAsmThreads)
chunk_assemble_kernel(int npoints, int ld, const float *__restrict__ scene, const int *__restrict__ sel,
                      const int *__restrict__ choice, const float *__restrict__ feature, int nfeat,
                      int feat_first, int use_color, int use_normal, int label_col,
                      const float *__restrict__ labelweights, int num_classes,
                      const unsigned char *__restrict__ wmask, const double *__restrict__ centre, Augment aug,
                      int ldo, float *__restrict__ out, long long *__restrict__ labels,
                      float *__restrict__ weights) {
  chunk_assemble(npoints, ld, scene, sel, choice, feature, nfeat, feat_first, use_color, use_normal, label_col,
                 labelweights, num_classes, wmask, centre, aug, ldo, out, labels, weights);
}

// ---- a whole batch in two launches: one descriptor per chunk (scenes differ per chunk) ----------------
struct ChunkDesc {              // mirrored by chunker.py's DESC dtype (pn2_chunk_desc_bytes checks the size)
  const float *scene;
  const float *feature;
  const int *sel;
  const int *choice;
  const unsigned char *wmask;
  int ld, nfeat;
  Augment aug;                  // ops (+ 4 bytes padding), t[3], R[9], f
  double centre[3];             // written by chunk_centre_batch_kernel
};

__global__ void __launch_bounds__(kSelThreads)
chunk_centre_batch_kernel(int npoints, ChunkDesc *__restrict__ desc) {
  ChunkDesc *d = desc + blockIdx.x;
  if (d->aug.ops == 0) return;
  chunk_centre(npoints, d->ld, d->scene, d->sel, d->choice, d->centre);
}

__global__ void __launch_bounds__(kAsmThreads)
chunk_assemble_batch_kernel(int npoints, const ChunkDesc *__restrict__ desc, int feat_first, int use_color,
                            int use_normal, int label_col, const float *__restrict__ labelweights,
                            int num_classes, int ldo, float *__restrict__ out, long long *__restrict__ labels,
                            float *__restrict__ weights) {
  const ChunkDesc *d = desc + blockIdx.y;
  const size_t off = static_cast<size_t>(blockIdx.y) * npoints;
  chunk_assemble(npoints, d->ld, d->scene, d->sel, d->choice, d->feature, d->nfeat, feat_first, use_color,
                 use_normal, label_col, labelweights, num_classes, d->wmask, d->centre, d->aug, ldo,
                 out + off * ldo, labels + off, weights + off);
}

}  // namespace pn2

using namespace pn2;

extern "C" int pn2_crop_try(int n, int ld, const float *scene, int label_col, int centre_row, double zmin,
                            double zmax, int *sel, int *stats, void *stream) {
  if (n <= 0 || ld < 3 || label_col < 0 || label_col >= ld || centre_row < 0 || centre_row >= n)
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(scene, sel, stats);
  CropSpec crop{centre_row, zmin, zmax};
  box_select_kernel<true><<<1, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      n, ld, scene, crop, nullptr, nullptr, label_col, sel, stats, nullptr);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_box_select(int n, int ld, const float *scene, int nbox, const double *lo, const double *hi,
                              int *sel, int *count, unsigned char *mask, void *stream) {
  if (n <= 0 || ld < 3 || nbox <= 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(scene, lo, hi, sel, count);
  if ((reinterpret_cast<uintptr_t>(lo) & 7u) != 0 || (reinterpret_cast<uintptr_t>(hi) & 7u) != 0)
    return PN2_ERR_ALIGNMENT;
  CropSpec crop{-1, 0.0, 0.0};
  box_select_kernel<false><<<nbox, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      n, ld, scene, crop, lo, hi, 0, sel, count, mask);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_chunk_centre(int npoints, int ld, const float *scene, const int *sel, const int *choice,
                                double *centre, void *stream) {
  if (npoints <= 0 || ld < 3) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(scene, choice, centre);
  if ((reinterpret_cast<uintptr_t>(centre) & 7u) != 0) return PN2_ERR_ALIGNMENT;
  chunk_centre_kernel<<<1, kSelThreads, 0, static_cast<cudaStream_t>(stream)>>>(npoints, ld, scene, sel,
                                                                                 choice, centre);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_chunk_assemble(int npoints, int ld, const float *scene, const int *sel, const int *choice,
                                  const float *feature, int nfeat, int feat_first, int use_color,
                                  int use_normal, int label_col, const float *labelweights, int num_classes,
                                  const unsigned char *wmask, const double *centre, int aug_ops,
                                  const double *aug_host, int ldo, float *out, long long *labels,
                                  float *weights, void *stream) {
  if (npoints <= 0 || ld < 3 || nfeat < 0 || num_classes <= 0 || label_col < 0 || label_col >= ld)
    return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(scene, choice, labelweights, out, labels, weights);
  if ((use_color || use_normal) && ld < 9) return PN2_ERR_INVALID_ARG;
  if (nfeat > 0 && feature == nullptr) return PN2_ERR_INVALID_ARG;
  if (ldo < 3 + nfeat + 3 * (use_color != 0) + 3 * (use_normal != 0)) return PN2_ERR_INVALID_ARG;
  if ((reinterpret_cast<uintptr_t>(labels) & 7u) != 0) return PN2_ERR_ALIGNMENT;
  Augment aug{};
  aug.ops = aug_ops & 7;
  if (aug.ops != 0) {
    if (aug_host == nullptr || centre == nullptr) return PN2_ERR_INVALID_ARG;
    for (int d = 0; d < 3; ++d) aug.t[d] = aug_host[d];
    for (int d = 0; d < 9; ++d) aug.R[d] = aug_host[3 + d];
    aug.f = aug_host[12];
  }
  const int blocks = (npoints + kAsmRows - 1) / kAsmRows;
  chunk_assemble_kernel<<<blocks, kAsmThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      npoints, ld, scene, sel, choice, feature, nfeat, feat_first != 0, use_color != 0, use_normal != 0,
      label_col, labelweights, num_classes, wmask, centre, aug, ldo, out, labels, weights);
  count_launch();
  PN2_RETURN_LAUNCH_STATUS();
}

extern "C" int pn2_chunk_desc_bytes(void) { return static_cast<int>(sizeof(ChunkDesc)); }

extern "C" int pn2_chunk_batch(int nchunk, int npoints, void *desc, int feat_first, int use_color,
                               int use_normal, int label_col, const float *labelweights, int num_classes,
                               int ldo, float *out, long long *labels, float *weights, void *stream) {
  if (nchunk <= 0 || nchunk > 65535 || npoints <= 0 || num_classes <= 0 || label_col < 0) return PN2_ERR_INVALID_ARG;
  PN2_CHECK_PTRS(desc, labelweights, out, labels, weights);
  if ((reinterpret_cast<uintptr_t>(desc) & 7u) != 0 || (reinterpret_cast<uintptr_t>(labels) & 7u) != 0)
    return PN2_ERR_ALIGNMENT;
  cudaStream_t s = static_cast<cudaStream_t>(stream);
  chunk_centre_batch_kernel<<<nchunk, kSelThreads, 0, s>>>(npoints, static_cast<ChunkDesc *>(desc));
  const dim3 grid((npoints + kAsmRows - 1) / kAsmRows, nchunk);
  chunk_assemble_batch_kernel<<<grid, kAsmThreads, 0, s>>>(
      npoints, static_cast<const ChunkDesc *>(desc), feat_first != 0, use_color != 0, use_normal != 0, label_col,
      labelweights, num_classes, ldo, out, labels, weights);
  count_launch(2);
  PN2_RETURN_LAUNCH_STATUS();
}
