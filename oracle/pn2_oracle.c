/*
 * pn2_oracle.c -- CPU restatement of the PointNet++ SA/FP custom ops of
 * ero1311/Pointnet2.ScanNet.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * legs may load this library.  It is the checker, never the product: the
 * product path (pointnet2/scannet_b200) is CUDA-only and fails loudly without
 * its extension.
 *
 * Every function restates, in plain scalar C, the arithmetic of one reference
 * kernel under /root/reference/pointnet2/_ext_src/src (cited per function) with
 * FMA contraction disabled (BASELINE.json north_star):
 * build with  gcc -O2 -ffp-contract=off  (see oracle/Makefile).  Where the
 * reference's answer depends on its thread layout (the FPS arg-max tie-break)
 * the thread layout itself is emulated: one loop iteration per CUDA thread,
 * one pass per tree level.
 *
 * Parity pin: this file is checked (tests/test_oracle_golden.py) against
 * outputs of the reference's own kernels, compiled unmodified from
 * /root/reference with -fmad=false (oracle/build_ref.py -> oracle/_ref/) and
 * run on a B200; those outputs are committed under tests/golden/.
 *
 * Layouts follow the reference: all float32 / int32, contiguous, batch-major.
 */
#include <float.h>
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PN2O_API __attribute__((visibility("default")))

/* every fp32 operation must round to fp32: no x87 extended precision, and the
 * Makefile passes -ffp-contract=off so that no a*b+c is fused */
_Static_assert(FLT_EVAL_METHOD == 0, "float expressions must evaluate in fp32");
#if defined(__FAST_MATH__)
#error "the oracle must not be built with -ffast-math"
#endif

/* ------------------------------------------------------------------------ */
/* batch-parallel helper (one cloud = one work item, like blockIdx.x)        */
/* ------------------------------------------------------------------------ */
static int g_threads = 1;

PN2O_API void pn2o_set_threads(int n) { g_threads = n < 1 ? 1 : n; }
PN2O_API int pn2o_get_threads(void) { return g_threads; }

typedef void (*cloud_fn)(int b, void *args);
typedef struct {
  cloud_fn fn;
  void *args;
  int begin, end;
} span_t;

static void *span_main(void *p) {
  span_t *s = (span_t *)p;
  for (int b = s->begin; b < s->end; ++b) s->fn(b, s->args);
  return NULL;
}

static void for_each_cloud(int nb, cloud_fn fn, void *args) {
  int nt = g_threads < nb ? g_threads : nb;
  if (nt <= 1) {
    for (int b = 0; b < nb; ++b) fn(b, args);
    return;
  }
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * nt);
  span_t *sp = (span_t *)malloc(sizeof(span_t) * nt);
  for (int t = 0; t < nt; ++t) {
    sp[t].fn = fn;
    sp[t].args = args;
    sp[t].begin = (int)((long long)nb * t / nt);
    sp[t].end = (int)((long long)nb * (t + 1) / nt);
    pthread_create(&th[t], NULL, span_main, &sp[t]);
  }
  for (int t = 0; t < nt; ++t) pthread_join(th[t], NULL);
  free(th);
  free(sp);
}

/* ------------------------------------------------------------------------ */
/* launch geometry: _ext_src/include/cuda_utils.h:13-19                      */
/* (TOTAL_THREADS = 512; block = largest power of two <= work, in [1,512];   */
/*  the exponent is the C-truncated quotient of two natural logs in double)  */
/* ------------------------------------------------------------------------ */
PN2O_API int pn2o_block_threads(int work) {
  int e = (int)(log((double)work) / log(2.0));
  int t = 1 << e;
  if (t > 512) t = 512;
  if (t < 1) t = 1;
  return t;
}

/* squared distance, source order, every operation rounded to fp32           */
/* (sampling_gpu.cu:103-104, ball_query_gpu.cu:31-32, interpolate_gpu.cu:33) */
static inline float sqdist(float ax, float ay, float az, float bx, float by,
                           float bz) {
  const float dx = ax - bx, dy = ay - by, dz = az - bz;
  const float xx = dx * dx, yy = dy * dy, zz = dz * dz;
  const float s = xx + yy;
  return s + zz;
}

/* ------------------------------------------------------------------------ */
/* furthest point sampling                                                   */
/*   kernel   sampling_gpu.cu:69-173  (tree step  :59-65)                    */
/*   launcher sampling_gpu.cu:175-229 (block = pn2o_block_threads(n))        */
/*   host     sampling.cpp:66-87      (idx zero-filled, temp filled 1e10)    */
/* ------------------------------------------------------------------------ */
typedef struct {
  int n, m;
  const float *xyz;
  float *temp;
  int *idx;
} fps_args;

static void fps_cloud(int b, void *pa) {
  fps_args *a = (fps_args *)pa;
  const int n = a->n, m = a->m;
  const float *p = a->xyz + (size_t)b * n * 3;
  float *temp = a->temp + (size_t)b * n;
  int *out = a->idx + (size_t)b * m;
  if (m <= 0) return; /* :72 */
  const int bs = pn2o_block_threads(n);
  float *slot_v = (float *)malloc(sizeof(float) * bs);
  int *slot_i = (int *)malloc(sizeof(int) * bs);

  int cur = 0;
  out[0] = 0; /* :85-86 -- emitted even when point 0 is a near-origin point */
  for (int j = 1; j < m; ++j) {
    const float cx = p[cur * 3 + 0], cy = p[cur * 3 + 1], cz = p[cur * 3 + 2];
    /* per-thread strided scan, :90-110 */
    for (int tid = 0; tid < bs; ++tid) {
      float best = -1.0f;
      int besti = 0;
      for (int k = tid; k < n; k += bs) {
        const float x = p[k * 3 + 0], y = p[k * 3 + 1], z = p[k * 3 + 2];
        const float xx = x * x, yy = y * y, zz = z * z;
        const float s = xx + yy;
        const float mag = s + zz;
        if ((double)mag <= 1e-3) continue; /* :101, compare is in double */
        const float d = sqdist(x, y, z, cx, cy, cz);
        const float d2 = d < temp[k] ? d : temp[k]; /* min(d, temp[k]) :106 */
        temp[k] = d2;
        if (d2 > best) { /* strict: first (lowest k) of a thread's ties */
          besti = k;
          best = d2;
        }
      }
      slot_v[tid] = best;
      slot_i[tid] = besti;
    }
    /* shared-memory tree, :115-167: level s merges slot t+s into slot t;    */
    /* the upper slot wins only on strictly greater value (:59-65)           */
    for (int s = bs >> 1; s >= 1; s >>= 1) {
      for (int t = 0; t < s; ++t) {
        const float v1 = slot_v[t], v2 = slot_v[t + s];
        const int i1 = slot_i[t], i2 = slot_i[t + s];
        slot_v[t] = v1 > v2 ? v1 : v2;
        slot_i[t] = v2 > v1 ? i2 : i1;
      }
    }
    cur = slot_i[0];
    out[j] = cur; /* :169-170 */
  }
  free(slot_v);
  free(slot_i);
}

/* temp may be NULL: then a scratch filled with 1e10f is used, as the host   */
/* binding does (sampling.cpp:74-76). A caller-supplied temp must be so      */
/* filled by the caller and is updated in place like the reference's.        */
PN2O_API int pn2o_furthest_point_sampling(int b, int n, int m,
                                          const float *dataset, float *temp,
                                          int *idxs) {
  float *own = NULL;
  if (!temp) {
    own = (float *)malloc(sizeof(float) * (size_t)b * n);
    for (size_t i = 0; i < (size_t)b * n; ++i) own[i] = 1e10f;
    temp = own;
  }
  memset(idxs, 0, sizeof(int) * (size_t)b * m);
  fps_args a = {n, m, dataset, temp, idxs};
  for_each_cloud(b, fps_cloud, &a);
  free(own);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* gather: sampling_gpu.cu:8-20 (fwd) / :34-47 (grad, atomicAdd)            */
/* host zero-fills both outputs: sampling.cpp:28-30, :51-53                  */
/* ------------------------------------------------------------------------ */
PN2O_API int pn2o_gather_points(int b, int c, int n, int npoints,
                                const float *points, const int *idx,
                                float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l)
      for (int j = 0; j < npoints; ++j)
        out[((size_t)i * c + l) * npoints + j] =
            points[((size_t)i * c + l) * n + idx[(size_t)i * npoints + j]];
  return 0;
}

/* The reference accumulates with fp32 atomicAdd in a nondeterministic order;
 * the oracle accumulates each destination in double and rounds once, i.e. it
 * returns the value every fp32 ordering is a perturbation of. */
PN2O_API int pn2o_gather_points_grad(int b, int c, int n, int npoints,
                                     const float *grad_out, const int *idx,
                                     float *grad_points) {
  double *acc = (double *)malloc(sizeof(double) * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l) {
      memset(acc, 0, sizeof(double) * n);
      for (int j = 0; j < npoints; ++j)
        acc[idx[(size_t)i * npoints + j]] +=
            grad_out[((size_t)i * c + l) * npoints + j];
      for (int k = 0; k < n; ++k)
        grad_points[((size_t)i * c + l) * n + k] = (float)acc[k];
    }
  free(acc);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* ball query: ball_query_gpu.cu:9-44; host zero-fill ball_query.cpp:19-21   */
/* ------------------------------------------------------------------------ */
typedef struct {
  int n, m, nsample;
  float radius;
  const float *new_xyz, *xyz;
  int *idx;
} bq_args;

static void bq_cloud(int b, void *pa) {
  bq_args *a = (bq_args *)pa;
  const int n = a->n, m = a->m, ns = a->nsample;
  const float *pts = a->xyz + (size_t)b * n * 3;
  const float *ctr = a->new_xyz + (size_t)b * m * 3;
  int *idx = a->idx + (size_t)b * m * ns;
  const float r2 = a->radius * a->radius; /* fp32 product, :22 */
  for (int j = 0; j < m; ++j) {
    const float qx = ctr[j * 3 + 0], qy = ctr[j * 3 + 1], qz = ctr[j * 3 + 2];
    int cnt = 0;
    for (int k = 0; k < n && cnt < ns; ++k) { /* early exit, :27 */
      const float d2 =
          sqdist(qx, qy, qz, pts[k * 3 + 0], pts[k * 3 + 1], pts[k * 3 + 2]);
      if (d2 < r2) {
        if (cnt == 0) /* first hit pre-fills every slot, :34-38 */
          for (int l = 0; l < ns; ++l) idx[j * ns + l] = k;
        idx[j * ns + cnt] = k;
        ++cnt;
      }
    }
  }
}

PN2O_API int pn2o_query_ball_point(int b, int n, int m, float radius,
                                   int nsample, const float *new_xyz,
                                   const float *xyz, int *idx) {
  memset(idx, 0, sizeof(int) * (size_t)b * m * nsample);
  bq_args a = {n, m, nsample, radius, new_xyz, xyz, idx};
  for_each_cloud(b, bq_cloud, &a);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* grouping: group_points_gpu.cu:8-28 (fwd) / :43-64 (grad, atomicAdd)       */
/* ------------------------------------------------------------------------ */
PN2O_API int pn2o_group_points(int b, int c, int n, int npoints, int nsample,
                               const float *points, const int *idx,
                               float *out) {
  const size_t e = (size_t)npoints * nsample;
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l) {
      const float *row = points + ((size_t)i * c + l) * n;
      float *dst = out + ((size_t)i * c + l) * e;
      const int *ix = idx + (size_t)i * e;
      for (size_t t = 0; t < e; ++t) dst[t] = row[ix[t]];
    }
  return 0;
}

PN2O_API int pn2o_group_points_grad(int b, int c, int n, int npoints,
                                    int nsample, const float *grad_out,
                                    const int *idx, float *grad_points) {
  const size_t e = (size_t)npoints * nsample;
  double *acc = (double *)malloc(sizeof(double) * n);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l) {
      memset(acc, 0, sizeof(double) * n);
      const float *g = grad_out + ((size_t)i * c + l) * e;
      const int *ix = idx + (size_t)i * e;
      for (size_t t = 0; t < e; ++t) acc[ix[t]] += g[t];
      float *dst = grad_points + ((size_t)i * c + l) * n;
      for (int k = 0; k < n; ++k) dst[k] = (float)acc[k];
    }
  free(acc);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* three_nn: interpolate_gpu.cu:9-59. Running bests are double 1e40 (:27),   */
/* candidates float; outputs are the SQUARED distances narrowed to float     */
/* ((float)1e40 = +inf when fewer than 3 known points) and indices (init 0). */
/* ------------------------------------------------------------------------ */
typedef struct {
  int n, m;
  const float *unknown, *known;
  float *dist2;
  int *idx;
} nn_args;

static void nn_cloud(int b, void *pa) {
  nn_args *a = (nn_args *)pa;
  const int n = a->n, m = a->m;
  const float *u = a->unknown + (size_t)b * n * 3;
  const float *kn = a->known + (size_t)b * m * 3;
  float *d2o = a->dist2 + (size_t)b * n * 3;
  int *io = a->idx + (size_t)b * n * 3;
  for (int j = 0; j < n; ++j) {
    const float ux = u[j * 3 + 0], uy = u[j * 3 + 1], uz = u[j * 3 + 2];
    double b1 = 1e40, b2 = 1e40, b3 = 1e40;
    int i1 = 0, i2 = 0, i3 = 0;
    for (int k = 0; k < m; ++k) {
      const float d =
          sqdist(ux, uy, uz, kn[k * 3 + 0], kn[k * 3 + 1], kn[k * 3 + 2]);
      if (d < b1) {
        b3 = b2, i3 = i2;
        b2 = b1, i2 = i1;
        b1 = d, i1 = k;
      } else if (d < b2) {
        b3 = b2, i3 = i2;
        b2 = d, i2 = k;
      } else if (d < b3) {
        b3 = d, i3 = k;
      }
    }
    d2o[j * 3 + 0] = (float)b1, d2o[j * 3 + 1] = (float)b2,
                d2o[j * 3 + 2] = (float)b3;
    io[j * 3 + 0] = i1, io[j * 3 + 1] = i2, io[j * 3 + 2] = i3;
  }
}

PN2O_API int pn2o_three_nn(int b, int n, int m, const float *unknown,
                           const float *known, float *dist2, int *idx) {
  nn_args a = {n, m, unknown, known, dist2, idx};
  for_each_cloud(b, nn_cloud, &a);
  return 0;
}

/* ------------------------------------------------------------------------ */
/* three_interpolate: interpolate_gpu.cu:72-101 (fwd), :116-143 (grad)       */
/* fwd: ((p1*w1) + (p2*w2)) + (p3*w3), each op rounded to fp32               */
/* ------------------------------------------------------------------------ */
PN2O_API int pn2o_three_interpolate(int b, int c, int m, int n,
                                    const float *points, const int *idx,
                                    const float *weight, float *out) {
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l) {
      const float *row = points + ((size_t)i * c + l) * m;
      float *dst = out + ((size_t)i * c + l) * n;
      const int *ix = idx + (size_t)i * n * 3;
      const float *w = weight + (size_t)i * n * 3;
      for (int j = 0; j < n; ++j) {
        const float t1 = row[ix[j * 3 + 0]] * w[j * 3 + 0];
        const float t2 = row[ix[j * 3 + 1]] * w[j * 3 + 1];
        const float t3 = row[ix[j * 3 + 2]] * w[j * 3 + 2];
        const float s = t1 + t2;
        dst[j] = s + t3;
      }
    }
  return 0;
}

PN2O_API int pn2o_three_interpolate_grad(int b, int c, int n, int m,
                                         const float *grad_out, const int *idx,
                                         const float *weight,
                                         float *grad_points) {
  double *acc = (double *)malloc(sizeof(double) * m);
  for (int i = 0; i < b; ++i)
    for (int l = 0; l < c; ++l) {
      memset(acc, 0, sizeof(double) * m);
      const float *g = grad_out + ((size_t)i * c + l) * n;
      const int *ix = idx + (size_t)i * n * 3;
      const float *w = weight + (size_t)i * n * 3;
      for (int j = 0; j < n; ++j)
        for (int t = 0; t < 3; ++t) {
          /* each addend is the fp32 product the kernel forms (:139-141) */
          const float prod = g[j] * w[j * 3 + t];
          acc[ix[j * 3 + t]] += prod;
        }
      float *dst = grad_points + ((size_t)i * c + l) * m;
      for (int k = 0; k < m; ++k) dst[k] = (float)acc[k];
    }
  free(acc);
  return 0;
}
