/*
 * flat_ip.c — plain-C restatement of the FAISS CPU flat inner-product search that sits
 * behind the reference's `index.search(q, k)` (src/search.py:115).  TEST INFRASTRUCTURE and
 * CPU baseline only: nothing in the product links or loads this file.
 *
 * The arithmetic is in the third-party dependency `faiss-cpu` (reference requirements.txt:29,
 * UNPINNED, not vendored under /root/reference, not installable here), so this file restates
 * the published algorithm of faiss/utils/distances.cpp + faiss/impl/ResultHandler.h (1.7.x /
 * 1.8.x line) and is anchored on the reference's call sites:
 *
 *   IndexFlat::search -> knn_inner_product(x, y, d, nx, ny, k, D, I)
 *     nx <  distance_compute_blas_threshold (20):
 *         exhaustive_inner_product_seq — `#pragma omp parallel for` over QUERIES; each query
 *         scans all rows with fvec_inner_product and pushes into its own result handler.
 *         (=> a batch-1 search is single-threaded, exactly as in FAISS.)
 *     nx >= 20:
 *         exhaustive_inner_product_blas — query blocks of 4096 x database blocks of 1024,
 *         one sgemm per block pair, then add_results per query row.
 *   result handler: min-heap CMin<float,int64> when k < distance_compute_min_k_reservoir
 *         (100): a row enters only if its score is STRICTLY greater than the heap top, so
 *         among exact ties the earlier row survives; reservoir (capacity 2k, shrink by
 *         partition to k) when k >= 100.  Both end with a best-first reorder.
 *   missing slots: id -1, score -FLT_MAX.
 *
 * Parity pinning: no FAISS here and the reference has no tests => "parity unpinned" at the
 * FAISS boundary.  This file is cross-checked against the numpy oracle
 * (oracle/flat_ip.py) and fp64 brute force in tests/test_oracle.py, and against the shipped
 * 1-row index fixture (KAT-3/4).
 *
 * Build: see oracle/Makefile (gcc -O3 -march=native -fopenmp -shared).
 */
#include <dlfcn.h>
#include <float.h>
#include <math.h>
#include <omp.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define BLAS_THRESHOLD 20
#define MIN_K_RESERVOIR 100
#define BS_X 4096
#define BS_Y 1024

/* ---- fvec_inner_product: 16 independent partial sums (gcc vectorises this to one zmm / two
 * ymm accumulators), tail handled scalar — same shape as FAISS's SIMD kernel. */
static inline float fvec_inner_product(const float* x, const float* y, size_t d) {
  float acc[16] = {0};
  size_t i = 0;
  for (; i + 16 <= d; i += 16)
    for (int j = 0; j < 16; ++j) acc[j] += x[i + j] * y[i + j];
  float s = 0.f;
  for (int j = 0; j < 16; ++j) s += acc[j];
  for (; i < d; ++i) s += x[i] * y[i];
  return s;
}

/* ---- min-heap on (val, id): root = smallest kept score; 1-based sift as in faiss/utils/Heap.h */
static inline int heap_less(float a, int64_t ia, float b, int64_t ib) { /* CMin::cmp2 */
  return (a < b) || (a == b && ia < ib);
}
static void heap_replace_top(size_t k, float* val, int64_t* ids, float v, int64_t id) {
  val--;
  ids--; /* 1-based */
  size_t i = 1, i1, i2;
  for (;;) {
    i1 = i << 1;
    i2 = i1 + 1;
    if (i1 > k) break;
    if (i2 == k + 1 || heap_less(val[i1], ids[i1], val[i2], ids[i2])) {
      if (heap_less(v, id, val[i1], ids[i1])) break;
      val[i] = val[i1];
      ids[i] = ids[i1];
      i = i1;
    } else {
      if (heap_less(v, id, val[i2], ids[i2])) break;
      val[i] = val[i2];
      ids[i] = ids[i2];
      i = i2;
    }
  }
  val[i] = v;
  ids[i] = id;
}
static void heap_init(size_t k, float* val, int64_t* ids) {
  for (size_t i = 0; i < k; ++i) {
    val[i] = -FLT_MAX; /* CMin::neutral() */
    ids[i] = -1;
  }
}

typedef struct {
  float v;
  int64_t id;
} pair_t;
static int pair_desc(const void* a, const void* b) {
  const pair_t *x = (const pair_t*)a, *y = (const pair_t*)b;
  if (x->v > y->v) return -1;
  if (x->v < y->v) return 1;
  /* -1 padding last, otherwise ascending id */
  if (x->id == y->id) return 0;
  if (x->id < 0) return 1;
  if (y->id < 0) return -1;
  return x->id < y->id ? -1 : 1;
}
/* heap_reorder: best first */
static void reorder(size_t k, float* val, int64_t* ids) {
  pair_t* p = (pair_t*)malloc(k * sizeof(pair_t));
  for (size_t i = 0; i < k; ++i) {
    p[i].v = val[i];
    p[i].id = ids[i];
  }
  qsort(p, k, sizeof(pair_t), pair_desc);
  for (size_t i = 0; i < k; ++i) {
    val[i] = p[i].v;
    ids[i] = p[i].id;
  }
  free(p);
}

/* ---- reservoir (k >= 100): buffer of capacity 2k, threshold = k-th best after a shrink */
typedef struct {
  size_t k, cap, n;
  float thr;
  pair_t* buf;
} reservoir_t;
static void res_init(reservoir_t* r, size_t k) {
  r->k = k;
  r->cap = 2 * k;
  r->n = 0;
  r->thr = -FLT_MAX;
  r->buf = (pair_t*)malloc(r->cap * sizeof(pair_t));
}
static void res_shrink(reservoir_t* r) {
  qsort(r->buf, r->n, sizeof(pair_t), pair_desc); /* FAISS uses a partition; same kept set */
  if (r->n > r->k) r->n = r->k;
  r->thr = r->buf[r->n - 1].v;
}
static inline void res_add(reservoir_t* r, float v, int64_t id) {
  if (v > r->thr) { /* strict, as ReservoirTopN::add */
    if (r->n == r->cap) res_shrink(r);
    r->buf[r->n].v = v;
    r->buf[r->n].id = id;
    r->n++;
  }
}
static void res_finish(reservoir_t* r, float* D, int64_t* I) {
  qsort(r->buf, r->n, sizeof(pair_t), pair_desc);
  size_t m = r->n < r->k ? r->n : r->k;
  for (size_t i = 0; i < m; ++i) {
    D[i] = r->buf[i].v;
    I[i] = r->buf[i].id;
  }
  for (size_t i = m; i < r->k; ++i) {
    D[i] = -FLT_MAX;
    I[i] = -1;
  }
  free(r->buf);
}

/* ---- optional BLAS (cblas_sgemm) resolved at run time, e.g. numpy's bundled OpenBLAS ------- */
typedef void (*sgemm32_fn)(int, int, int, int, int, int, float, const float*, int, const float*, int, float, float*,
                           int);
typedef void (*sgemm64_fn)(int, int, int, int64_t, int64_t, int64_t, float, const float*, int64_t, const float*,
                           int64_t, float, float*, int64_t);
static sgemm32_fn g_sgemm32 = 0;
static sgemm64_fn g_sgemm64 = 0;

int oracle_set_blas(const char* lib_path, const char* symbol, int ilp64) {
  void* h = dlopen(lib_path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return 1;
  void* f = dlsym(h, symbol);
  if (!f) return 2;
  g_sgemm32 = 0;
  g_sgemm64 = 0;
  if (ilp64) g_sgemm64 = (sgemm64_fn)f;
  else g_sgemm32 = (sgemm32_fn)f;
  return 0;
}
int oracle_has_blas(void) { return g_sgemm32 || g_sgemm64; }

/* ip[nxb][nyb] = X[nxb][d] . Y[nyb][d]^T */
static void block_ip(const float* X, const float* Y, size_t nxb, size_t nyb, size_t d, float* ip) {
  if (g_sgemm64 || g_sgemm32) {
    /* row-major C = A * B^T : CblasRowMajor=101, NoTrans=111, Trans=112 */
    if (g_sgemm64)
      g_sgemm64(101, 111, 112, (int64_t)nxb, (int64_t)nyb, (int64_t)d, 1.0f, X, (int64_t)d, Y, (int64_t)d, 0.0f, ip,
                (int64_t)nyb);
    else
      g_sgemm32(101, 111, 112, (int)nxb, (int)nyb, (int)d, 1.0f, X, (int)d, Y, (int)d, 0.0f, ip, (int)nyb);
    return;
  }
/* portable fallback: register-blocked 4 queries at a time, OpenMP over database rows */
#pragma omp parallel for schedule(static)
  for (long j = 0; j < (long)nyb; ++j) {
    const float* y = Y + (size_t)j * d;
    size_t i = 0;
    for (; i + 4 <= nxb; i += 4) {
      const float *x0 = X + i * d, *x1 = x0 + d, *x2 = x1 + d, *x3 = x2 + d;
      float a0[16] = {0}, a1[16] = {0}, a2[16] = {0}, a3[16] = {0};
      size_t t = 0;
      for (; t + 16 <= d; t += 16)
        for (int u = 0; u < 16; ++u) {
          const float yv = y[t + u];
          a0[u] += x0[t + u] * yv;
          a1[u] += x1[t + u] * yv;
          a2[u] += x2[t + u] * yv;
          a3[u] += x3[t + u] * yv;
        }
      float s0 = 0, s1 = 0, s2 = 0, s3 = 0;
      for (int u = 0; u < 16; ++u) {
        s0 += a0[u];
        s1 += a1[u];
        s2 += a2[u];
        s3 += a3[u];
      }
      for (; t < d; ++t) {
        s0 += x0[t] * y[t];
        s1 += x1[t] * y[t];
        s2 += x2[t] * y[t];
        s3 += x3[t] * y[t];
      }
      ip[(i + 0) * nyb + j] = s0;
      ip[(i + 1) * nyb + j] = s1;
      ip[(i + 2) * nyb + j] = s2;
      ip[(i + 3) * nyb + j] = s3;
    }
    for (; i < nxb; ++i) ip[i * nyb + j] = fvec_inner_product(X + i * d, y, d);
  }
}

/* ---- knn_inner_product ------------------------------------------------------------------- */
static void search_seq(const float* x, const float* y, size_t d, size_t nx, size_t ny, size_t k, float* D,
                       int64_t* I) {
  int nt = omp_get_max_threads();
  if ((size_t)nt > nx) nt = (int)nx;
#pragma omp parallel for num_threads(nt) schedule(static)
  for (long i = 0; i < (long)nx; ++i) {
    const float* xi = x + (size_t)i * d;
    float* Di = D + (size_t)i * k;
    int64_t* Ii = I + (size_t)i * k;
    if (k < MIN_K_RESERVOIR) {
      heap_init(k, Di, Ii);
      for (size_t j = 0; j < ny; ++j) {
        const float ip = fvec_inner_product(xi, y + j * d, d);
        if (ip > Di[0]) heap_replace_top(k, Di, Ii, ip, (int64_t)j);
      }
      reorder(k, Di, Ii);
    } else {
      reservoir_t r;
      res_init(&r, k);
      for (size_t j = 0; j < ny; ++j) res_add(&r, fvec_inner_product(xi, y + j * d, d), (int64_t)j);
      res_finish(&r, Di, Ii);
    }
  }
}

static void search_blas(const float* x, const float* y, size_t d, size_t nx, size_t ny, size_t k, float* D,
                        int64_t* I) {
  float* ip = (float*)malloc((size_t)BS_X * BS_Y * sizeof(float));
  for (size_t i0 = 0; i0 < nx; i0 += BS_X) {
    const size_t i1 = i0 + BS_X < nx ? i0 + BS_X : nx;
    const int use_heap = k < MIN_K_RESERVOIR;
    reservoir_t* res = 0;
    if (use_heap) {
      for (size_t i = i0; i < i1; ++i) heap_init(k, D + i * k, I + i * k);
    } else {
      res = (reservoir_t*)malloc((i1 - i0) * sizeof(reservoir_t));
      for (size_t i = i0; i < i1; ++i) res_init(&res[i - i0], k);
    }
    for (size_t j0 = 0; j0 < ny; j0 += BS_Y) {
      const size_t j1 = j0 + BS_Y < ny ? j0 + BS_Y : ny;
      block_ip(x + i0 * d, y + j0 * d, i1 - i0, j1 - j0, d, ip);
/* add_results: one thread per query row of the block */
#pragma omp parallel for schedule(static)
      for (long i = (long)i0; i < (long)i1; ++i) {
        const float* row = ip + (size_t)(i - (long)i0) * (j1 - j0);
        if (use_heap) {
          float* Di = D + (size_t)i * k;
          int64_t* Ii = I + (size_t)i * k;
          for (size_t j = j0; j < j1; ++j) {
            const float v = row[j - j0];
            if (v > Di[0]) heap_replace_top(k, Di, Ii, v, (int64_t)j);
          }
        } else {
          reservoir_t* r = &res[i - (long)i0];
          for (size_t j = j0; j < j1; ++j) res_add(r, row[j - j0], (int64_t)j);
        }
      }
    }
#pragma omp parallel for schedule(static)
    for (long i = (long)i0; i < (long)i1; ++i) {
      if (use_heap) reorder(k, D + (size_t)i * k, I + (size_t)i * k);
      else res_finish(&res[i - (long)i0], D + (size_t)i * k, I + (size_t)i * k);
    }
    free(res);
  }
  free(ip);
}

/* IndexFlatIP::search.  Returns 0; 1 on bad arguments (FAISS throws for k <= 0). */
int oracle_flat_ip_search(const float* xb, int64_t n, int64_t d, const float* xq, int64_t nq, int64_t k, float* D,
                          int64_t* I) {
  if (k <= 0 || d <= 0 || n < 0 || nq < 0) return 1;
  if (nq == 0) return 0;
  if (nq < BLAS_THRESHOLD) search_seq(xq, xb, (size_t)d, (size_t)nq, (size_t)n, (size_t)k, D, I);
  else search_blas(xq, xb, (size_t)d, (size_t)nq, (size_t)n, (size_t)k, D, I);
  return 0;
}

/* "Optimistic CPU" variant for batch-1 (NOT what FAISS does): the single query's scan is split
 * over all threads by rows and the per-thread heaps are merged.  Reported next to the faithful
 * number so the GPU/CPU ratio is not flattered by FAISS's one-thread-per-query rule. */
int oracle_flat_ip_search_rowpar(const float* xb, int64_t n, int64_t d, const float* xq, int64_t nq, int64_t k,
                                 float* D, int64_t* I) {
  if (k <= 0 || d <= 0 || n < 0 || nq < 0) return 1;
  const int nt = omp_get_max_threads();
  float* Dt = (float*)malloc((size_t)nt * k * sizeof(float));
  int64_t* It = (int64_t*)malloc((size_t)nt * k * sizeof(int64_t));
  for (int64_t q = 0; q < nq; ++q) {
    const float* xi = xq + q * d;
#pragma omp parallel num_threads(nt)
    {
      const int t = omp_get_thread_num();
      float* Di = Dt + (size_t)t * k;
      int64_t* Ii = It + (size_t)t * k;
      heap_init((size_t)k, Di, Ii);
      const int64_t j0 = n * t / nt, j1 = n * (t + 1) / nt;
      for (int64_t j = j0; j < j1; ++j) {
        const float ip = fvec_inner_product(xi, xb + j * d, (size_t)d);
        if (ip > Di[0]) heap_replace_top((size_t)k, Di, Ii, ip, j);
      }
    }
    pair_t* p = (pair_t*)malloc((size_t)nt * k * sizeof(pair_t));
    for (size_t i = 0; i < (size_t)nt * k; ++i) {
      p[i].v = Dt[i];
      p[i].id = It[i];
    }
    qsort(p, (size_t)nt * k, sizeof(pair_t), pair_desc);
    for (int64_t i = 0; i < k; ++i) {
      D[q * k + i] = p[i].v;
      I[q * k + i] = p[i].id;
    }
    free(p);
  }
  free(Dt);
  free(It);
  return 0;
}

int oracle_num_threads(void) { return omp_get_max_threads(); }
void oracle_set_num_threads(int n) { omp_set_num_threads(n); }
