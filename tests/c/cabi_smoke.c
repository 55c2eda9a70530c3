/* The C ABI from plain C (include/sgic.h): no Python, no torch, no C++ — what a C / cgo / JNI host of the reference's
 * search path would link.  Builds an index of pseudo-random unit vectors, searches it (src/search.py:115 semantics: k
 * best inner products, descending, ties by ascending row, -1 / -FLT_MAX padding), checks the answer against a brute
 * force in this file on the fp16-rounded values' neighbourhood (1e-3, the north-star tolerance), repeats the search on a
 * two-shard multi-GPU handle (both shards on device 0 when only one GPU is there) and on an index re-read from an IxFI
 * file.  Exit code 0 = all good; 77 = no usable GPU (the caller skips).
 *
 *   gcc -std=c99 -O2 -I include tests/c/cabi_smoke.c -L <pkg> -lsgic -Wl,-rpath,<pkg> -lm -o cabi_smoke
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sgic.h"

#define N 20000
#define D 64
#define NQ 5
#define K 7

static uint64_t rng_state = 88172645463325252ull;
static double rnd(void) { /* xorshift64*, uniform in (0,1) */
  rng_state ^= rng_state >> 12;
  rng_state ^= rng_state << 25;
  rng_state ^= rng_state >> 27;
  return (double)((rng_state * 2685821657736338717ull) >> 11) / 9007199254740992.0 + 1e-12;
}
static float gauss(void) { return (float)(sqrt(-2.0 * log(rnd())) * cos(6.283185307179586 * rnd())); }

static void unit_rows(float* x, int n) {
  for (int i = 0; i < n; ++i) {
    double s = 0;
    for (int j = 0; j < D; ++j) {
      x[i * D + j] = gauss();
      s += (double)x[i * D + j] * x[i * D + j];
    }
    for (int j = 0; j < D; ++j) x[i * D + j] = (float)(x[i * D + j] / sqrt(s));
  }
}

static int check(const char* what, const float* xb, const float* xq, const float* Dg, const int64_t* Ig) {
  for (int q = 0; q < NQ; ++q) {
    /* brute force: the K best rows by fp64 dot product */
    double best[K];
    for (int r = 0; r < K; ++r) best[r] = -1e30;
    for (int i = 0; i < N; ++i) {
      double s = 0;
      for (int j = 0; j < D; ++j) s += (double)xb[i * D + j] * xq[q * D + j];
      if (s > best[K - 1]) {
        int r = K - 1;
        while (r > 0 && best[r - 1] < s) {
          best[r] = best[r - 1];
          --r;
        }
        best[r] = s;
      }
    }
    for (int r = 0; r < K; ++r) {
      const int64_t id = Ig[q * K + r];
      if (id < 0 || id >= N) return fprintf(stderr, "%s: q%d r%d id %lld out of range\n", what, q, r, (long long)id), 1;
      double s = 0;
      for (int j = 0; j < D; ++j) s += (double)xb[id * D + j] * xq[q * D + j];
      if (fabs(s - Dg[q * K + r]) > 1e-3) return fprintf(stderr, "%s: q%d r%d score %f vs %f\n", what, q, r, Dg[q * K + r], s), 1;
      if (fabs(best[r] - Dg[q * K + r]) > 2e-3) return fprintf(stderr, "%s: q%d rank %d is %f, best is %f\n", what, q, r, Dg[q * K + r], best[r]), 1;
      if (r > 0 && Dg[q * K + r] > Dg[q * K + r - 1]) return fprintf(stderr, "%s: q%d not sorted\n", what, q), 1;
    }
  }
  return 0;
}

#define OK(call)                                                        \
  do {                                                                  \
    if ((call) != 0) {                                                  \
      fprintf(stderr, "%s failed: %s\n", #call, sgic_last_error());     \
      return 1;                                                         \
    }                                                                   \
  } while (0)

int main(int argc, char** argv) {
  const char* tmp = argc > 1 ? argv[1] : "/tmp/cabi_smoke.index";
  float* xb = (float*)malloc(sizeof(float) * N * D);
  float* xq = (float*)malloc(sizeof(float) * NQ * D);
  unit_rows(xb, N);
  unit_rows(xq, NQ);
  memcpy(xq, xb + 4242 * D, sizeof(float) * D); /* query 0 is a database row: it must come back first */
  float Dg[NQ * K];
  int64_t Ig[NQ * K];

  sgic_index* h = NULL;
  if (sgic_index_create(D, SGIC_F16, 0, 0, SGIC_RETAIN_F32, &h) != 0) {
    fprintf(stderr, "no usable GPU: %s\n", sgic_last_error());
    return 77;
  }
  OK(sgic_index_add_f32(h, N / 2, xb));
  OK(sgic_index_add_f32(h, N - N / 2, xb + (size_t)(N / 2) * D));
  if (sgic_index_ntotal(h) != N || sgic_index_d(h) != D) return fprintf(stderr, "ntotal / d wrong\n"), 1;
  OK(sgic_index_search(h, NQ, xq, K, Dg, Ig));
  if (check("one GPU", xb, xq, Dg, Ig)) return 1;
  if (Ig[0] != 4242) return fprintf(stderr, "self query came back as %lld\n", (long long)Ig[0]), 1;
  /* one query at a time (the reference's do_search) gives the same rows */
  for (int q = 0; q < NQ; ++q) {
    float d1[K];
    int64_t i1[K];
    OK(sgic_index_search(h, 1, xq + q * D, K, d1, i1));
    if (memcmp(i1, Ig + q * K, sizeof(i1)) != 0) return fprintf(stderr, "batch 1 and batch %d disagree on q%d\n", NQ, q), 1;
  }
  /* k beyond the rows: padding */
  {
    sgic_index* tiny = NULL;
    float d3[3];
    int64_t i3[3];
    OK(sgic_index_create(D, SGIC_F16, 0, 0, 0, &tiny));
    OK(sgic_index_add_f32(tiny, 2, xb));
    OK(sgic_index_search(tiny, 1, xq, 3, d3, i3));
    if (i3[2] != -1 || d3[2] != -FLT_MAX || i3[0] < 0 || i3[1] < 0) return fprintf(stderr, "padding wrong\n"), 1;
    OK(sgic_index_destroy(tiny));
  }
  /* IxFI round trip (faiss.write_index / read_index) */
  OK(sgic_index_write(h, tmp));
  sgic_index* back = NULL;
  OK(sgic_index_read(tmp, SGIC_F16, 0, 0, &back));
  float Db[NQ * K];
  int64_t Ib[NQ * K];
  OK(sgic_index_search(back, NQ, xq, K, Db, Ib));
  if (memcmp(Ib, Ig, sizeof(Ig)) != 0 || memcmp(Db, Dg, sizeof(Dg)) != 0) return fprintf(stderr, "re-read index answers differently\n"), 1;
  OK(sgic_index_destroy(back));
  /* the same rows behind a two-shard handle: identical answer */
  {
    int devs[2] = {0, 0};
    sgic_index* many = NULL;
    float Dm[NQ * K];
    int64_t Im[NQ * K];
    OK(sgic_index_create_sharded(D, SGIC_F16, 2, devs, 0, 0, &many));
    OK(sgic_index_add_f32(many, N, xb));
    if (sgic_index_n_shards(many) != 2 || sgic_index_ntotal(many) != N) return fprintf(stderr, "sharded handle wrong\n"), 1;
    OK(sgic_index_search(many, NQ, xq, K, Dm, Im));
    if (memcmp(Im, Ig, sizeof(Ig)) != 0 || memcmp(Dm, Dg, sizeof(Dg)) != 0) return fprintf(stderr, "two shards answer differently\n"), 1;
    OK(sgic_index_destroy(many));
  }
  OK(sgic_index_destroy(h));
  remove(tmp);
  free(xb);
  free(xq);
  printf("cabi_smoke ok: %d rows, %d queries, k=%d (C99, libsgic %d)\n", N, NQ, K, sgic_version());
  return 0;
}
