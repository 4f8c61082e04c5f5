/* knn_demo.c — libfenix_knn.so from plain C: the drop-in boundary needs no Python and no torch.
 *
 *   gcc -O2 -I include examples/knn_demo.c -L fenix_b200 -lfenix_knn -Wl,-rpath,$PWD/fenix_b200 -lm -o knn_demo && ./knn_demo
 *
 * Uploads a small random shard (fx_corpus_create / append / finalize), answers a single query (the latency path) and a
 * batch (the tensor-core path) with every metric, and checks the neighbours against a double-precision brute force on the
 * host that follows the reference's formulas (src/fenix/io/coder/coder.py:38-50). Exit code 0 = all equal. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "fenix_knn.h"

#define CHECK(call)                                                            \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ != FX_OK) { fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, fx_last_error()); return 2; } \
  } while (0)

static double ref_distance(int metric, const float* q, const float* x, int d) {
  double qq = 0, xx = 0, qx = 0;
  for (int i = 0; i < d; ++i) { qq += (double)q[i] * q[i]; xx += (double)x[i] * x[i]; qx += (double)q[i] * x[i]; }
  if (metric == FX_METRIC_L2) { double d2 = qq - 2 * qx + xx; return sqrt(d2 > 0 ? d2 : 0); }
  if (metric == FX_METRIC_COSINE) {
    double nq = sqrt(qq), nx = sqrt(xx);
    return 0.5 - 0.5 * qx / ((nq > 1e-12 ? nq : 1e-12) * (nx > 1e-12 ? nx : 1e-12));
  }
  return -qx;
}

static float frand(uint64_t* s) {   /* xorshift, uniform in (-1, 1) */
  *s ^= *s << 13; *s ^= *s >> 7; *s ^= *s << 17;
  return (float)((double)(*s >> 11) / 9007199254740992.0 * 2.0 - 1.0);
}

int main(void) {
  const int64_t n = 20000; const int d = 96, k = 5, nq = 200;
  uint64_t seed = 88172645463325252ull;
  float* x = malloc(sizeof(float) * n * d); float* q = malloc(sizeof(float) * nq * d);
  int64_t* rows = malloc(sizeof(int64_t) * nq * k); float* dist = malloc(sizeof(float) * nq * k);
  if (!x || !q || !rows || !dist) return 3;
  for (int64_t i = 0; i < n * d; ++i) x[i] = frand(&seed);
  for (int64_t i = 0; i < (int64_t)nq * d; ++i) q[i] = frand(&seed);

  fx_ctx* ctx = NULL; fx_corpus* c = NULL;
  CHECK(fx_init(0, &ctx));
  CHECK(fx_corpus_create(ctx, n, d, FX_DTYPE_F32, 0, &c));
  CHECK(fx_corpus_append(c, x, n / 2));                       /* two record batches */
  CHECK(fx_corpus_append(c, x + (n / 2) * d, n - n / 2));
  CHECK(fx_corpus_finalize(c));

  int bad = 0;
  for (int metric = 0; metric < 3; ++metric) {
    for (int pass = 0; pass < 2; ++pass) {                    /* one query (direct scan), then the batch (tcgen05 filter) */
      const int n_q = pass == 0 ? 1 : nq;
      CHECK(fx_search(c, q, n_q, metric, k, FX_PREC_FP32, NULL, rows, dist));
      fx_stats st; CHECK(fx_get_stats(c, &st));
      for (int qi = 0; qi < n_q; qi += (n_q > 1 ? 37 : 1)) {
        /* brute force: the k smallest (distance, row) */
        int64_t best_r[5]; double best_d[5];
        for (int j = 0; j < k; ++j) { best_r[j] = -1; best_d[j] = INFINITY; }
        for (int64_t r = 0; r < n; ++r) {
          double dd = (double)(float)ref_distance(metric, q + (int64_t)qi * d, x + r * d, d);
          int j = k - 1;
          if (dd >= best_d[j]) continue;
          while (j > 0 && dd < best_d[j - 1]) { best_d[j] = best_d[j - 1]; best_r[j] = best_r[j - 1]; --j; }
          best_d[j] = dd; best_r[j] = r;
        }
        for (int j = 0; j < k; ++j) {
          if (rows[qi * k + j] != best_r[j] || fabs(dist[qi * k + j] - best_d[j]) > 1e-5 * fmax(1.0, fabs(best_d[j]))) {
            if (bad < 5) fprintf(stderr, "metric %d query %d rank %d: got (%lld, %.7g) want (%lld, %.7g)\n", metric, qi, j,
                                 (long long)rows[qi * k + j], dist[qi * k + j], (long long)best_r[j], best_d[j]);
            ++bad;
          }
        }
      }
      printf("metric %d, %3d quer%s: path %d, %.1f us on the device\n", metric, n_q, n_q == 1 ? "y" : "ies", st.last_path,
             st.last_search_ms * 1e3);
    }
  }
  CHECK(fx_corpus_destroy(c));
  CHECK(fx_shutdown(ctx));
  free(x); free(q); free(rows); free(dist);
  printf(bad ? "KNN DEMO FAILED (%d mismatches)\n" : "KNN DEMO OK\n", bad);
  return bad ? 1 : 0;
}
