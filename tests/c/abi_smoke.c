/* abi_smoke.c -- the drop-in boundary exercised from plain C (no Python, no torch): proves include/proxb200.h is valid C and that a
 * host with nothing but a C FFI (the reference's host language is Julia: `ccall`) can own device vectors, run the fused step and run
 * a whole solve.  Compiled with gcc and run by tests/test_gpu_abi_c.py, which passes the reference's lasso_small fixture as a raw
 * little-endian file:  int64 m, int64 n, double lam, double A[m*n] (column-major), double b[m].
 *
 *   usage: abi_smoke <problem.bin>      prints "iterations=<k> persistent_ctas=<g> objective=<v>" and exits 0 on success
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "proxb200.h"

#define CHECK(call)                                                              \
  do {                                                                           \
    int rc__ = (call);                                                           \
    if (rc__ != PB_OK) {                                                         \
      fprintf(stderr, "%s -> %d: %s\n", #call, rc__, pb_last_error());            \
      return 2;                                                                  \
    }                                                                            \
  } while (0)

static int solve(pb_ctx* ctx, int persistent_mode, int64_t m, int64_t n, const void* dA, const void* db, double lam,
                 double* z_host, pb_solve_result* res) {
  void *x, *grad, *z, *z_prev, *x_next, *scratch, *r;
  const size_t nb = (size_t)n * sizeof(double);
  CHECK(pb_malloc(ctx, nb, &x));
  CHECK(pb_malloc(ctx, nb, &grad));
  CHECK(pb_malloc(ctx, nb, &z));
  CHECK(pb_malloc(ctx, nb, &z_prev));
  CHECK(pb_malloc(ctx, nb, &x_next));
  CHECK(pb_malloc(ctx, nb, &scratch));
  CHECK(pb_malloc(ctx, (size_t)m * sizeof(double), &r));
  CHECK(pb_memset_zero(ctx, x, nb));                       /* x0 = 0 */
  pb_smooth f;
  memset(&f, 0, sizeof f);
  f.kind = PB_F_LSQ_DENSE;
  f.m = m;
  f.n = n;
  f.lda = m;
  f.A = dA;
  f.b = db;
  f.r = r;
  pb_prox g;
  memset(&g, 0, sizeof g);
  g.kind = PB_PROX_L1;
  g.p0 = lam;
  pb_solve_opts o;
  memset(&o, 0, sizeof o);
  o.algorithm = PB_ALG_FFB;                                /* FastForwardBackward(tol = 1e-6), no Lf => adaptive */
  o.adaptive = 1;
  o.sequence = PB_SEQ_ADAPTIVE;
  o.maxit = 10000;
  o.tol = 1e-6;
  o.gamma = 0.0;
  o.minimum_gamma = 1e-7;
  o.reduce_gamma = 0.5;
  o.increase_gamma = 1.0;
  CHECK(pb_ctx_set_option(ctx, PB_OPT_PERSISTENT, persistent_mode));
  CHECK(pb_solve(ctx, PB_F64, n, &f, &g, &o, x, grad, z, z_prev, x_next, NULL, scratch, res));
  CHECK(pb_download(ctx, z_host, res->z, nb));
  CHECK(pb_free(ctx, x));
  CHECK(pb_free(ctx, grad));
  CHECK(pb_free(ctx, z));
  CHECK(pb_free(ctx, z_prev));
  CHECK(pb_free(ctx, x_next));
  CHECK(pb_free(ctx, scratch));
  CHECK(pb_free(ctx, r));
  return 0;
}

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s problem.bin\n", argv[0]);
    return 2;
  }
  FILE* fh = fopen(argv[1], "rb");
  if (!fh) {
    perror(argv[1]);
    return 2;
  }
  int64_t m, n;
  double lam;
  if (fread(&m, 8, 1, fh) != 1 || fread(&n, 8, 1, fh) != 1 || fread(&lam, 8, 1, fh) != 1) return 2;
  double* A = (double*)malloc((size_t)m * n * 8);
  double* b = (double*)malloc((size_t)m * 8);
  if (fread(A, 8, (size_t)m * n, fh) != (size_t)(m * n) || fread(b, 8, (size_t)m, fh) != (size_t)m) return 2;
  fclose(fh);

  int ndev = 0;
  CHECK(pb_device_count(&ndev));
  if (ndev < 1) {
    fprintf(stderr, "no CUDA device\n");
    return 3;
  }
  pb_ctx* ctx = NULL;
  CHECK(pb_ctx_create(0, NULL, 0, &ctx));                  /* the context owns its stream */
  void *dA, *db;
  CHECK(pb_malloc(ctx, (size_t)m * n * 8, &dA));
  CHECK(pb_malloc(ctx, (size_t)m * 8, &db));
  CHECK(pb_upload(ctx, dA, A, (size_t)m * n * 8));
  CHECK(pb_upload(ctx, db, b, (size_t)m * 8));

  /* 1. one fused FISTA step (K2) on x = 1, grad = A'(A x - b), z_prev = 0, checked against the same arithmetic in C */
  {
    void *x, *grad, *zp, *z, *xn, *r;
    const size_t nb = (size_t)n * 8;
    CHECK(pb_malloc(ctx, nb, &x));
    CHECK(pb_malloc(ctx, nb, &grad));
    CHECK(pb_malloc(ctx, nb, &zp));
    CHECK(pb_malloc(ctx, nb, &z));
    CHECK(pb_malloc(ctx, nb, &xn));
    CHECK(pb_malloc(ctx, (size_t)m * 8, &r));
    double* ones = (double*)malloc(nb);
    for (int64_t j = 0; j < n; ++j) ones[j] = 1.0;
    CHECK(pb_upload(ctx, x, ones, nb));
    CHECK(pb_memset_zero(ctx, zp, nb));
    CHECK(pb_lsq_dense_residual(ctx, PB_F64, m, n, dA, m, x, db, r));
    CHECK(pb_lsq_dense_gradient(ctx, PB_F64, m, n, dA, m, r, grad));
    pb_prox g;
    memset(&g, 0, sizeof g);
    g.kind = PB_PROX_L1;
    g.p0 = lam;
    const double gamma = 1e-3, beta = 0.25;
    CHECK(pb_ffb_step(ctx, PB_F64, n, x, grad, zp, gamma, beta, &g, NULL, z, NULL, xn));
    double sc[PB_NSCALARS];
    CHECK(pb_read_scalars(ctx, sc));
    double *gh = (double*)malloc(nb), *zh = (double*)malloc(nb), *xnh = (double*)malloc(nb);
    CHECK(pb_download(ctx, gh, grad, nb));
    CHECK(pb_download(ctx, zh, z, nb));
    CHECK(pb_download(ctx, xnh, xn, nb));
    double resinf = 0.0;
    for (int64_t j = 0; j < n; ++j) {
      volatile double gg = gamma * gh[j];
      volatile double y = 1.0 - gg;
      const double gl = gamma * lam;
      volatile double zz = y + (y <= -gl ? gl : (y >= gl ? -gl : -y));
      volatile double d = zz - 0.0;
      volatile double bd = beta * d;
      volatile double xx = zz + bd;
      if (zz != zh[j] || xx != xnh[j]) {
        fprintf(stderr, "fused step mismatch at %lld: z %.17g vs %.17g, x_next %.17g vs %.17g\n", (long long)j, zh[j], (double)zz, xnh[j], (double)xx);
        return 1;
      }
      if (fabs(1.0 - zz) > resinf) resinf = fabs(1.0 - zz);
    }
    if (sc[PB_S_RESINF] != resinf) {
      fprintf(stderr, "norm(res, Inf) mismatch: %.17g vs %.17g\n", sc[PB_S_RESINF], resinf);
      return 1;
    }
    free(ones); free(gh); free(zh); free(xnh);
    CHECK(pb_free(ctx, x)); CHECK(pb_free(ctx, grad)); CHECK(pb_free(ctx, zp)); CHECK(pb_free(ctx, z)); CHECK(pb_free(ctx, xn)); CHECK(pb_free(ctx, r));
  }

  /* 2. the whole solve: one persistent kernel (auto) and one kernel per operation (-1) -- same iterations, same bits */
  pb_solve_result r_auto, r_multi;
  double* z_auto = (double*)malloc((size_t)n * 8);
  double* z_multi = (double*)malloc((size_t)n * 8);
  if (solve(ctx, 0, m, n, dA, db, lam, z_auto, &r_auto)) return 2;
  if (solve(ctx, -1, m, n, dA, db, lam, z_multi, &r_multi)) return 2;
  if (r_auto.iterations != r_multi.iterations || memcmp(z_auto, z_multi, (size_t)n * 8) != 0 || r_auto.gamma != r_multi.gamma ||
      r_auto.f_x != r_multi.f_x || r_auto.g_z != r_multi.g_z) {
    fprintf(stderr, "persistent (%lld it) and multi-kernel (%lld it) solves differ\n", (long long)r_auto.iterations, (long long)r_multi.iterations);
    return 1;
  }
  if (r_auto.persistent_ctas < 1 || r_multi.persistent_ctas != 0) {
    fprintf(stderr, "unexpected driver selection: %d / %d\n", r_auto.persistent_ctas, r_multi.persistent_ctas);
    return 1;
  }
  double obj = 0.0;
  for (int64_t i = 0; i < m; ++i) {
    double s = -b[i];
    for (int64_t j = 0; j < n; ++j) s += A[i + j * m] * z_auto[j];
    obj += 0.5 * s * s;
  }
  for (int64_t j = 0; j < n; ++j) obj += lam * fabs(z_auto[j]);
  printf("iterations=%lld persistent_ctas=%d objective=%.17g\n", (long long)r_auto.iterations, r_auto.persistent_ctas, obj);
  CHECK(pb_free(ctx, dA));
  CHECK(pb_free(ctx, db));
  CHECK(pb_ctx_destroy(ctx));
  free(A); free(b); free(z_auto); free(z_multi);
  return 0;
}
