/* cabi_smoke.c — the C-ABI of libblmm_b200.so driven from plain C, with the structs of include/blmm_b200.h built by
 * hand exactly as the Julia shim's `ccall`s (bulklmm.jl_b200/julia/BulkLMMB200.jl) and any other FFI would build them:
 * no Python, no torch, no ctypes in between.
 *
 *     gcc -O1 -I include tests/cabi_smoke.c -o /tmp/cabi_smoke -L <libdir> -lblmm_b200 -Wl,-rpath,<libdir> -lm
 *     /tmp/cabi_smoke tests/golden/cabi_smoke.bin [ndev]
 *
 * Reads the committed fixture (inputs + ORACLE outputs, tests/golden/make_cabi_smoke_fixture.py), runs
 *   blmm_create / blmm_create_multi -> blmm_bulkscan(alt-grid) -> blmm_scan_perms -> blmm_thresholds -> blmm_destroy
 * with pageable malloc'ed host buffers, and compares: alt-grid LODs within 1e-8 * max(1, |ref|), h2 panel equal up to
 * rounding-level ties (< 0.1 % of entries), permutation LODs within 1e-5 (two FP64 Brent fits), per-permutation
 * maxima == column maxima.  Exit status 0 = pass.  Also checks an error path: the reference's message for h2 = 1. */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "blmm_b200.h"

static double* rd(FILE* f, size_t n) {
  double* a = (double*)malloc(n * sizeof(double));
  if (!a || fread(a, sizeof(double), n, f) != n) {
    fprintf(stderr, "fixture truncated\n");
    exit(2);
  }
  return a;
}

static double rel_err(const double* a, const double* b, size_t n) {
  double e = 0.0;
  for (size_t i = 0; i < n; ++i) {
    const double s = fabs(b[i]) > 1.0 ? fabs(b[i]) : 1.0;
    const double d = fabs(a[i] - b[i]) / s;
    if (!(d <= e)) e = d; /* NaN propagates */
  }
  return e;
}

#define CHECK(call)                                                                  \
  do {                                                                               \
    int st_ = (call);                                                                \
    if (st_ != BLMM_OK) {                                                            \
      fprintf(stderr, "%s -> %d: %s\n", #call, st_, blmm_last_error(ctx));           \
      return 1;                                                                      \
    }                                                                                \
  } while (0)

int main(int argc, char** argv) {
  if (argc < 2) {
    fprintf(stderr, "usage: %s cabi_smoke.bin [ndev]\n", argv[0]);
    return 2;
  }
  const int ndev = argc > 2 ? atoi(argv[2]) : 1;
  FILE* f = fopen(argv[1], "rb");
  if (!f) {
    perror(argv[1]);
    return 2;
  }
  int64_t hdr[5];
  if (fread(hdr, sizeof(int64_t), 5, f) != 5) return 2;
  const int64_t n = hdr[0], p = hdr[1], m = hdr[2], nperms = hdr[3], ngrid = hdr[4];
  double *Y = rd(f, n * m), *G = rd(f, n * p), *U = rd(f, n * n), *lam = rd(f, n), *grid = rd(f, ngrid);
  double *refL = rd(f, p * m), *refH = rd(f, p * m), *y1 = rd(f, n);
  int32_t* perm = (int32_t*)malloc(n * nperms * sizeof(int32_t));
  if (fread(perm, sizeof(int32_t), n * nperms, f) != (size_t)(n * nperms)) return 2;
  double *refLod = rd(f, p), *refLp = rd(f, p * nperms), *refh2 = rd(f, 1), *refs2 = rd(f, 1);
  fclose(f);

  if (blmm_abi_version() != BLMM_ABI_VERSION) {
    fprintf(stderr, "ABI version mismatch: header %d, library %d\n", BLMM_ABI_VERSION, blmm_abi_version());
    return 1;
  }
  blmm_ctx* ctx = NULL;
  int devs[64];
  for (int i = 0; i < 64; ++i) devs[i] = i;
  int st = ndev > 1 ? blmm_create_multi(&ctx, devs, ndev) : blmm_create(&ctx, 0);
  if (st != BLMM_OK) {
    fprintf(stderr, "blmm_create -> %d (no usable B200?)\n", st);
    return 1;
  }
  if (blmm_device_count(ctx) != ndev) return 1;

  double* ones = (double*)malloc(n * sizeof(double));
  for (int64_t i = 0; i < n; ++i) ones[i] = 1.0;

  /* bulkscan(Y, G, K; method = "alt-grid") */
  blmm_problem pr;
  memset(&pr, 0, sizeof pr);
  pr.n = n; pr.p = p; pr.m = m; pr.c = 1;
  pr.Y = Y; pr.G = G; pr.Covar = ones; pr.U = U; pr.lambda = lam; pr.obs_weights = NULL;
  blmm_opts op;
  memset(&op, 0, sizeof op);
  op.method = BLMM_METHOD_ALT_GRID; op.reml = 0; op.prior_variance = 1.0; op.prior_sample_size = 0.0;
  op.h2_grid = grid; op.ngrid = (int32_t)ngrid; op.optim_interval = 1; op.h2_panel_mode = BLMM_H2PANEL_REFERENCE;
  op.mem_space = BLMM_MEM_HOST; op.ld_out = 0; op.chisq_df = 0; op.log10p_out = NULL;
  double* L = (double*)malloc(p * m * sizeof(double));
  double* H = (double*)malloc(p * m * sizeof(double));
  CHECK(blmm_bulkscan(ctx, &pr, &op, L, H));
  const double eL = rel_err(L, refL, p * m);
  int64_t mism = 0;
  for (int64_t i = 0; i < p * m; ++i) mism += H[i] != refH[i];
  printf("alt-grid: max rel err %.3e, h2_panel mismatches %lld of %lld\n", eL, (long long)mism, (long long)(p * m));
  if (!(eL < 1e-8) || mism * 1000 > p * m) return 1;

  /* scan(y, G, K; permutation_test = true) */
  pr.Y = y1; pr.m = 1;
  op.method = BLMM_METHOD_NULL_EXACT; op.prior_variance = 0.0; op.h2_grid = NULL; op.ngrid = 0;
  double* lod = (double*)malloc(p * sizeof(double));
  double* Lp = (double*)malloc(p * nperms * sizeof(double));
  double* mx = (double*)malloc(nperms * sizeof(double));
  double s2 = 0.0, h2 = 0.0;
  CHECK(blmm_scan_perms(ctx, &pr, &op, perm, nperms, lod, Lp, mx, &s2, &h2));
  const double e1 = rel_err(lod, refLod, p), e2 = rel_err(Lp, refLp, p * nperms);
  printf("perms: lod err %.3e, L_perms err %.3e, h2 %.10f (ref %.10f), sigma2 %.10f (ref %.10f)\n", e1, e2, h2, *refh2,
         s2, *refs2);
  if (!(e1 < 1e-5) || !(e2 < 1e-5) || !(fabs(h2 - *refh2) < 2e-6) || !(fabs(s2 - *refs2) < 1e-5 * *refs2)) return 1;
  for (int64_t s = 0; s < nperms; ++s) {
    double cm = 0.0;
    for (int64_t i = 0; i < p; ++i) cm = Lp[s * p + i] > cm ? Lp[s * p + i] : cm;
    if (cm != mx[s]) {
      fprintf(stderr, "per-permutation maximum %lld differs from the column maximum\n", (long long)s);
      return 1;
    }
  }
  double sig[2] = {0.10, 0.05}, thr[2] = {0, 0};
  CHECK(blmm_thresholds(ctx, mx, nperms, sig, 2, thr, BLMM_MEM_HOST));
  printf("thresholds: %.6f %.6f\n", thr[0], thr[1]);
  if (!(thr[1] >= thr[0]) || !(thr[0] > 0.0)) return 1;

  /* 1-based indices (the natural Julia slip) are refused, not read */
  for (int64_t i = 0; i < n * nperms; ++i) perm[i] += 1;
  st = blmm_scan_perms(ctx, &pr, &op, perm, nperms, lod, Lp, mx, &s2, &h2);
  if (st != BLMM_E_INVALID) {
    fprintf(stderr, "1-based perm_idx: expected BLMM_E_INVALID, got %d\n", st);
    return 1;
  }

  /* the reference's own error string crosses the boundary */
  double badgrid[2] = {0.5, 1.0};
  pr.Y = Y; pr.m = m;
  op.method = BLMM_METHOD_NULL_GRID; op.h2_grid = badgrid; op.ngrid = 2; op.prior_variance = 1.0;
  st = blmm_bulkscan(ctx, &pr, &op, L, H);
  if (st != BLMM_E_H2_ONE || strcmp(blmm_last_error(ctx), "Heritability of 1 is not allowed.") != 0) {
    fprintf(stderr, "h2 = 1: got %d '%s'\n", st, blmm_last_error(ctx));
    return 1;
  }
  blmm_destroy(ctx);
  printf("cabi_smoke ok (ndev %d)\n", ndev);
  return 0;
}
