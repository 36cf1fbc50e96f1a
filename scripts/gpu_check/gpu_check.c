/* scripts/gpu_check/gpu_check.c -- a torch-free, python-free GPU parity check through the C ABI (include/grb.h).
 *
 * For each case file exported by scripts/gpu_check/export_cases.py from tests/golden (npz files; the UNMODIFIED reference's weights
 * on fixed points): parse the card, create the engine on device 0, install the screening table if the case is screened,
 * evaluate the points with grb_eval (host buffers) and compare flags bit-exactly and weights to max(1e-10, 8 kappa).
 * Then one VEGAS iteration (uniform grid, 2^18 points) per distinct process as a smoke test of the k_points<VEGAS> variant.
 * Exists because a fresh GPU box spends ~1 min importing torch: this binary starts in a second.
 *   gcc -O2 -Iinclude scripts/gpu_check/gpu_check.c -o scripts/gpu_check/gpu_check -Lgraniitti_b200 -lgrb -Wl,-rpath,'$ORIGIN/../../graniitti_b200' -lm
 *   scripts/gpu_check/gpu_check scripts/gpu_check/data/msa.bin scripts/gpu_check/data/<case>.bin ...            */
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "grb.h"

static int cmp_d(const void *a, const void *b) { const double x = *(const double *)a, y = *(const double *)b; return x < y ? -1 : x > y; }

int main(int argc, char **argv) {
  if (argc < 3) { fprintf(stderr, "usage: gpu_check msa.bin case.bin ...\n"); return 2; }
  FILE *f = fopen(argv[1], "rb");
  if (!f) { perror(argv[1]); return 2; }
  int64_t np = 0;
  if (fread(&np, 8, 1, f) != 1) return 2;
  double *msa = (double *)malloc(sizeof(double) * 3 * np);
  if (fread(msa, 8, 3 * np, f) != (size_t)(3 * np)) return 2;
  fclose(f);
  int n_fail = 0;
  for (int c = 2; c < argc; ++c) {
    f = fopen(argv[c], "rb");
    if (!f) { perror(argv[c]); return 2; }
    int64_t h[4];
    char card[256], proc[512];
    if (fread(h, 8, 4, f) != 4 || fread(card, 1, 256, f) != 256 || fread(proc, 1, 512, f) != 512) return 2;
    const int64_t n = h[0], dim = h[1], dt = h[2], pom = h[3];
    double *u = (double *)malloc(8 * n * dim), *tape = (double *)malloc(8 * n * dt), *W = (double *)malloc(8 * n), *kap = (double *)malloc(8 * n);
    uint8_t *fl = (uint8_t *)malloc(n);
    if (fread(u, 8, n * dim, f) != (size_t)(n * dim) || fread(tape, 8, n * dt, f) != (size_t)(n * dt) || fread(W, 8, n, f) != (size_t)n ||
        fread(kap, 8, n, f) != (size_t)n || fread(fl, 1, n, f) != (size_t)n) return 2;
    fclose(f);
    grb_engine *E = NULL;
    int rc = grb_create_from_card(card, proc, (int)pom, NULL, 0, &E);
    if (rc != GRB_OK) { printf("%s: CREATE FAILED rc=%d\n", argv[c], rc); ++n_fail; continue; }
    if (pom) {
      rc = grb_set_eikonal(E, msa, msa + np, msa + 2 * np, (int)np);
      if (rc != GRB_OK) { printf("%s: set_eikonal rc=%d %s\n", argv[c], rc, grb_last_error(E)); ++n_fail; grb_destroy(E); continue; }
    }
    double *w = (double *)calloc(n, 8), *rel = (double *)calloc(n, 8);
    uint8_t *g = (uint8_t *)calloc(n, 1);
    rc = grb_eval(E, n, u, (int)dim, tape, (int)dt, w, g, NULL);
    if (rc != GRB_OK) { printf("%s: grb_eval rc=%d %s\n", argv[c], rc, grb_last_error(E)); ++n_fail; grb_destroy(E); continue; }
    int64_t nz = 0, bad_flags = 0, bad_zero = 0, over10 = 0, over_k = 0;
    double mx = 0;
    for (int64_t i = 0; i < n; ++i) {
      if ((g[i] & 1) != (fl[i] & 1) || ((fl[i] & 1) && g[i] != fl[i])) ++bad_flags;
      if ((w[i] != 0) != (W[i] != 0)) { ++bad_zero; continue; }
      if (W[i] == 0) continue;
      const double r = fabs(w[i] - W[i]) / fabs(W[i]);
      rel[nz++] = r;
      if (r > mx) mx = r;
      if (r > 1e-10) ++over10;
      if (r > fmax(1e-10, 8 * kap[i])) ++over_k;
    }
    qsort(rel, nz, 8, cmp_d);
    const int ok = !bad_flags && !bad_zero && !over_k && nz > 0;
    printf("%-52s n %5lld nz %5lld flags!= %lld zero!= %lld  median %.2e max %.2e  >1e-10: %lld  >max(1e-10,8k): %lld  %s\n", strrchr(argv[c], '/') ? strrchr(argv[c], '/') + 1 : argv[c],
           (long long)n, (long long)nz, (long long)bad_flags, (long long)bad_zero, nz ? rel[nz / 2] : 0.0, mx, (long long)over10, (long long)over_k, ok ? "OK" : "FAIL");
    if (!ok) ++n_fail;
    /* VEGAS smoke: uniform grid, 2^18 points */
    {
      grb_process_desc d;
      grb_get_desc(E, &d);
      const int bins = 128, D = d.lipsdim;
      double *xmat = (double *)malloc(8 * bins * D);
      for (int b = 0; b < bins; ++b)
        for (int k = 0; k < D; ++k) xmat[b * D + k] = (b + 1) / (double)bins;
      grb_accum *acc = (grb_accum *)calloc(1, sizeof(grb_accum));
      rc = grb_vegas_iterate(E, xmat, bins, D, 0, 1 << 18, 12345678, 0, GRB_MODE_INTEGRATE, 0.0, acc);
      double ms = 0; int nl = 0;
      grb_last_timing(E, &ms, &nl);
      const double N = acc->evaluations, mean = acc->fsum / N, var = (acc->f2sum / N - mean * mean) / N;
      printf("    vegas rc=%d evals %.0f kin %.0f fid %.0f amp_ok %.0f  sigma(flat grid) %.6e +- %.2e  %.3f ms %d launches\n", rc, N, acc->n_kin_ok, acc->n_fid_ok,
             acc->n_amp_ok, mean, sqrt(var > 0 ? var : 0), ms, nl);
      if (rc != GRB_OK || !(acc->fsum > 0)) ++n_fail;
      /* GRB_CHECK_BENCH=<log2 points>: device-timed VEGAS iterations on the uniform grid (3 warm-up + 5 timed; grb_last_timing =
       * CUDA events around the kernels of the iteration on the engine's stream) */
      const char *be = getenv("GRB_CHECK_BENCH");
      if (be && rc == GRB_OK) {
        const int64_t nb = (int64_t)1 << atoi(be);
        double tot = 0, best = 1e30, items = 0;
        for (int it = 0; it < 8; ++it) {
          rc = grb_vegas_iterate(E, xmat, bins, D, 0, nb, 777, (uint64_t)it, GRB_MODE_INTEGRATE, 0.0, acc);
          grb_last_timing(E, &ms, &nl);
          if (it >= 3) { tot += ms; if (ms < best) best = ms; items = acc->n_fid_ok; }
        }
        printf("    bench 2^%s points/iteration: %.3f ms mean, %.3f ms best -> %.4e points/s; %.0f points pass the cuts -> %.4e amplitude%s evaluations/s\n", be, tot / 5,
               best, nb / (tot / 5 * 1e-3), items, items / (tot / 5 * 1e-3), pom ? " x screening-loop" : "");
      }
      free(xmat); free(acc);
    }
    grb_destroy(E);
    free(u); free(tape); free(W); free(kap); free(fl); free(w); free(rel); free(g);
    fflush(stdout);
  }
  printf("gpu_check: %d failing case(s)\n", n_fail);
  return n_fail ? 1 : 0;
}
