// oracle/ref_harness/ref_vegas.cc -- TEST INFRASTRUCTURE ONLY.  Drives the reference's header-only VEGAS grid refinement
// (gra::VEGASData::OptimizeGrid / Rebin, include/Graniitti/MVEGAS.h:87-142, unmodified) on caller-supplied grids and
// accumulated <f^2> histograms, so that grb_vegas_optimize_grid can be pinned bit for bit (tests/test_vegas_grid.py).
//   ref_vegas <in.bin> <out.bin>    in: int64 bins, int64 dim, double lambda, xmat[bins][dim], f2mat[bins][dim]; out: xmat
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <stdexcept>
#include <string>
#include <vector>

#include "Graniitti/MVEGAS.h"

int main(int argc, char **argv) {
  if (argc < 3) return 2;
  FILE *fi = fopen(argv[1], "rb");
  if (!fi) return 2;
  int64_t bins, dim;
  double  lambda;
  if (fread(&bins, 8, 1, fi) != 1 || fread(&dim, 8, 1, fi) != 1 || fread(&lambda, 8, 1, fi) != 1) return 2;
  std::vector<double> x(bins * dim), f2(bins * dim);
  if (fread(x.data(), 8, x.size(), fi) != x.size() || fread(f2.data(), 8, f2.size(), fi) != f2.size()) return 2;
  fclose(fi);
  gra::VEGASPARAM par;
  par.BINS = (unsigned)bins; par.LAMBDA = lambda;
  gra::VEGASData VD;
  VD.InitRegion((unsigned)dim);
  VD.ClearAll(par);
  VD.BINS_prev = par.BINS;
  for (int64_t i = 0; i < bins; ++i)
    for (int64_t j = 0; j < dim; ++j) { VD.xmat[i][j] = x[i * dim + j]; VD.f2mat[i][j] = f2[i * dim + j]; }
  VD.OptimizeGrid(par);
  for (int64_t i = 0; i < bins; ++i)
    for (int64_t j = 0; j < dim; ++j) x[i * dim + j] = VD.xmat[i][j];
  FILE *fo = fopen(argv[2], "wb");
  fwrite(x.data(), 8, x.size(), fo);
  fclose(fo);
  return 0;
}
