// grb_eikonal.h -- host interface of the device eikonal table builder (grb_eikonal.cu)
#pragma once
#include <cuda_runtime.h>

#include <string>
#include <vector>

#include "../../include/grb.h"

namespace grb {

struct EikonalTables {
  std::vector<double> bt, om_re, om_im;   // opacity Omega(s,b)            (reference MEikonal::MBT)
  std::vector<double> kt2, sa_re, sa_im;  // screening amplitude A(s,kt^2) (reference MEikonal::MSA)
  double xs[3] = {0, 0, 0};               // sigma_tot, sigma_el, sigma_inel [barn]
  float  kernel_ms = 0;                   // device time of the two Fourier-Bessel launches
};

// returns 0, a cudaError_t (> 0), or -1 bad argument / -3 unsupported model switch
int build_eikonal_tables(const grb_param_soft &S, double s, int n_bt, int n_kt2, cudaStream_t st, EikonalTables &T,
                         std::string &err);

}  // namespace grb
