// grb_eikonal.cu -- eikonal screening tables on the device.
//
// The reference builds the proton opacity Omega(s,b) and the screening amplitude A_eik(s,kt^2) with
// 2 x 10 001 Fourier-Bessel integrals of 10 001 nodes each, one std::async future per table point
// (reference src/MEikonal.cc:177-242 S3Density/S3Screening, :379-401 S3CalculateArray; ~6 s on 8 cores).
// Both are the same transform  out_i = sum_j c_j J0(x_i y_j)  once everything that does not depend on the
// output index is folded into c_j on the host (elastic amplitude with the non-linear pomeron trajectory, the
// eikonalised amplitude i(1 - e^{i Omega/2}), Simpson weights, Jacobians): one kernel, one CTA per output
// point, J0 by the reference's own rational/asymptotic approximation (libs/extmath.hpp:12-49, kept: it is
// part of the numbers).  The 15-significant-digit ASCII round trip of the reference's table cache
// (MEikonal.cc:84-91,404-436) is reproduced on the host so that the tables are interchangeable with files
// written by the reference.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <string>
#include <vector>

#include "grb_dev.cuh"
#include "grb_eikonal.h"

namespace grb {

__device__ __forceinline__ double bessj0_ref(double X) {
  if (X == 0.0) return 1.0;
  const double AX = fabs(X);
  if (AX < 8.0) {
    const double Y  = X * X;
    const double FR = 57568490574.0 + Y * (-13362590354.0 + Y * (651619640.7 + Y * (-11214424.18 + Y * (77392.33017 + Y * -184.9052456))));
    const double FS = 57568490411.0 + Y * (1029532985.0 + Y * (9494680.718 + Y * (59272.64853 + Y * (267.8532712 + Y * 1.0))));
    return FR / FS;
  }
  const double Z = 8.0 / AX, Y = Z * Z, XX = AX - 0.785398164;
  const double FP = 1.0 + Y * (-0.1098628627E-2 + Y * (0.2734510407E-4 + Y * (-0.2073370639E-5 + Y * 0.2093887211E-6)));
  const double FQ = -0.1562499995E-1 + Y * (0.1430488765E-3 + Y * (-0.6911147651E-5 + Y * (0.7621095161E-6 + Y * -0.9349451520E-7)));
  double sn, cs;
  sincos(XX, &sn, &cs);
  return sqrt(0.636619772 / AX) * (FP * cs - Z * FQ * sn);
}

// out[i] = sum_j (c_re[j] + i c_im[j]) * J0(x[i] * y[j]);  one CTA per i
__global__ void __launch_bounds__(256) k_fourier_bessel(const double *__restrict__ x, int n_out, const double *__restrict__ y,
                                                        const double *__restrict__ c_re, const double *__restrict__ c_im,
                                                        int n_in, double *__restrict__ out_re, double *__restrict__ out_im) {
  __shared__ double sr[8], si[8];
  for (int i = blockIdx.x; i < n_out; i += gridDim.x) {
    const double xi = x[i];
    double ar = 0.0, ai = 0.0;
    for (int j = threadIdx.x; j < n_in; j += blockDim.x) {
      const double b = bessj0_ref(xi * y[j]);
      ar = fma(c_re[j], b, ar);
      ai = fma(c_im[j], b, ai);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      ar += __shfl_xor_sync(0xffffffffu, ar, o);
      ai += __shfl_xor_sync(0xffffffffu, ai, o);
    }
    if ((threadIdx.x & 31) == 0) { sr[threadIdx.x >> 5] = ar; si[threadIdx.x >> 5] = ai; }
    __syncthreads();
    if (threadIdx.x == 0) {
      double tr = 0.0, ti = 0.0;
      for (int w = 0; w < 8; ++w) { tr += sr[w]; ti += si[w]; }
      out_re[i] = tr;
      out_im[i] = ti;
    }
    __syncthreads();
  }
}

namespace {
constexpr double kPI = GRB_PI;
constexpr double kMpi = 0.13957018;  // PDG::mpi (reference MPDG.h:56)
const double kGeV2barn = ((6.582119514E-25 * 2.99792458E8) * (6.582119514E-25 * 2.99792458E8)) / 1E-28;
const double kGeV2fm = (6.582119514E-25 * 2.99792458E8) * 1E15;

double msq(double v) { return std::sqrt(v > 0.0 ? v : 0.0); }
double round15(double v) {
  char buf[64];
  snprintf(buf, sizeof buf, "%.15g", v);
  return strtod(buf, nullptr);
}
// non-linear pomeron trajectory with pion loop (reference src/MForm.cc:402-431)
double pom_alpha(const grb_param_soft &S, double t) {
  const double BETApi = (2.0 / 3.0) * S.gN_P;
  const double PIC = ((BETApi * BETApi) * (kMpi * kMpi)) / (32.0 * (kPI * (kPI * kPI)));
  const double tau = 4.0 * (kMpi * kMpi) / std::fabs(t), sqrtau = msq(1 + tau);
  const double Fpi = 1.0 / (1.0 - t / S.fc3);
  const double h = (4.0 / tau) * (Fpi * Fpi) *
                   (2.0 * tau - std::pow(1.0 + tau, 3.0 / 2.0) * std::log((sqrtau + 1.0) / (sqrtau - 1.0)) + std::log((1.0 * 1.0) / (kMpi * kMpi)));
  return 1.0 + S.DELTA_P + S.ALPHA_P * t - PIC * h;
}
// single pomeron exchange elastic amplitude (reference src/MEikonal.cc:129-153)
void single_amp(const grb_param_soft &S, double s, double t, double &re, double &im) {
  const double F = (1.0 / (1 - t / S.fc1)) * (1.0 / (1 - t / S.fc2));
  const double a = pom_alpha(S, t), denom = std::sin(kPI * a), y = (-1.0 * kPI) * a;
  const double f = (S.gN_P * S.gN_P) * F * F, p = std::pow(s / 1.0, a);
  re = f * (-(1.0 + std::cos(y)) / denom) * p;
  im = f * (-(std::sin(y)) / denom) * p;
}
// composite Simpson coefficient of node j of N intervals times h/3 (reference MMath.h:378-396)
double simpson_w(int j, int N, double h) { return ((j == 0 || j == N) ? 1.0 : ((j % 2) ? 4.0 : 2.0)) * h / 3.0; }
}  // namespace

int build_eikonal_tables(const grb_param_soft &S, double s, int n_bt, int n_kt2, cudaStream_t st, EikonalTables &T,
                         std::string &err) {
  if (S.odderon_on) { err = "ODDERON_ON eikonal is not supported"; return -3; }
  if (n_bt < 2 || n_kt2 < 2) { err = "table sizes must be >= 2"; return -1; }
  const int NFB = 10000;  // MEikonalNumerics::FBIntegralN (MEikonal.h:40)
  const double MinBT = 1E-6, MaxBT = 10.0 / kGeV2fm, MinKT2 = 1E-6, MaxKT2 = 25.0, FBmin = 1E-9, FBmax = 30.0;
  const int nmax = std::max(std::max(n_bt, n_kt2), NFB) + 1;
  struct Dev {  // x | y | c_re | c_im | out_re | out_im
    double *p = nullptr;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};
    ~Dev() { cudaFree(p); for (auto e : ev) if (e) cudaEventDestroy(e); }
  } D;
  cudaError_t ce;
#define CKE(c) if ((ce = (c)) != cudaSuccess) { err = std::string(#c) + ": " + cudaGetErrorString(ce); return (int)ce; }
  CKE(cudaMalloc(&D.p, sizeof(double) * nmax * 6));
  for (auto &e : D.ev) CKE(cudaEventCreate(&e));
  double *d_in = D.p, *d_out = D.p + 4 * (size_t)nmax;
  std::vector<double> in(4 * (size_t)nmax), out(2 * (size_t)nmax);
  double *x = in.data(), *y = x + nmax, *cr = y + nmax, *ci = cr + nmax, *o_r = out.data(), *o_i = o_r + nmax;
  auto transform = [&](int n_out, int which) -> int {
    CKE(cudaMemcpyAsync(d_in, in.data(), sizeof(double) * in.size(), cudaMemcpyHostToDevice, st));
    CKE(cudaEventRecord(D.ev[which], st));
    k_fourier_bessel<<<std::min(n_out, 148 * 8), 256, 0, st>>>(d_in, n_out, d_in + nmax, d_in + 2 * (size_t)nmax,
                                                              d_in + 3 * (size_t)nmax, NFB + 1, d_out, d_out + nmax);
    CKE(cudaGetLastError());
    CKE(cudaEventRecord(D.ev[which + 1], st));
    CKE(cudaMemcpyAsync(out.data(), d_out, sizeof(double) * out.size(), cudaMemcpyDeviceToHost, st));
    CKE(cudaStreamSynchronize(st));
    float ms = 0;
    CKE(cudaEventElapsedTime(&ms, D.ev[which], D.ev[which + 1]));
    T.kernel_ms += ms;
    return 0;
  };
  T.kernel_ms = 0;

  // ---- opacity Omega(s, b): S3Density (MEikonal.cc:177-204) ----------------------------------------------
  const double ktS = (FBmax - FBmin) / NFB, FACT1 = 1 / (2.0 * kPI * s);
  for (int j = 0; j <= NFB; ++j) {
    const double kt = FBmin + j * ktS;
    double ar, ai;
    single_amp(S, s, -(kt * kt), ar, ai);
    const double w = simpson_w(j, NFB, ktS) * kt * FACT1;
    y[j] = kt; cr[j] = ar * w; ci[j] = ai * w;
  }
  const double bS = (MaxBT - MinBT) / n_bt;
  for (int i = 0; i <= n_bt; ++i) x[i] = MinBT + i * bS;
  if (int rc = transform(n_bt + 1, 0)) return rc;
  std::vector<double> &bx = T.bt, &br = T.om_re, &bi = T.om_im;
  bx.resize(n_bt + 1); br.resize(n_bt + 1); bi.resize(n_bt + 1);
  for (int i = 0; i <= n_bt; ++i) { bx[i] = round15(x[i]); br[i] = round15(o_r[i]); bi[i] = round15(o_i[i]); }
  auto interp = [&](double a, double &re, double &im) {  // IArray1D::Interpolate1D, MEikonal.cc:489-512
    const double STEP = (MaxBT - MinBT) / n_bt;
    if (a < MinBT) a = MinBT;
    int i = (int)std::floor((a - MinBT) / STEP);
    if (i < 0) i = 0;
    if (i >= n_bt) i = n_bt - 1;
    const double dx = a - bx[i], h = bx[i + 1] - bx[i];
    re = br[i] + dx * ((br[i + 1] - br[i]) / h);
    im = bi[i] + dx * ((bi[i + 1] - bi[i]) / h);
  };
  auto a_el = [&](double bt, double &re, double &im) {  // i (1 - exp(i Omega / 2)), MEikonal.cc:207-213
    double orr, oi;
    interp(bt, orr, oi);
    const double er = std::exp(-oi / 2.0), c = std::cos(orr / 2.0), sn = std::sin(orr / 2.0);
    re = er * sn;
    im = 1.0 - er * c;
  };
  // ---- sigma_tot / el / inel: S3CalcXS (single channel, MEikonal.cc:262-340), host, 6001 nodes ------------
  {
    const int N = 2 * 3000;
    const double STEP = (MaxBT - MinBT) / N;
    double t = 0, e = 0, in = 0;
    for (int i = 0; i <= N; ++i) {
      const double bt = MinBT + i * STEP;
      double re, im;
      a_el(bt, re, im);
      const double a2 = re * re + im * im, w = simpson_w(i, N, STEP) * bt;
      t += w * 2.0 * im; e += w * a2; in += w * (2.0 * im - a2);
    }
    T.xs[0] = 2.0 * kPI * t * kGeV2barn; T.xs[1] = 2.0 * kPI * e * kGeV2barn; T.xs[2] = 2.0 * kPI * in * kGeV2barn;
  }
  // ---- screening amplitude A_eik(s, kt^2): S3Screening (MEikonal.cc:216-242) -------------------------------
  const double bS2 = (MaxBT - MinBT) / NFB, FACT2 = (2.0 * s) * (2.0 * kPI);
  for (int j = 0; j <= NFB; ++j) {
    const double bt = MinBT + j * bS2;
    double re, im;
    a_el(bt, re, im);
    const double w = simpson_w(j, NFB, bS2) * bt * FACT2;
    y[j] = bt; cr[j] = re * w; ci[j] = im * w;
  }
  const double kS = (MaxKT2 - MinKT2) / n_kt2;
  T.kt2.resize(n_kt2 + 1); T.sa_re.resize(n_kt2 + 1); T.sa_im.resize(n_kt2 + 1);
  for (int i = 0; i <= n_kt2; ++i) { T.kt2[i] = MinKT2 + i * kS; x[i] = msq(T.kt2[i]); }
  if (int rc = transform(n_kt2 + 1, 1)) return rc;
  for (int i = 0; i <= n_kt2; ++i) { T.kt2[i] = round15(T.kt2[i]); T.sa_re[i] = round15(o_r[i]); T.sa_im[i] = round15(o_i[i]); }
#undef CKE
  return 0;
}

}  // namespace grb
