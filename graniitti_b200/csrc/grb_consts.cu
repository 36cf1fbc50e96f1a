// grb_consts.cu -- the __constant__ blocks shared by every kernel translation unit (relocatable device code), their
// uploads, the k_points dispatch and the fp64 FMA peak microbenchmark.
#include <cuda_runtime.h>

#include <cstdlib>

#include "grb_point.cuh"
#include "grb_fastmath.cuh"

namespace grb {

__constant__ DevProc c_proc;
__constant__ DevLoop c_loop;
__constant__ DevRes  c_res;
__constant__ FastConst c_fm = {-1.0 / 6.0, 1.0 / 5.0, 1.0 / 3.0, GRB_LN2,
                               GRB_64_OVER_LN2, GRB_LN2_64_LO, 1.0 / 120.0, 1.0 / 24.0, 1.0 / 6.0,
                               GRB_64_OVER_PI, GRB_PI_64_LO, -1.0 / 5040.0, -1.0 / 720.0};
// lookup tables of grb_fastmath.cuh (global memory, read through the L1 / read-only path: indices diverge in a warp)
__device__ const double2 g_logtab[128] = {GRB_LOGTAB_ROWS};
__device__ const double  g_exptab[64]  = {GRB_EXPTAB_ROWS};
__device__ const double2 g_sincostab[128] = {GRB_SINCOSTAB_ROWS};

cudaError_t upload_proc(const DevProc &p, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_proc, &p, sizeof(DevProc), 0, cudaMemcpyHostToDevice, st);
}
cudaError_t upload_res(const DevRes &r, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_res, &r, sizeof(DevRes), 0, cudaMemcpyHostToDevice, st);
}
cudaError_t upload_loop(const DevLoop &l, cudaStream_t st) {
  return cudaMemcpyToSymbolAsync(c_loop, &l, sizeof(DevLoop), 0, cudaMemcpyHostToDevice, st);
}

template <int AMP>
void launch_points_amp(const PointArgs &a, int n_nodes, unsigned blocks, int threads, size_t smem, cudaStream_t st);

int launch_points(const PointArgs &a_in, int n_nodes, int amp_kind, int sm_count, cudaStream_t st) {
  PointArgs a = a_in;
  a.sm_count  = sm_count;
  const int     threads = 256;
  int64_t       blocks  = (a.n + threads - 1) / threads;
  const int64_t cap     = (int64_t)sm_count * 8;
  if (blocks > cap) blocks = cap;
  if (a.side_launch) blocks = sm_count;
  if (blocks < 1) blocks = 1;
  const size_t smem = a.vegas ? sizeof(double) * 2 * a.bins * a.dim : 0;   // fm, f2m [bins][lipsdim] (hist_add indexing)
  switch (amp_kind) {
    case AMP_TENSOR_2XS: launch_points_amp<AMP_TENSOR_2XS>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_REGGE_RES: launch_points_amp<AMP_REGGE_RES>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_REGGE_RES_CON: launch_points_amp<AMP_REGGE_RES_CON>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_YY_LL: launch_points_amp<AMP_YY_LL>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_TENSOR_2XV: launch_points_amp<AMP_TENSOR_2XV>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_TENSOR_2XF: launch_points_amp<AMP_TENSOR_2XF>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_YY_QED: launch_points_amp<AMP_YY_QED>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_REGGE_ME6: launch_points_amp<AMP_REGGE_ME6>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_REGGE_ME8: launch_points_amp<AMP_REGGE_ME8>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_TENSOR_RES: launch_points_amp<AMP_TENSOR_RES>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    case AMP_TENSOR_RES_CON: launch_points_amp<AMP_TENSOR_RES_CON>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
    default: launch_points_amp<AMP_REGGE>(a, n_nodes, (unsigned)blocks, threads, smem, st); break;
  }
  return 1;
}

// ---- fp64 FMA peak microbenchmark (roofline denominator) -------------------------------------------------
__global__ void __launch_bounds__(256) k_dfma_peak(double *out, int iters) {
  double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  const double b = 1.0000001, c = 1e-9;
  for (int i = 0; i < iters; ++i) {
    a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
    a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
  }
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = ((a0 + a1) + (a2 + a3)) + ((a4 + a5) + (a6 + a7));
}

cudaError_t run_dfma_peak(int sm_count, double *tflops) {
  const int blocks = sm_count * 8, threads = 256, iters = 1 << 16;
  double *  d      = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(double) * blocks * threads);
  if (e != cudaSuccess) return e;
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0);
  cudaEventCreate(&t1);
  k_dfma_peak<<<blocks, threads>>>(d, 1024);  // warm-up
  double best = 0.0;
  for (int rep = 0; rep < 5; ++rep) {
    cudaEventRecord(t0);
    k_dfma_peak<<<blocks, threads>>>(d, iters);
    cudaEventRecord(t1);
    e = cudaEventSynchronize(t1);
    if (e != cudaSuccess) break;
    float ms = 0;
    cudaEventElapsedTime(&ms, t0, t1);
    const double flops = 2.0 * 8.0 * (double)iters * blocks * threads;
    best = fmax(best, flops / (ms * 1e-3) / 1e12);
  }
  cudaEventDestroy(t0);
  cudaEventDestroy(t1);
  cudaFree(d);
  *tflops = best;
  return e;
}

// ---- unit-test hook for grb_fastmath.cuh (grb_test_fastmath) ----------------------------------------------------
__global__ void k_fastmath(int kind, int64_t n, const double *__restrict__ x, double *__restrict__ y) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    switch (kind) {
      case 0: y[i] = fast_log(x[i]); break;
      case 1: y[i] = fast_exp(x[i]); break;
      case 2: y[i] = sqrt_fast_ok(x[i]) ? sqrt_fast_core(x[i]) : sqrt(x[i]); break;   // must equal sqrt(x) bit for bit
      case 3: y[i] = sqrt(x[i]); break;
      case 4: y[i] = div_checked(x[i], x[(i + 1) % n]); break;                         // must equal x[i] / x[i+1]
      case 8: case 9: case 10: case 11: {
        // SolvePz: checked reference-order form (8: A, 10: B) vs the power-of-two-folded fast form (9: A, 11: B) on
        // arguments derived from six consecutive inputs (loop-node magnitudes at sqrt(s) = 13 TeV); must agree bit for bit
        const double u0 = x[i], u1 = x[(i + 1) % n], u2 = x[(i + 2) % n], u3 = x[(i + 3) % n], u4 = x[(i + 4) % n], u5 = x[(i + 5) % n];
        PzConst pc; pc.t7 = 13000.0; pc.s = 13000.0 * 13000.0; pc.t9 = pow(pc.s, 1.5); pc.t14 = pc.s * pc.s;
        const double m3 = 0.938272081 + 1e-9 * (u0 - 0.5), m4 = (kind < 10) ? m3 : 0.938272081 + 3.0 * u1 * u1;
        const double pt3 = 3.0 * u2 * u2 + 1e-6, pt4 = 3.0 * u3 * u3 + 1e-6;
        const double mt = 0.3 + 6.0 * u4, yy = 6.0 * (u5 - 0.5), pz5 = mt * sinh(yy), E5 = mt * cosh(yy);
        bool dom = true;
        double r;
        if (kind == 8) r = solve_pz3_A(m3, pt3, pt4, pz5, E5, pc);
        else if (kind == 10) r = solve_pz3_B(m3, m4, pt3, pt4, pz5, E5, pc);
        else {
          const PzInv I = make_pz_inv(m3, m4, pc, kind == 9);
          r = (kind == 9) ? solve_pz3_A_fast(I, pt3, pt4, pz5, E5, dom) : solve_pz3_B_fast(I, pt3, pt4, pz5, E5, dom);
          if (!dom) r = (kind == 9) ? solve_pz3_A(m3, pt3, pt4, pz5, E5, pc) : solve_pz3_B(m3, m4, pt3, pt4, pz5, E5, pc);
        }
        y[i] = r;
        break;
      }
      case 6: { double sn, cs; fast_sincos(x[i], sn, cs); y[i] = sn; break; }
      case 7: { double sn, cs; fast_sincos(x[i], sn, cs); y[i] = cs; break; }
      default: y[i] = x[i] / x[(i + 1) % n]; break;
    }
}
cudaError_t run_fastmath(int kind, int64_t n, const double *x_host, double *y_host) {
  double *d = nullptr;
  cudaError_t e = cudaMalloc(&d, sizeof(double) * 2 * (size_t)n);
  if (e != cudaSuccess) return e;
  e = cudaMemcpy(d, x_host, sizeof(double) * n, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) {
    k_fastmath<<<592, 256>>>(kind, n, d, d + n);
    e = cudaMemcpy(y_host, d + n, sizeof(double) * n, cudaMemcpyDeviceToHost);
  }
  cudaFree(d);
  return e;
}

}  // namespace grb
