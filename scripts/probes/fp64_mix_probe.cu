// fp64_mix_probe.cu -- what fp64 issue rate can an sm_100a SM sustain for instruction mixes like k_screen's?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o /tmp/fp64_mix_probe scripts/probes/fp64_mix_probe.cu
// Variants: pure DFMA (the roofline denominator), DADD/DMUL/DFMA round robin on 8 chains with distinct register
// operands, the same with one integer instruction per fp64 instruction, and with dependent chains only (ILP 1..8).
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE, int ILP>
__global__ void __launch_bounds__(128) k(double *out, int iters, double b0, double c0, int salt) {
  double a[8], b[8], c[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { a[i] = threadIdx.x * 1e-3 + i; b[i] = b0 + 1e-9 * i; c[i] = c0 + 1e-10 * i; }
  int z = salt + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) {
      if (MODE == 0) a[i] = fma(a[i], b[i], c[i]);
      if (MODE >= 1) {
        a[i] = fma(a[i], b[i], c[i]);
        if (MODE == 2) z = (z ^ (z >> 3)) + 0x9e37;
        a[i] = a[i] * b[(i + 1) & 7];
        if (MODE == 2) z = (z ^ (z << 5)) + i;
        a[i] = a[i] + c[(i + 3) & 7];
        if (MODE == 2) z = (z * 3) ^ 0x55;
      }
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += a[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s + z;
}

template <int MODE, int ILP>
void run(const char *name, int ctas_per_sm) {
  int sm = 148;
  cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, 0);
  const int blocks = sm * ctas_per_sm, threads = 128, iters = 1 << 14;
  double *d;
  cudaMalloc(&d, sizeof(double) * blocks * threads);
  cudaEvent_t t0, t1;
  cudaEventCreate(&t0); cudaEventCreate(&t1);
  k<MODE, ILP><<<blocks, threads>>>(d, 64, 1.0000001, 1e-9, 1);
  float best = 1e30f;
  for (int r = 0; r < 3; ++r) {
    cudaEventRecord(t0);
    k<MODE, ILP><<<blocks, threads>>>(d, iters, 1.0000001, 1e-9, r);
    cudaEventRecord(t1); cudaEventSynchronize(t1);
    float ms; cudaEventElapsedTime(&ms, t0, t1);
    if (ms < best) best = ms;
  }
  const double per = (MODE == 0 ? 1.0 : 3.0) * ILP;
  const double inst = per * iters * (double)blocks * threads;   // thread-level fp64 instructions
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("%-34s warps/SM %2d  %.3f ms  %.2f fp64 thread-inst/clk/SM (peak 64)\n", name, ctas_per_sm * 4, best,
         inst / (best * 1e-3) / (clk * 1e3) / sm);
  cudaFree(d);
}

int main() {
  run<0, 8>("DFMA x8 chains", 4);
  run<1, 8>("DFMA/DMUL/DADD x8 chains", 4);
  run<2, 8>("same + 1 int op per fp64 op", 4);
  run<1, 1>("DFMA/DMUL/DADD dependent, ILP1", 4);
  run<1, 2>("DFMA/DMUL/DADD ILP2", 4);
  run<1, 1>("DFMA/DMUL/DADD dependent, ILP1", 3);
  run<1, 2>("DFMA/DMUL/DADD ILP2", 3);
  run<2, 2>("ILP2 + int ops", 4);
  run<2, 1>("ILP1 + int ops", 4);
  run<2, 1>("ILP1 + int ops", 8);
  run<1, 1>("ILP1", 8);
  run<1, 1>("ILP1", 16);
  return 0;
}
