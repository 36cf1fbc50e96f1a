// tests/host_tres.cc -- TEST INFRASTRUCTURE.  The product's device header graniitti_b200/csrc/grb_tres.cuh compiled by the
// HOST compiler (it is plain C++ outside nvcc), so that the closed-form tensor-pomeron resonance algebra the CUDA kernels run
// can be pinned against the CPU oracle's explicit-tensor restatement without a GPU, in double and in long double (the latter
// measures the rounding noise of the reference's explicit rank-6 contraction).  Built by tests/test_tensor_res.py with
//   g++ -O2 -ffp-contract=off -shared -fPIC tests/host_tres.cc -o tests/_host_tres.so
// Never linked into, or loaded by, the product.
#include "../graniitti_b200/csrc/grb_tres.cuh"

using namespace grb::tres;

template <typename R>
static void run(const ResPar *res, int n_res, const double *m, double *re, double *im) {
  V4<R> v[7];
  for (int i = 0; i < 7; ++i) { v[i].x = m[4 * i]; v[i].y = m[4 * i + 1]; v[i].z = m[4 * i + 2]; v[i].e = m[4 * i + 3]; }
  R a, b;
  resonance_sum<R>(res, n_res, v[0], v[1], v[2], v[3], v[4], v[5], v[6], a, b);
  *re = (double)a;
  *im = (double)b;
}

extern "C" {
int tres_sizeof_respar(void) { return (int)sizeof(ResPar); }
// mom: n points x 28 doubles (pa, pb, p1, p2, pX, p3, p4 as x, y, z, e); out: n x 2
void tres_eval(const ResPar *res, int n_res, long n, const double *mom, int stride, double *out, int extended) {
  for (long k = 0; k < n; ++k) {
    if (extended) run<long double>(res, n_res, mom + k * stride, out + 2 * k, out + 2 * k + 1);
    else run<double>(res, n_res, mom + k * stride, out + 2 * k, out + 2 * k + 1);
  }
}
}
