// grb_accum.cuh -- work-list layout, block-level VEGAS accumulation and the hit-or-miss test shared by the point
// kernels (grb_kernels.cu) and the screening kernels (grb_kernels_screen.cu).
#pragma once
#include "grb_point.cuh"

namespace grb {

enum { IT_P1X = 0, IT_P1Y, IT_P2X, IT_P2Y, IT_M1, IT_M2, IT_MXSQ, IT_SHY, IT_CHY, IT_P3 = 9, IT_P4 = 13,
       IT_H0RE = 17, IT_H0IM, IT_PREF, IT_VW, IT_AMP2,
       // Born-level kinematics for the multi-helicity loop: p1z, p1e, p2z, p2e, pX(x,y,z,e)
       IT_BK = 22,
       // <C> class loop invariants: q0y, sinh / cosh of the two leg rapidities (m3, m4, q0x ride in IT_MXSQ, IT_SHY, IT_CHY)
       // the ladders (AMP_REGGE_ME6 / ME8, K = 4 / 6 legs) keep 5K - 2 invariants -- q_i (x, y; K - 1 of them), m[K], sinh y[K],
       // cosh y[K] -- in the rows ladder_row(0 .. 5K - 3): rows 6..16, then 22..34, then 36..39
       IT_CK = 30,
       IT_EXC = 35,  // lts.excite1/2 bits of the point (forward N* excitation)
       IT_NF = 40 };
static_assert(IT_NF == ITEM_ROWS, "the work-list allocation (grb_api.cu) and the row layout disagree");
__host__ __device__ constexpr int ladder_row(int r) { return r < 11 ? 6 + r : (r < 24 ? 22 + (r - 11) : 36 + (r - 24)); }


// ---- block-level VEGAS accumulation ---------------------------------------------------------------------
struct ThreadSums {
  double fsum = 0, f2sum = 0, n_eval = 0, n_kin = 0, n_fid = 0, n_veto = 0, n_amp = 0, n_naninf = 0;
  double maxW = 0, maxf = 0;
};

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ void atomic_max_pos(double *addr, double v) {
  // non-negative doubles order like their bit patterns
  atomicMax(reinterpret_cast<long long *>(addr), __double_as_longlong(v));
}

static __device__ void flush_sums(const ThreadSums &t, grb_accum *acc) {
  double s[8] = {t.fsum, t.f2sum, t.n_eval, t.n_kin, t.n_fid, t.n_veto, t.n_amp, t.n_naninf};
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = warp_sum(s[k]);
  const double mw = warp_max(t.maxW), mf = warp_max(t.maxf);
  if ((threadIdx.x & 31) == 0) {
    double *a = &acc->fsum;  // fsum,f2sum,evaluations,n_kin_ok,n_fid_ok,n_veto_ok,n_amp_ok,n_nan_inf
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (s[k] != 0.0) atomicAdd(a + k, s[k]);
    if (mw > 0) atomic_max_pos(&acc->maxW, mw);
    if (mf > 0) atomic_max_pos(&acc->maxf, mf);
  }
}

// histograms: shared fm[bins*dim], f2m[bins*dim]
extern __shared__ double s_hist[];
__device__ __forceinline__ void hist_clear(int nb) {
  for (int i = threadIdx.x; i < 2 * nb; i += blockDim.x) s_hist[i] = 0.0;
  __syncthreads();
}
__device__ __forceinline__ void hist_add(int nb, int dim, uint64_t packed, double f, uint64_t packed_hi = 0, uint64_t packed_x = 0) {
  const double f2 = f * f;
#pragma unroll 1
  for (int j = 0; j < dim; ++j) {
    const int b = (int)(((j < 8 ? packed : (j < 16 ? packed_hi : packed_x)) >> (8 * (j & 7))) & 0xff);
    atomicAdd(&s_hist[b * dim + j], f);
    atomicAdd(&s_hist[nb + b * dim + j], f2);
  }
}
__device__ __forceinline__ void hist_flush(int nb, int dim, grb_accum *acc) {
  __syncthreads();
  for (int i = threadIdx.x; i < nb; i += blockDim.x) {
    const int    b = i / dim, j = i % dim;
    const double a = s_hist[i], c = s_hist[nb + i];
    if (a != 0.0) atomicAdd(&acc->fmat[b * GRB_MAX_DIM + j], a);
    if (c != 0.0) atomicAdd(&acc->f2mat[b * GRB_MAX_DIM + j], c);
  }
}

// ---- hit-or-miss (reference MGraniitti::SaveEvent accept test, src/MGraniitti.cc:1657-1668) ------------------
// The uniform is draw #63 of the point's Philox stream (the kinematics never reaches it), so the decision is
// a pure function of (seed, iteration, global point index, maxf).
__device__ __forceinline__ double hitmiss_uniform(uint64_t seed, uint32_t iter, uint64_t gp) {
  Philox g;
  g.k0 = (uint32_t)seed; g.k1 = (uint32_t)(seed >> 32);
  uint32_t o[4];
  g.block((uint32_t)gp, (uint32_t)(gp >> 32), iter, 31u, o);
  return u01_from_bits(o[3], o[2]);
}
// returns the event slot (or -1): accepted iff U < f / maxf; f > maxf counts as an overflow (still accepted)
__device__ __forceinline__ int64_t hitmiss_slot(const PointArgs &a, double f, uint64_t gp) {
  if (!(f > 0.0)) return -1;
  if (f > a.maxf) atomicAdd(a.ev_count + 1, 1u);
  if (!(hitmiss_uniform(a.seed, a.iter, gp) < f / a.maxf)) return -1;
  const unsigned int slot = atomicAdd(a.ev_count, 1u);
  return slot < a.ev_cap ? (int64_t)slot : -1;
}
}  // namespace grb
