// grb_kernels_screen.cu -- the screening loop kernels (their own translation unit: the loop kernel dominates both the
// run time and the compile time).
//   k_screen : one thread per accepted point, the eikonal screening loop integral of MProcess::S3ScreenedAmp2
//              (reference src/MProcess.cc:157-288) with every loop invariant in constant memory; uniform trip
//              count, so no divergence after the compaction.
//   k_finish : weight assembly, outputs and VEGAS accumulation for the screened work list.
#include <cuda_runtime.h>

#include <cstdlib>

#include "grb_accum.cuh"

namespace grb {

// k_screen is the pure loop integral: it reads one work-list item, writes |A_born + A_loop|^2 back into the item.
// Weight assembly, NaN handling, outputs and VEGAS accumulation happen in k_finish, which keeps the hot kernel's
// register budget free of histogram/atomic code.
// FF >= 0: single-trajectory Regge continuum with PARAM_REGGE::offshellFF = FF, the one-basic-block node (grb_point.cuh)
template <int AMP, bool SAME, int FF>
__device__ __forceinline__ void screen_item(const PointArgs &a, int64_t s) {
  const DevProc &P = c_proc;
  const int64_t  c = a.item_cap;
  double *       it = a.items + s;
  if (AMP == AMP_REGGE_ME6 || AMP == AMP_REGGE_ME8) {  // ladders: their own invariants and loop (grb_point.cuh)
    constexpr int K = AMP == AMP_REGGE_ME8 ? 6 : 4;
    LoopInvK<K>   LK;
    LK.p1x = it[IT_P1X * c]; LK.p1y = it[IT_P1Y * c]; LK.p2x = it[IT_P2X * c]; LK.p2y = it[IT_P2Y * c];
    LK.m1 = it[IT_M1 * c]; LK.m2 = it[IT_M2 * c]; LK.m1sq = pow2(LK.m1); LK.m2sq = pow2(LK.m2);
#pragma unroll
    for (int k = 0; k < K - 1; ++k) { LK.qx[k] = it[ladder_row(2 * k) * c]; LK.qy[k] = it[ladder_row(2 * k + 1) * c]; }
#pragma unroll
    for (int k = 0; k < K; ++k) {
      LK.m[k]  = it[ladder_row(2 * (K - 1) + k) * c];
      LK.sh[k] = it[ladder_row(2 * (K - 1) + K + k) * c];
      LK.ch[k] = it[ladder_row(2 * (K - 1) + 2 * K + k) * c];
    }
    cx h0;
    h0.re = it[IT_H0RE * c]; h0.im = it[IT_H0IM * c];
    it[IT_AMP2 * c] = screened_amp2_ladder<SAME, K>(LK, h0);
    return;
  }
  LoopInv        L;
  L.p1x = it[IT_P1X * c]; L.p1y = it[IT_P1Y * c]; L.p2x = it[IT_P2X * c]; L.p2y = it[IT_P2Y * c];
  L.m1 = it[IT_M1 * c]; L.m2 = it[IT_M2 * c]; L.mXsq = it[IT_MXSQ * c];
  L.shY = it[IT_SHY * c]; L.chY = it[IT_CHY * c];
  L.m1sq = pow2(L.m1); L.m2sq = pow2(L.m2);
  L.emax = P.sqrt_s - (L.m1 + L.m2);
  L.same_mass = SAME;
  L.exc = P.excitation ? (int)it[IT_EXC * c] : 0;
  if ((FF < 0 || (FF >= 10 && FF < 20)) && P.ps_class == GRB_CLASS_C) {
    L.m3 = L.mXsq; L.m4 = L.shY; L.qx = L.chY;
    L.qy = it[(IT_CK + 0) * c]; L.sh3 = it[(IT_CK + 1) * c]; L.ch3 = it[(IT_CK + 2) * c];
    L.sh4 = it[(IT_CK + 3) * c]; L.ch4 = it[(IT_CK + 4) * c];
  }
  L.p3.x = it[(IT_P3 + 0) * c]; L.p3.y = it[(IT_P3 + 1) * c]; L.p3.z = it[(IT_P3 + 2) * c]; L.p3.e = it[(IT_P3 + 3) * c];
  L.p4.x = it[(IT_P4 + 0) * c]; L.p4.y = it[(IT_P4 + 1) * c]; L.p4.z = it[(IT_P4 + 2) * c]; L.p4.e = it[(IT_P4 + 3) * c];
  if (AmpHel<AMP>::NH > 1) {
    // Born-level kinematics -> the same scalars lorentz_scalars() formed in k_points (same expressions, same bits)
    NodeKin B;
    B.p1.x = L.p1x; B.p1.y = L.p1y; B.p1.z = it[(IT_BK + 0) * c]; B.p1.e = it[(IT_BK + 1) * c];
    B.p2.x = L.p2x; B.p2.y = L.p2y; B.p2.z = it[(IT_BK + 2) * c]; B.p2.e = it[(IT_BK + 3) * c];
    B.exc = L.exc;
    B.pX.x = it[(IT_BK + 4) * c]; B.pX.y = it[(IT_BK + 5) * c]; B.pX.z = it[(IT_BK + 6) * c]; B.pX.e = it[(IT_BK + 7) * c];
    const v4 &p3 = L.p3, &p4 = L.p4;
    B.ss13 = m2_of(B.p1.x + p3.x, B.p1.y + p3.y, B.p1.z + p3.z, B.p1.e + p3.e);
    B.ss14 = m2_of(B.p1.x + p4.x, B.p1.y + p4.y, B.p1.z + p4.z, B.p1.e + p4.e);
    B.ss23 = m2_of(B.p2.x + p3.x, B.p2.y + p3.y, B.p2.z + p3.z, B.p2.e + p3.e);
    B.ss24 = m2_of(B.p2.x + p4.x, B.p2.y + p4.y, B.p2.z + p4.z, B.p2.e + p4.e);
    const double q1x = 0.0 - B.p1.x, q1y = 0.0 - B.p1.y, q1z = P.pb1z - B.p1.z, q1e = P.pb1e - B.p1.e;
    B.t1 = m2_of(q1x, q1y, q1z, q1e);
    B.t2 = m2_of(0.0 - B.p2.x, 0.0 - B.p2.y, P.pb2z - B.p2.z, P.pb2e - B.p2.e);
    B.t_hat = m2_of(q1x - p3.x, q1y - p3.y, q1z - p3.z, q1e - p3.e);
    B.u_hat = m2_of(q1x - p4.x, q1y - p4.y, q1z - p4.z, q1e - p4.e);
    it[IT_AMP2 * c] = screened_amp2_multi<AMP, SAME>(L, B);
  } else {
    cx h0;
    h0.re = it[IT_H0RE * c]; h0.im = it[IT_H0IM * c];
    if (AMP == AMP_REGGE && FF >= 10 && FF < 20) it[IT_AMP2 * c] = screened_amp2_regge_fast_C<SAME, (FF >= 10 && FF < 20 ? FF - 10 : 0)>(L, h0);
    else if (AMP == AMP_REGGE && FF >= 20) it[IT_AMP2 * c] = screened_amp2_regge_fast<SAME, (FF >= 20 ? FF - 20 : 0), true>(L, h0);
    else if (AMP == AMP_REGGE && FF >= 0) it[IT_AMP2 * c] = screened_amp2_regge_fast<SAME, (FF >= 0 && FF < 10 ? FF : 0), false>(L, h0);
    else if (AMP == AMP_TENSOR_2XS && FF >= 0) it[IT_AMP2 * c] = screened_amp2_tensor_fast<SAME>(L, h0);
    else it[IT_AMP2 * c] = screened_amp2_h1<AMP, SAME>(L, h0);
  }
}

template <int MINB, int AMP, int FF>
__global__ void __launch_bounds__(128, MINB) k_screen(const PointArgs a) {
  const int64_t nA = min((int64_t)a.item_count[0], a.item_cap), nB = min((int64_t)a.item_count[1], a.item_cap - nA);
  // class A occupies slots [0, nA), class B slots [cap - nB, cap).  Work is handed out in tiles of 32 consecutive items
  // of one class (one warp, branch-uniform, coalesced) from a global tile counter (item_count[2], zeroed with the list):
  // the persistent grid stays balanced for any list length -- a static stride loop wastes up to one item per thread,
  // 11 % of a launch at 2^20 points -- and the unequal-mass class (more instructions per node) is started first.
  const unsigned tilesA = (unsigned)((nA + 31) >> 5), tilesB = (unsigned)((nB + 31) >> 5);
  const int      lane   = threadIdx.x & 31;
  for (;;) {
    unsigned t = 0;
    if (lane == 0) t = atomicAdd(&a.item_count[2], 1u);
    t = __shfl_sync(0xffffffffu, t, 0);
    if (t >= tilesA + tilesB) break;
    if (t < tilesB) {
      const int64_t i = ((int64_t)t << 5) + lane;
      if (i < nB) screen_item<AMP, false, FF>(a, a.item_cap - nB + i);
    } else {
      const int64_t i = ((int64_t)(t - tilesB) << 5) + lane;
      if (i < nA) screen_item<AMP, true, FF>(a, i);
    }
  }
}

// weight assembly + outputs + VEGAS accumulation for the screened work list
template <bool VEGAS>
__global__ void __launch_bounds__(256) k_finish(const PointArgs a) {
  const int dim = c_proc.lipsdim;
  const int nb  = a.bins * dim;
  if (VEGAS) hist_clear(nb);
  ThreadSums    ts;
  const int64_t c = a.item_cap;
  const int64_t nA = min((int64_t)a.item_count[0], c), nB = min((int64_t)a.item_count[1], c - nA);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < nA + nB; t += stride) {
    const int64_t s  = t < nA ? t : c - nB + (t - nA);
    const double *it = a.items + s;
    const double  amp2 = it[IT_AMP2 * c], pref = it[IT_PREF * c], vw = it[IT_VW * c];
    double        W    = pref * amp2;
    bool          ok   = true;
    if (isnan(W) || isinf(W)) { W = 0.0; ok = false; }
    const int32_t idx = a.item_idx[s];
    if (a.w) a.w[idx] = W;
    if (a.flags && !ok) a.flags[idx] &= (uint8_t)~GRB_FLAG_AMP;
    if (a.dump) a.dump[GRB_DUMP_AMP2 * a.ld + idx] = amp2;
    ts.n_amp += ok ? 1.0 : 0.0;
    ts.n_naninf += ok ? 0.0 : 1.0;
    if (VEGAS && W != 0.0) {
      const double f = W * vw;
      ts.fsum += f;
      ts.f2sum += f * f;
      ts.maxW = fmax(ts.maxW, W);
      ts.maxf = fmax(ts.maxf, f);
      hist_add(nb, dim, a.item_bins[s], f, dim > 8 ? a.item_bins_hi[s] : 0, dim > 16 ? a.item_bins_hi[c + s] : 0);
      if (a.mode == GRB_MODE_GENERATE) {
        const int64_t es = hitmiss_slot(a, f, (uint64_t)(a.first_call + idx));
        if (es >= 0) {
          for (int r = 0; r < 4 * a.npart; ++r) a.ev_out[(int64_t)r * a.ev_cap + es] = a.pay[(int64_t)r * c + s];
          a.ev_out[(4 * a.npart + 0) * a.ev_cap + es] = W;
          a.ev_out[(4 * a.npart + 1) * a.ev_cap + es] = f;
        }
      }
    }
  }
  flush_sums(ts, a.acc);
  if (VEGAS) hist_flush(nb, dim, a.acc);
}

template <int MINB, int AMP, int FF = -1>
static void launch_screen_t(const PointArgs &a, int sm_count, cudaStream_t st) {
  // GRB_SCREEN_OCC=n (profiling knob): pad dynamic shared memory so that only n CTAs fit an SM, same code
  static const int occ = [] { const char *e = getenv("GRB_SCREEN_OCC"); return e ? atoi(e) : 0; }();
  if (occ > 0 && occ < MINB) {
    const int pad = (227 * 1024) / (occ + 1) + 2048;
    cudaFuncSetAttribute(k_screen<MINB, AMP, FF>, cudaFuncAttributeMaxDynamicSharedMemorySize, pad);
    k_screen<MINB, AMP, FF><<<sm_count * occ, 128, pad, st>>>(a);
  } else
    k_screen<MINB, AMP, FF><<<sm_count * ((a.screen_ctas > 0 && a.screen_ctas < MINB) ? a.screen_ctas : MINB), 128, 0, st>>>(a);  // persistent: exactly the resident CTAs
  const size_t smem = a.vegas ? sizeof(double) * 2 * a.bins * a.dim : 0;
  if (a.vegas) k_finish<true><<<sm_count * 4, 256, smem, st>>>(a);
  else k_finish<false><<<sm_count * 4, 256, 0, st>>>(a);
}

// regge_ff: >= 0 when the one-basic-block node applies -- PARAM_REGGE::offshellFF for the Regge continuum with one active
// trajectory (+10 in the <C> class, +20 in the <F> class with s0 = 1 GeV^2), 0 for the tensor pomeron 2xS (<F> class);
// no forward excitation -- else -1
int launch_screen(const PointArgs &a, int amp_kind, int regge_ff, int sm_count, cudaStream_t st) {
  // tuning knob for profiling runs; the default is the measured best (profiles/)
  static const int minb = [] { const char *e = getenv("GRB_SCREEN_MINB"); return e ? atoi(e) : 4; }();
  if (amp_kind == AMP_TENSOR_2XS) {
    const bool nofast_t = getenv("GRB_SCREEN_NOFAST") != nullptr;
    if (regge_ff >= 0 && !nofast_t) launch_screen_t<4, AMP_TENSOR_2XS, 0>(a, sm_count, st);  // one-basic-block node (<F> class)
    else launch_screen_t<4, AMP_TENSOR_2XS>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_TENSOR_RES) {
    launch_screen_t<2, AMP_TENSOR_RES>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_TENSOR_RES_CON) {
    launch_screen_t<2, AMP_TENSOR_RES_CON>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_REGGE_RES) {
    launch_screen_t<2, AMP_REGGE_RES>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_TENSOR_2XV) {
    launch_screen_t<2, AMP_TENSOR_2XV>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_TENSOR_2XF) {
    launch_screen_t<2, AMP_TENSOR_2XF>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_YY_QED) {
    launch_screen_t<2, AMP_YY_QED>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_REGGE_ME6) {
    launch_screen_t<2, AMP_REGGE_ME6>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_REGGE_ME8) {
    launch_screen_t<2, AMP_REGGE_ME8>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_YY_LL) {
    launch_screen_t<2, AMP_YY_LL>(a, sm_count, st);
    return 2;
  }
  if (amp_kind == AMP_REGGE_RES_CON) {
    launch_screen_t<2, AMP_REGGE_RES_CON>(a, sm_count, st);
    return 2;
  }
  const bool nofast = getenv("GRB_SCREEN_NOFAST") != nullptr;  // profiling / test knob (read per launch): the checked node everywhere
  if (!nofast && minb == 4) {
    if (regge_ff == 0) { launch_screen_t<4, AMP_REGGE, 0>(a, sm_count, st); return 2; }
    if (regge_ff == 1) { launch_screen_t<4, AMP_REGGE, 1>(a, sm_count, st); return 2; }
    if (regge_ff == 2) { launch_screen_t<4, AMP_REGGE, 2>(a, sm_count, st); return 2; }
    if (regge_ff == 20) { launch_screen_t<4, AMP_REGGE, 20>(a, sm_count, st); return 2; }  // <F> class, s0 = 1 GeV^2
    if (regge_ff == 21) { launch_screen_t<4, AMP_REGGE, 21>(a, sm_count, st); return 2; }
    if (regge_ff == 22) { launch_screen_t<4, AMP_REGGE, 22>(a, sm_count, st); return 2; }
    if (regge_ff == 10) { launch_screen_t<4, AMP_REGGE, 10>(a, sm_count, st); return 2; }  // <C> class
    if (regge_ff == 11) { launch_screen_t<4, AMP_REGGE, 11>(a, sm_count, st); return 2; }
    if (regge_ff == 12) { launch_screen_t<4, AMP_REGGE, 12>(a, sm_count, st); return 2; }
  }
  if (!nofast && (regge_ff == 2 || regge_ff == 22) && minb == 3) { launch_screen_t<3, AMP_REGGE, 2>(a, sm_count, st); return 2; }  // occupancy probes
  if (!nofast && (regge_ff == 2 || regge_ff == 22) && minb == 5) { launch_screen_t<5, AMP_REGGE, 2>(a, sm_count, st); return 2; }
  switch (minb) {
    case 3: launch_screen_t<3, AMP_REGGE>(a, sm_count, st); break;
    case 5: launch_screen_t<5, AMP_REGGE>(a, sm_count, st); break;
    case 6: launch_screen_t<6, AMP_REGGE>(a, sm_count, st); break;
    default: launch_screen_t<4, AMP_REGGE>(a, sm_count, st); break;
  }
  return 2;
}

}  // namespace grb
