// grb_kernels_points.cuh -- the two hot kernels of the event-weight path and their launchers.
//
//   k_points : one thread per phase-space point. Draws (caller-supplied SoA or Philox through the VEGAS
//              grid) -> <F> kinematics -> cuts -> Born amplitude -> weight prefactor.  Unscreened
//              processes finish here (weight + VEGAS accumulation).  Screened processes append the
//              accepted points to a compact SoA work list (warp-aggregated atomics).
//   k_screen : one thread per accepted point, the eikonal screening loop integral of
//              MProcess::S3ScreenedAmp2 (reference src/MProcess.cc:157-288) with every loop invariant in
//              constant memory; uniform trip count, so no divergence after the compaction.
//
// VEGAS accumulators (reference src/MGraniitti.cc:1389-1419, a mutex per point there): per-thread
// running sums -> warp shuffle -> one global atomic per block; the [bins][dim] histograms live in shared
// memory per block and are flushed once at block exit.
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>

#include "grb_accum.cuh"

namespace grb {

template <int NN>
__device__ __forceinline__ void store_momenta(double *dst, int64_t stride, const PointState<NN> &S, int n_nodes) {
  const v4 *three[3] = {&S.pX, &S.p1, &S.p2};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    dst[(4 * k + 0) * stride] = three[k]->x; dst[(4 * k + 1) * stride] = three[k]->y;
    dst[(4 * k + 2) * stride] = three[k]->z; dst[(4 * k + 3) * stride] = three[k]->e;
  }
#pragma unroll
  for (int i = 0; i < NN; ++i)
    if (i < n_nodes) {
      dst[(12 + 4 * i + 0) * stride] = S.np4[i].x; dst[(12 + 4 * i + 1) * stride] = S.np4[i].y;
      dst[(12 + 4 * i + 2) * stride] = S.np4[i].z; dst[(12 + 4 * i + 3) * stride] = S.np4[i].e;
    }
}

// launch shape per amplitude kind: CTA size and resident CTAs per SM (register cap = 65536 / (threads x CTAs)).  Measured
// per 2^23 points (profiles/r01_variants.txt): Regge 2.66 -> 1.96 ms and tensor 2xS 6.15 -> 4.60 ms at 256 x 2 (128
// registers); resonances 36.8 -> 30.4 ms at 128 x 3 (168 registers; 256 x 2 is slower); HELAS yy -> l+ l- is fastest with
// the full register file (45.5 ms; 48.8 at 128 x 3, 49.6 at 256 x 2)
// Round 2 (profiles/r02_kp_threads_sweep.txt): HELAS at 64 x 4 (same 255 registers, shorter waits at the final barrier:
// 2.02e8 -> 2.16e8 points/s), tensor 2xS at 128 x 4 (1.77e9 -> 1.90e9)
#ifndef GRB_KP_RES_THREADS   // (build-time knobs of the launch-shape sweeps: -DGRB_KP_RES_THREADS=.. -DGRB_KP_RES_MINB=..)
#define GRB_KP_RES_THREADS 128
#endif
#ifndef GRB_KP_RES_MINB
#define GRB_KP_RES_MINB 3
#endif
__host__ __device__ constexpr int kp_threads(int amp) {
  return (amp == AMP_REGGE_RES || amp == AMP_REGGE_RES_CON) ? GRB_KP_RES_THREADS : (amp == AMP_TENSOR_2XS ? 128 : (amp == AMP_YY_LL ? 64 : 256));
}
__host__ __device__ constexpr int kp_minb(int amp) {
  return amp == AMP_REGGE ? 2 : ((amp == AMP_REGGE_RES || amp == AMP_REGGE_RES_CON) ? GRB_KP_RES_MINB : ((amp == AMP_TENSOR_2XS || amp == AMP_YY_LL) ? 4 : 1));
}

// ---- Born amplitude dispatch (reference seam B, MSubProc::GetBareAmplitude2) -----------------------------
template <int NN, int AMP>
__device__ __forceinline__ cx born_amplitude(const PointState<NN> &S) {
  NodeKin in;
  in.ss13 = S.ss13; in.ss24 = S.ss24; in.ss14 = S.ss14; in.ss23 = S.ss23;
  in.t1 = S.t1; in.t2 = S.t2; in.t_hat = S.t_hat; in.u_hat = S.u_hat;
  in.p1 = S.p1; in.p2 = S.p2; in.pX = S.pX; in.exc = S.exc;
  return amplitude<AMP>(in, S.np4[0], S.np4[1]);
}

// weight prefactor: everything of MFactorized.cc:173-174 except |A|^2
template <int NN>
__device__ __forceinline__ double weight_prefactor(const PointState<NN> &S) {
  const double J   = 1.0 / fabs(S.p1.z / S.p1.e - S.p2.z / S.p2.e);                      // :441-443
  const double pt1 = msqrt(pt2_of(S.p1)), pt2 = msqrt(pt2_of(S.p2));
  const double psw = (pt1 / S.p1.e) * (pt2 / S.p2.e) * J;                               // x 1/2 (2pi)^-5 / 4
  const double vol = (pow2(S.M_MAX) - pow2(S.M_MIN)) * (pt1 * pt2);                     // x (2pi)^2 dlog^2 dY
  const double C   = c_proc.ps_active ? S.DW / (2 * GRB_PI) : 1.0;
  double       fwd = 1.0;  // ForwardVolume with excitation: log-change of variable jacobian M^2 (ln M2max - ln M2min) per leg
  if (S.exc & 1) fwd *= m2_of(S.p1) * c_proc.dlog_M2f;
  if (S.exc & 2) fwd *= m2_of(S.p2) * c_proc.dlog_M2f;
  return C * psw * vol * fwd * c_proc.w_const;
}
// <C> class: MContinuum.cc:181-183 with BNPhaseSpaceWeight (:639-658) and BNIntegralVolume (:619-637), Kf = 2
template <int NN>
__device__ __forceinline__ double weight_prefactor_C(const PointState<NN> &S) {
  const double J   = 1.0 / fabs(S.p1.z / S.p1.e - S.p2.z / S.p2.e);
  const double pt1 = msqrt(pt2_of(S.p1)), pt2 = msqrt(pt2_of(S.p2));
  const double psw = (pt1 / S.p1.e) * (pt2 / S.p2.e) * J * S.qprod;                          // x 2^-2Kf (2pi)^-(3Nf-4) / 4
  double       fwd = 1.0;  // ForwardVolume with excitation (as in weight_prefactor)
  if (S.exc & 1) fwd *= m2_of(S.p1) * c_proc.dlog_M2f;
  if (S.exc & 2) fwd *= m2_of(S.p2) * c_proc.dlog_M2f;
  return psw * (pt1 * pt2) * fwd * c_proc.w_const_C;                                         // x (dkt 2pi)^(Kf-1) drap^Kf (2pi)^2 dlog^2
}

template <int NN, bool VEGAS, int AMP>
__global__ void __launch_bounds__(kp_threads(AMP), kp_minb(AMP)) k_points(const PointArgs a) {
  const DevProc &P   = c_proc;
  const int      dim = P.lipsdim;
  const int      nb  = a.bins * dim;
  if (VEGAS) hist_clear(nb);
  ThreadSums ts;
  // Warps take tiles of 32 consecutive points from a global counter (a.tile_count, zeroed per launch) instead of a static
  // stride: the per-point cost varies (rejections, Breit-Wigner retries, light-cone iterations), and with a static split the
  // fast warps of a CTA wait at the final barrier for the slowest one (18-25 % of all stall samples for the heavy Born
  // amplitudes, profiles/r02_k_points_c1_opcode_mix.txt); now all warps run out of work within one tile of each other.
  const int     lane_id = threadIdx.x & 31;
  const int64_t ntiles  = (a.n + 31) >> 5;
  for (;;) {
    unsigned tile = 0;
    if (lane_id == 0) tile = atomicAdd(a.tile_count, 1u);
    tile = __shfl_sync(0xffffffffu, tile, 0);
    if ((int64_t)tile >= ntiles) break;
    const int64_t idx    = ((int64_t)tile << 5) + lane_id;
    const bool    active = idx < a.n;
    constexpr int KL = (AMP == AMP_REGGE_ME8) ? 6 : 4;     // legs of the ladder amplitudes
    constexpr bool LADDER = (AMP == AMP_REGGE_ME6 || AMP == AMP_REGGE_ME8);
    constexpr int UD = LADDER ? 3 * KL + 4 : 10;           // coordinates: <F> 6 (+2 N*), <C> 8 (two legs) / 14 (four) / 20 (six) (+2 N*)
    double        u[UD] = {};
    uint64_t      packed = 0, packed_hi = 0, packed_x = 0;  // VEGAS bin per dimension, 8 bits each
    Ladder<KL>    X4;                                      // ladder scalars (AMP_REGGE_ME6 / ME8 only)
    double        vw     = 1.0;
    PointState<NN> S;
    bool kin = false, fid = false, veto = false, overrun = false;
    if (active) {
      if (VEGAS) {
        DrawsPhilox dr;
        const uint64_t gp = (uint64_t)(a.first_call + idx);
        dr.g.k0 = (uint32_t)a.seed; dr.g.k1 = (uint32_t)(a.seed >> 32);
        dr.p_lo = (uint32_t)gp; dr.p_hi = (uint32_t)(gp >> 32); dr.iter = a.iter;
        dr.pos = 0; dr.overrun = false;
        // reference src/MGraniitti.cc:1357-1375: bin pick + in-bin offset, vegasweight = prod zo*BINS
#pragma unroll
        for (int j = 0; j < UD; ++j)
          if (j < dim) {
            const double zn  = dr.next() * a.bins + 1.0;
            unsigned     ind = (unsigned)zn;
            ind = max(1u, min(ind, (unsigned)a.bins));
            double zo, ac;
            if (ind > 1) {
              const double x1 = a.xmat[(ind - 1) * dim + j], x0 = a.xmat[(ind - 2) * dim + j];
              zo = x1 - x0;
              ac = x0 + (zn - ind) * zo;
            } else {
              zo = a.xmat[(ind - 1) * dim + j];
              ac = (zn - ind) * zo;
            }
            u[j] = 0.0 + ac * 1.0;
            vw *= zo * a.bins;
            if (j < 8) packed |= (uint64_t)(ind - 1) << (8 * j);
            else if (j < 16) packed_hi |= (uint64_t)(ind - 1) << (8 * (j - 8));
            else packed_x |= (uint64_t)(ind - 1) << (8 * (j - 16));
          }
        if (LADDER) kin = point_kinematics_CK<NN, KL>(u, dr, S, X4);
        else kin = P.ps_class == GRB_CLASS_C ? point_kinematics_C<NN>(u, dr, S) : point_kinematics<NN>(u, dr, S);
        overrun = dr.overrun;
      } else {
        DrawsSoA dr;
        dr.u = a.u; dr.tape = a.tape; dr.n = a.ld; dr.idx = idx; dr.dtape = a.dtape; dr.pos = 0; dr.overrun = false;
        dr.rng = a.tape_rng != 0;
        if (dr.rng) {
          const uint64_t gp = (uint64_t)(a.first_call + idx);
          dr.g.k0 = (uint32_t)a.seed; dr.g.k1 = (uint32_t)(a.seed >> 32);
          dr.p_lo = (uint32_t)gp; dr.p_hi = (uint32_t)(gp >> 32); dr.iter = a.iter; dr.d0 = dim;
        }
#pragma unroll
        for (int j = 0; j < UD; ++j)
          if (j < dim) u[j] = dr.coord(j);
        if (LADDER) kin = point_kinematics_CK<NN, KL>(u, dr, S, X4);
        else kin = P.ps_class == GRB_CLASS_C ? point_kinematics_C<NN>(u, dr, S) : point_kinematics<NN>(u, dr, S);
        overrun = dr.overrun;
      }
      if (kin) {  // on failure the reference evaluates the cuts on stale state: flags are don't-care
        fid  = fiducial_cuts<NN>(S);
        veto = veto_cuts<NN>(S);
      }
    }
    __syncwarp();   // (the amplitude starts with the lanes of the warp together, whatever the kinematics and the cuts did)
    const bool valid = active && kin && fid && veto && !overrun;
    double     W = 0.0;
    bool       amp_ok = true;
    cx         h0 = {0.0, 0.0};
    double     pref = 0.0;
    if (valid) {
      pref = P.ps_class == GRB_CLASS_C ? weight_prefactor_C<NN>(S) : weight_prefactor<NN>(S);
      if (LADDER) h0 = regge_ladder<KL>(S.t1, S.t2, X4);
      else h0 = born_amplitude<NN, AMP>(S);
      if (!P.screening) {
        W = pref * (h0.re * h0.re + h0.im * h0.im);
        if (isnan(W) || isinf(W)) { W = 0.0; amp_ok = false; }
      }
    }
    // screened: append to the work list (warp-aggregated slot reservation).  Two classes, filled from the two
    // ends of the buffer: equal-mass points (SolvePz3_A in the loop) from slot 0 up, the others from the top down.
    if (P.screening) {
      LoopInv L;
      bool    same = false;
      if (valid) {
        make_loop_inv<NN>(S, L);
        same = L.same_mass;
      }
      const unsigned mA = __ballot_sync(0xffffffffu, valid && same), mB = __ballot_sync(0xffffffffu, valid && !same);
      if (mA | mB) {
        const int lane = threadIdx.x & 31;
        unsigned int a0 = 0, b0 = 0;
        if (lane == 0) {
          if (mA) a0 = atomicAdd(a.item_count, (unsigned)__popc(mA));
          if (mB) b0 = atomicAdd(a.item_count + 1, (unsigned)__popc(mB));
        }
        a0 = __shfl_sync(0xffffffffu, a0, 0);
        b0 = __shfl_sync(0xffffffffu, b0, 0);
        if (valid) {
          const unsigned below = (1u << lane) - 1;
          const int64_t slot = same ? (int64_t)a0 + __popc(mA & below) : a.item_cap - 1 - ((int64_t)b0 + __popc(mB & below));
          if (slot >= 0 && slot < a.item_cap) {
            double *it = a.items + slot;
            const int64_t c = a.item_cap;
            it[IT_P1X * c] = L.p1x; it[IT_P1Y * c] = L.p1y; it[IT_P2X * c] = L.p2x; it[IT_P2Y * c] = L.p2y;
            it[IT_M1 * c] = L.m1; it[IT_M2 * c] = L.m2; it[IT_MXSQ * c] = L.mXsq;
            it[IT_SHY * c] = L.shY; it[IT_CHY * c] = L.chY;
            if (P.excitation) it[IT_EXC * c] = (double)L.exc;
            if (LADDER) {  // ladders: leg invariants re-derived from the saved 4-vectors (MContinuum.cc:108-112)
#pragma unroll
              for (int k = 0; k < KL - 1; ++k) { it[ladder_row(2 * k) * c] = X4.qx[k]; it[ladder_row(2 * k + 1) * c] = X4.qy[k]; }
#pragma unroll
              for (int k = 0; k < KL; ++k) {
                const v4    &pk = S.np4[k % NN];
                const double yk = rap_of(pk), shk = sinh(yk), chk = cosh(yk);
                it[ladder_row(2 * (KL - 1) + k) * c] = m_of(pk);
                it[ladder_row(2 * (KL - 1) + KL + k) * c] = shk;
                it[ladder_row(2 * (KL - 1) + 2 * KL + k) * c] = chk;
              }
            }
            if (dim > 8) a.item_bins_hi[slot] = packed_hi;   // VEGAS bins of the dimensions 8.. (second packed word)
            if (dim > 16) a.item_bins_hi[c + slot] = packed_x;  // ... and 16.. (third)
            if (P.ps_class == GRB_CLASS_C && !LADDER) {
              it[IT_MXSQ * c] = L.m3; it[IT_SHY * c] = L.m4; it[IT_CHY * c] = L.qx;
              it[(IT_CK + 0) * c] = L.qy; it[(IT_CK + 1) * c] = L.sh3; it[(IT_CK + 2) * c] = L.ch3;
              it[(IT_CK + 3) * c] = L.sh4; it[(IT_CK + 4) * c] = L.ch4;
            }
            if (!LADDER) {  // (the ladders keep their own invariants in these rows)
              it[(IT_P3 + 0) * c] = L.p3.x; it[(IT_P3 + 1) * c] = L.p3.y; it[(IT_P3 + 2) * c] = L.p3.z; it[(IT_P3 + 3) * c] = L.p3.e;
              it[(IT_P4 + 0) * c] = L.p4.x; it[(IT_P4 + 1) * c] = L.p4.y; it[(IT_P4 + 2) * c] = L.p4.z; it[(IT_P4 + 3) * c] = L.p4.e;
            }
            it[IT_H0RE * c] = h0.re; it[IT_H0IM * c] = h0.im;
            it[IT_PREF * c] = pref; it[IT_VW * c] = vw;
            if (AmpHel<AMP>::NH > 1) {
              it[(IT_BK + 0) * c] = S.p1.z; it[(IT_BK + 1) * c] = S.p1.e; it[(IT_BK + 2) * c] = S.p2.z; it[(IT_BK + 3) * c] = S.p2.e;
              it[(IT_BK + 4) * c] = S.pX.x; it[(IT_BK + 5) * c] = S.pX.y; it[(IT_BK + 6) * c] = S.pX.z; it[(IT_BK + 7) * c] = S.pX.e;
            }
            a.item_idx[slot]  = (int32_t)idx;
            a.item_bins[slot] = packed;
            if (VEGAS && a.mode == GRB_MODE_GENERATE) store_momenta<NN>(a.pay + slot, c, S, P.n_nodes);
          }
        }
      }
    }
    if (active) {
      const uint8_t fl = (uint8_t)((kin ? GRB_FLAG_KIN : 0) | (fid ? GRB_FLAG_FID : 0) | (veto ? GRB_FLAG_VETO : 0) |
                                   (amp_ok ? GRB_FLAG_AMP : 0));
      if (a.w) a.w[idx] = W;
      if (a.flags) a.flags[idx] = fl;
      if (a.dump) {
        double *D = a.dump + idx;
        const int64_t n = a.ld;
        D[GRB_DUMP_T1 * n] = kin ? S.t1 : 0; D[GRB_DUMP_T2 * n] = kin ? S.t2 : 0;
        D[GRB_DUMP_S1 * n] = kin ? S.s1 : 0; D[GRB_DUMP_S2 * n] = kin ? S.s2 : 0;
        D[GRB_DUMP_M2 * n] = kin ? S.m2 : 0; D[GRB_DUMP_Y * n] = kin ? S.Y : 0; D[GRB_DUMP_PT * n] = kin ? S.Pt : 0;
        D[GRB_DUMP_THAT * n] = kin ? S.t_hat : 0; D[GRB_DUMP_UHAT * n] = kin ? S.u_hat : 0;
        D[GRB_DUMP_AMP2 * n] = valid ? (h0.re * h0.re + h0.im * h0.im) : 0;
        D[GRB_DUMP_HAMP0_RE * n] = h0.re; D[GRB_DUMP_HAMP0_IM * n] = h0.im;
      }
      // statistics (reference stat.Accumulate(aux)); screened weights are added by k_screen
      ts.n_eval += 1.0;
      ts.n_kin += kin ? 1.0 : 0.0;
      ts.n_fid += (kin && fid) ? 1.0 : 0.0;
      ts.n_veto += (kin && veto) ? 1.0 : 0.0;
      if (!P.screening) {
        ts.n_amp += amp_ok ? 1.0 : 0.0;
        ts.n_naninf += amp_ok ? 0.0 : 1.0;
        if (VEGAS && W != 0.0) {
          const double f = W * vw;
          ts.fsum += f;
          ts.f2sum += f * f;
          ts.maxW = fmax(ts.maxW, W);
          ts.maxf = fmax(ts.maxf, f);
          hist_add(nb, dim, packed, f, packed_hi, packed_x);
          if (a.mode == GRB_MODE_GENERATE) {
            const int64_t es = hitmiss_slot(a, f, (uint64_t)(a.first_call + idx));
            if (es >= 0) {
              store_momenta<NN>(a.ev_out + es, a.ev_cap, S, P.n_nodes);
              a.ev_out[(4 * a.npart + 0) * a.ev_cap + es] = W;
              a.ev_out[(4 * a.npart + 1) * a.ev_cap + es] = f;
            }
          }
        }
      }
    }
    // Reconverge before the next tile.  Without this the lanes that leave the compare-and-swap loops of the shared-memory
    // histogram one after the other stay split, and the warp walks the next point's kinematics and amplitude in several
    // partial passes (ncu: 15.7 active threads per instruction for HELAS gamma gamma -> l+ l- where 24 are accepted;
    // profiles/r02_kp_probe2.txt: 2.6e8 -> 6.2e8 points/s for that workload, 3.8e8 -> 5.8e8 for the resonances).
    __syncwarp();
  }
  flush_sums(ts, a.acc);
  if (VEGAS) hist_flush(nb, dim, a.acc);
}

// points that were not Valid() count as amplitude_ok in the reference (aux default true): added on the host.

template <int NN, int AMP>
inline void launch_points_t(const PointArgs &a, unsigned blocks, int threads, size_t smem, cudaStream_t st) {
  int th = kp_threads(AMP);
  // GRB_KP_THREADS (tuning knob): CTA size of k_points.  The kernel ends with one __syncthreads() (histogram flush); with
  // few, long-running warps per CTA the fast warps wait there for the slowest one
  static const int th_env = [] { const char *e = getenv("GRB_KP_THREADS"); return e ? atoi(e) : 0; }();
  if (th_env >= 32 && th_env <= th && (th_env & 31) == 0) th = th_env;
  if (a.side_launch) th = 128;                                       // (blocks = one CTA per SM, set by the caller)
  else {
    // persistent grid: the CTAs that are resident at once, each taking tiles until the counter runs out
    int occ = 0;
    const cudaError_t rc = a.vegas ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_points<NN, true, AMP>, th, smem)
                                   : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_points<NN, false, AMP>, th, 0);
    const unsigned sms = a.sm_count > 0 ? (unsigned)a.sm_count : 1u;
    const int64_t  want = (a.n + th - 1) / th;
    if (rc == cudaSuccess && occ > 0) blocks = (unsigned)(want < (int64_t)sms * occ ? want : (int64_t)sms * occ);
    else if (th != threads) blocks = blocks * (unsigned)(threads / th);
  }
  if (a.vegas) k_points<NN, true, AMP><<<blocks, th, smem, st>>>(a);
  else k_points<NN, false, AMP><<<blocks, th, 0, st>>>(a);
}
template <int AMP>
inline void launch_points_n(const PointArgs &a, int n_nodes, unsigned blocks, int threads, size_t smem, cudaStream_t st) {
  if (n_nodes <= 2) launch_points_t<2, AMP>(a, blocks, threads, smem, st);
  else if (n_nodes <= 6) launch_points_t<6, AMP>(a, blocks, threads, smem, st);  // config 5: rho rho -> 4 pi; six-leg ladder
  else launch_points_t<GRB_MAX_NODES, AMP>(a, blocks, threads, smem, st);
}


// one explicit instantiation per amplitude kind lives in its own translation unit (grb_points_*.cu)
template <int AMP>
void launch_points_amp(const PointArgs &a, int n_nodes, unsigned blocks, int threads, size_t smem, cudaStream_t st) {
  launch_points_n<AMP>(a, n_nodes, blocks, threads, smem, st);
}

}  // namespace grb
