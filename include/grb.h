/* grb.h -- C ABI of the B200-native batched event-weight engine for GRANIITTI's hot path.
 *
 * The reference (mieskolainen/GRANIITTI v1.051) has no FFI: its "operator API" for this path is C++
 * virtual dispatch across two seams:
 *
 *   seam A  double MProcess::EventWeight(const std::vector<double>& randvec, AuxIntData& aux)
 *           reference include/Graniitti/MProcess.h:84, called per VEGAS/FLAT point from
 *           MGraniitti::VEGASMultiThread (src/MGraniitti.cc:1341-1443) and SampleFlat (:1446-1482)
 *   seam B  double MSubProc::GetBareAmplitude2(gra::LORENTZSCALAR&)   (include/Graniitti/MSubProc.h:605)
 *           selected by the process string "ISTATE[CHANNEL]<class> -> final" (src/MSubProc.cc:90-156)
 *
 * This header is what a cgo/ctypes/C++ binding of those seams binds instead: plain pointers and sizes,
 * no C++ or torch types, errors as negative return codes (never exceptions), all buffers caller-owned.
 * One engine per host thread / GPU rank; an engine is not thread-safe.
 *
 * Every entry point names the reference interface it replaces.  INTEGRATION.md shows the reference-side
 * stub (a patched MGraniitti::VEGAS) that a maintainer would add.
 */
#ifndef GRB_H
#define GRB_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GRB_ABI_VERSION 3

/* ---- error codes ---------------------------------------------------------------------------------- */
enum {
  GRB_OK              = 0,
  GRB_ERR_ARG         = -1, /* bad argument (replaces std::invalid_argument thrown by the reference)    */
  GRB_ERR_CUDA        = -2, /* CUDA runtime error; no CPU fallback exists by design                      */
  GRB_ERR_UNSUPPORTED = -3, /* process / option outside the hot-path scope        */
  GRB_ERR_PARSE       = -4, /* steering card / model file parse error                                    */
  GRB_ERR_STATE       = -5, /* call sequence error (e.g. screening on but no eikonal table set)          */
  GRB_ERR_NCCL        = -6
};

/* ---- limits --------------------------------------------------------------------------------------- */
#define GRB_MAX_NODES 16   /* flattened central decay tree nodes (config 5: 2 rho + 4 pi = 6)            */
#define GRB_MAX_RES 12     /* resonances in a coherent sum (cards use <= 8)                              */
#define GRB_MAX_VETO 8     /* veto domains                                                               */
#define GRB_MAX_DIM 20     /* VEGAS dimension bound (<F>: 6 + NSTARS; <C>: 3 Nf - 4 + NSTARS, Nf - 2 = 6 legs: 20)  */
#define GRB_MAX_TAPE 32    /* internal uniforms per point (config 5 uses 12)                             */
#define GRB_MAX_BINS 128   /* VEGAS bins per dimension (reference MVEGAS.h BINS default 128)             */
#define GRB_MAX_LOOP_KT 64 /* screening loop kt nodes (Nkt+1 <= 46 for ND<=10)                           */
#define GRB_MAX_LOOP_PHI 64
#define GRB_RHO_DIM 5      /* spin density matrices up to J=2                                            */
#define GRB_MAX_HAMP 36    /* helicity amplitudes per point (tensor pomeron 2xV)                         */

/* ---- process ids: the reference's MSubProc registry keys (include/Graniitti/MSubProc.h PROC_n) ---- */
enum {
  GRB_PROC_NONE         = 0,
  GRB_PROC_YY_CON       = 3,  /* "yy[CON]"       PROC_3  MSubProc.h:165  MG5 HELAS yy->l+l- + kt-EPA flux */
  GRB_PROC_YY_QED       = 4,  /* "yy[QED]"       PROC_4  :175            tensor-pomeron code, QED branch   */
  GRB_PROC_PP_RES       = 9,  /* "PP[RES]"       PROC_9                  sum of MRegge::ME3/PhotoME3       */
  GRB_PROC_PP_RESTENSOR = 11, /* "PP[RESTENSOR]" PROC_11 :278            MTensorPomeron::ME3 (J = 0+, 2+ -> two pseudoscalars) */
  GRB_PROC_PP_CONTENSOR = 12, /* "PP[CONTENSOR]" PROC_12 :288            MTensorPomeron::ME4               */
  GRB_PROC_PP_RES_CONTENSOR = 14, /* "PP[RES+CONTENSOR]" PROC_14 :321    ME4 (2xS) + ME3, helicity amplitudes added index by index */
  GRB_PROC_PP_CON       = 15, /* "PP[CON]"       PROC_15 :361            MRegge::ME4(+1)                   */
  GRB_PROC_PP_CON_MINUS = 16, /* "PP[CON-]"      PROC_16 :382            MRegge::ME4(-1)                   */
  GRB_PROC_PP_RES_CON   = 17  /* "PP[RES+CON]"   PROC_17 :399            ME4 + sum of ME3/PhotoME3         */
};

/* phase-space class: "<F>" MFactorized (all BASELINE configs), "<C>" MContinuum (2 -> 4 direct)  */
enum { GRB_CLASS_F = 0, GRB_CLASS_C = 1 };

/* flag bits of one evaluated point = AuxIntData of the reference (include/Graniitti/MProcess.h:37-52) */
enum {
  GRB_FLAG_KIN  = 1, /* kinematics_ok                                                                   */
  GRB_FLAG_FID  = 2, /* fidcuts_ok   (meaningful only if GRB_FLAG_KIN; the reference evaluates it on     */
  GRB_FLAG_VETO = 4, /* vetocuts_ok   stale state otherwise, MFactorized.cc:146-148)                     */
  GRB_FLAG_AMP  = 8  /* amplitude_ok = weight finite (MProcess.h:317-340 CheckInfNan)                    */
};
#define GRB_FLAGS_VALID (GRB_FLAG_KIN | GRB_FLAG_FID | GRB_FLAG_VETO)

/* ---- flattened process description (POD twin of MProcess + LORENTZSCALAR + PARAM_* namespaces) ----- */

/* one node of the flattened central decay tree (reference MDecayBranch, MParticle.h:62-93) */
typedef struct {
  int32_t pdg;
  int32_t spinX2;
  int32_t parent;      /* index of mother node, -1 for first-level legs of the central system           */
  int32_t first_child; /* index of first daughter, -1 if final state                                    */
  int32_t n_children;  /* 0 (final) or 2 (two-body decays only on the device path)                      */
  int32_t pad_;
  double  mass;        /* PDG pole mass                                                                 */
  double  width;       /* PDG width; BW sampled when width >= 1e-40 (MProcess.cc:753-757)               */
} grb_node;

/* reference gra::GENCUT / FIDCUT / VETOCUT (include/Graniitti/MKinematics.h:1554-1625) */
typedef struct {
  double Y_min, Y_max;                   /* GENCUTS.<F>.Rap                                             */
  double M_min, M_max;                   /* GENCUTS.<F>.M                                               */
  double forward_pt_min, forward_pt_max; /* after SetTechnicalBoundaries (MProcess.cc:1369-1386)        */
  double XI_min, XI_max;
  double kt_min, kt_max;                 /* GENCUTS.<C>.Kt (default [0, 10 / n_legs], MContinuum.cc:70-84)  */
  double rap_min, rap_max;               /* GENCUTS.<C>.Rap: rapidity of each central leg                   */
} grb_gencut;

typedef struct {
  int32_t active;
  int32_t usercuts; /* USERCUTS id (MUserCuts.cc:21-215); 0, +-3, +-111 supported                        */
  double  eta_min, eta_max, rap_min, rap_max, pt_min, pt_max, Et_min, Et_max; /* PARTICLE                */
  double  M_min, M_max, Y_min, Y_max, Pt_min, Pt_max;                         /* SYSTEM                  */
  double  forward_t_min, forward_t_max, forward_M_min, forward_M_max;         /* FORWARD                 */
} grb_fidcut;

typedef struct {
  double eta_min, eta_max, pt_min, pt_max;
} grb_vetodomain;

/* PARAM_SOFT (src/MForm.cc:71-130), PARAM_REGGE (src/MRegge.cc:924-1033) */
typedef struct {
  double  DELTA_P, ALPHA_P, gN_P, gN_O, g3P, gamma, fc1, fc2, fc3;
  int32_t odderon_on;
  int32_t f2_model; /* PARAM_STRUCTURE.F2 (proton structure function of the incoherent photon flux): 0 CKMT, 1 DL */
} grb_param_soft;

typedef struct {
  double  a0[3], ap[3], sgn[3];
  double  s0;
  int32_t offshellFF; /* 0 exp, 1 Orear, 2 power law                                                     */
  int32_t reggeize;
  double  b_EXP, a_OREAR, b_OREAR, b_POW, omega;
  double  c[3];       /* couplings of the first central leg (PARAM_PI / PARAM_K / ...)                   */
  int32_t n[3];       /* trajectory on/off                                                               */
  int32_t pad_;
} grb_param_regge;

/* PARAM_TENSORPOM (modeldata GENERAL.json), MTensorPomeron.h */
typedef struct {
  double delta_P, delta_O, delta_1R, delta_2R;
  double ap_P, ap_O, ap_1R, ap_2R, eta_O;
  double gPNN, gPpipi, gPKK;
  double gPrhorho[2], gPphiphi[2];
  double M_O, M_1R;
} grb_param_tensorpom;

/* one resonance (reference gra::PARAM_RES, include/Graniitti/MResonance.h) */
typedef struct {
  int32_t pdg, spinX2, P, BW;
  int32_t spingen, spindec, frame /* 0 CM,1 HX,2 CS,3 GJ,4 PG,5 SR */, jmax;
  double  mass, width;
  double  g_re, g_im; /* production coupling                                                             */
  double  g_FF;
  double  g_decay;    /* hel.g_decay                                                                     */
  double  BR;
  double  g_Tensor[8];       /* tensor-pomeron production couplings (2 for J = 0, 1; 7 for J = 2), PARAM_RES.g_Tensor          */
  double  g_decay_tensor[2]; /* tensor-pomeron decay couplings hel.g_decay_tensor: BRANCHING.json, else MTensorPomeron::GDecay
                              * from J, M, W, BR and the daughter mass (src/MProcess.cc:594-700, src/MTensorPomeron.cc:138-156) */
  double  rho_re[GRB_RHO_DIM * GRB_RHO_DIM]; /* spin density matrix, row-major, (2J+1)^2 used            */
  double  rho_im[GRB_RHO_DIM * GRB_RHO_DIM];
  double  T_re[9], T_im[9];                  /* decay helicity T matrix (hel.T), row-major n1 x n2       */
  int32_t T_rows, T_cols;
} grb_resonance;

typedef struct {
  int32_t abi_version;
  int32_t ps_class;   /* GRB_CLASS_F | GRB_CLASS_C                                                       */
  int32_t proc;       /* GRB_PROC_*                                                                      */
  int32_t screening;  /* POMLOOP                                                                         */
  int32_t excitation; /* NSTARS: 0, or 1 / 2 forward N* systems (PP[CON], PP[CON-], PP[CONTENSOR], yy[CON]; <F> class) */
  int32_t flatamp;    /* @FLATAMP (0 only)                                                               */
  int32_t ps_active;  /* !ISOLATE ("->" true, "&>" false)                                                */
  int32_t flatmass2;
  int32_t lipsdim;    /* VEGAS dimension (MFactorized.cc:69-78)                                          */
  int32_t n_tape;     /* internal uniforms consumed per point when no rejection loop repeats             */
  int32_t n_nodes;    /* decay tree nodes                                                                */
  int32_t n_legs;     /* first-level central legs (nodes[0..n_legs-1])                                   */
  int32_t n_final;    /* leaves                                                                          */
  int32_t n_res;
  int32_t n_veto;
  int32_t veto_active;
  int32_t em_ff;      /* PARAM_STRUCTURE.EM: 0 DIPOLE, 1 KELLY                                           */
  int32_t qed_alpha;  /* PARAM_STRUCTURE.QED_alpha: 0 ZERO, 1 MG                                         */
  int32_t loop_nkt;   /* NumberLoopKT  (MEikonal.h:44-51)                                                */
  int32_t loop_nphi;  /* NumberLoopPHI                                                                   */
  double  loop_min_kt, loop_max_kt;
  double  s, sqrt_s;
  double  pbeam1_pz, pbeam1_E, pbeam2_pz, pbeam2_E;
  double  beam1_mass, beam2_mass;
  double  S_factor;   /* identical-particle symmetry factor (MProcess.cc:449-468)                        */
  double  offshell;   /* OFFSHELL: number of widths in BW sampling (default 5)                           */
  double  flat_b;     /* PARAM_FLAT::b                                                                   */
  grb_node            nodes[GRB_MAX_NODES];
  grb_gencut          gcuts;
  grb_fidcut          fcuts;
  grb_vetodomain      veto[GRB_MAX_VETO];
  grb_param_soft      soft;
  grb_param_regge     regge;
  grb_param_tensorpom tpom;
  grb_resonance       res[GRB_MAX_RES];
} grb_process_desc;

/* VEGAS accumulators of one iteration = what VEGASMultiThread adds under the mutex
 * (src/MGraniitti.cc:1389-1419) + the NaN/Inf counters of MProcess::CheckInfNan.
 * All doubles so that one fp64 sum all-reduce + one max all-reduce covers it.                            */
typedef struct {
  double fsum, f2sum;
  double evaluations;   /* stat.evaluations                                                              */
  double n_kin_ok, n_fid_ok, n_veto_ok, n_amp_ok; /* pass counters (stat.*_ok)                            */
  double n_nan_inf;
  double n_accepted;    /* hit-or-miss accepted (mode generate)                                          */
  double n_overflow;    /* f > maxf during generation (MGraniitti.cc:1660)                                */
  double maxW, maxf;    /* max-reduced                                                                   */
  double fmat[GRB_MAX_BINS * GRB_MAX_DIM];  /* [bin][dim], dim fastest                                    */
  double f2mat[GRB_MAX_BINS * GRB_MAX_DIM];
} grb_accum;

typedef struct grb_engine grb_engine;

/* ---- lifecycle ------------------------------------------------------------------------------------ */

/* Build an engine from a flattened process description. Replaces `new MFactorized(process, syntax)` +
 * `post_Constructor()` (src/MGraniitti.cc:245-309,1016-1018).                                            */
int grb_create(const grb_process_desc *desc, int device, grb_engine **out);

/* Build an engine from the reference's own steering inputs: JSON5 card (e.g. input/test.json), optional process
 * string override (`gr -p`), optional POMLOOP override (`gr -l`; -1 = card), and the GRANIITTI base
 * directory holding modeldata/ (NULL = built-in TUNE0 + PDG subset). Replaces MGraniitti::ReadInput
 * (src/MGraniitti.cc:317-330) + ReadModelParam + post_Constructor.                                       */
int grb_create_from_card(const char *card_path, const char *process_override, int pomloop_override,
                         const char *graniitti_base, int device, grb_engine **out);

/* Host-only variant of the above (no GPU needed): parse and flatten only.                                */
int grb_parse_card(const char *card_path, const char *process_override, int pomloop_override,
                   const char *graniitti_base, grb_process_desc *out, char *errbuf, int errbuf_len);

int         grb_get_desc(const grb_engine *e, grb_process_desc *out);
const char *grb_last_error(const grb_engine *e);
void        grb_destroy(grb_engine *e);

/* ---- eikonal screening table ---------------------------------------------------------------------- */

/* Load the screening amplitude table A_eik(kt^2) (reference MEikonal::MSA, IArray1D, src/MEikonal.cc:
 * 404-512): n+1 points on a uniform kt^2 grid. Replaces MEikonal::S3Constructor's file read
 * (src/MEikonal.cc:46-128).  Host pointers.                                                              */
int grb_set_eikonal(grb_engine *e, const double *kt2, const double *re, const double *im, int n_points);

/* Build the density (MBT) and amplitude (MSA) tables on the device from PARAM_SOFT and install the MSA table,
 * replacing MEikonal::S3Density/S3Screening/S3CalculateArray (src/MEikonal.cc:177-242,379-401; one
 * Fourier-Bessel integral of 10 001 nodes per table row) and S3CalcXS (:262-376, single-channel part).
 * number_bt / number_kt2 <= 0 select the NUMERICS.json defaults (10000).  Optional host outputs: the MBT table
 * (number_bt+1 rows), the MSA table (number_kt2+1 rows), sigma_tot/el/inel in mb.  Values are rounded to the
 * 15 significant digits of the reference's ASCII table cache (:84-91,404-436), as the reference reads them.  */
int grb_build_eikonal(grb_engine *e, int number_bt, int number_kt2, double *bt_out, double *omega_re_out,
                      double *omega_im_out, double *kt2_out, double *re_out, double *im_out,
                      double *xs_tot_el_inel /* [3] or NULL */);

/* Nodes of the screening loop as the kernels execute them: nkt_nodes = NumberLoopKT + 1, nphi_nodes = NumberLoopPHI
 * (phi = 2 pi is folded onto phi = 0 with both Simpson end weights) or NumberLoopPHI + 1 with GRB_NO_PHI_FOLD=1.
 * GRB_ERR_STATE before a table was installed.                                                                          */
int grb_loop_nodes(const grb_engine *e, int *nkt_nodes, int *nphi_nodes);

/* The reference's table cache (src/MEikonal.cc:60-128): file names "MBT_<pdg1>_<pdg2>_<sqrts>_<djb2 hash>" /
 * "MSA_..." (hash over s, beams, PARAM_SOFT, MEikonalNumerics; src/MForm.cc:94-100, MEikonal.h:60-67,
 * MAux.cc:281-287) and the 3-column "%.15g" CSV rows of IArray1D::WriteArray/ReadArray (:404-487), so that
 * tables are interchangeable with the reference's eikonal/ directory.  Host-only.                          */
int grb_eikonal_cache_names(const grb_process_desc *desc, int beam1_pdg, int beam2_pdg, int number_bt,
                            int number_kt2, char *mbt_name, char *msa_name, int cap);
int grb_write_eikonal_file(const char *path, const double *x, const double *re, const double *im, int n_points);
int grb_read_eikonal_file(const char *path, double *x, double *re, double *im, int n_points);

/* ---- seam A x n : batched EventWeight --------------------------------------------------------------- */

/* Evaluate n caller-supplied points. u is structure-of-arrays [dim][n] (VEGAS-visible coordinates, the
 * `randvec` of EventWeight); tape is [dtape][n] (the internal uniforms the reference draws from its
 * thread RNG inside EventWeight: BW masses, decay angles; MFactorized.cc:374, MKinematics.h:405-409).
 * w[n] receives the weight W (0 if not Valid()), flags[n] the GRB_FLAG_* bits. dump (optional, [ndump][n],
 * see GRB_DUMP_*) receives intermediate scalars for parity tests. Pointers may be host or device.
 * Host buffers are pipelined in slices (GRB_HOST_SLICE points, default 2^20): copy-in, kernels and copy-out of
 * consecutive slices overlap on three streams; pinned host memory makes the copies asynchronous.          */
int grb_eval(grb_engine *e, int64_t n, const double *u, int dim, const double *tape, int dtape, double *w,
             uint8_t *flags, double *dump);

/* Seam A exactly as the reference calls it: the caller supplies only the VEGAS-visible coordinates (`randvec` of
 * MProcess::EventWeight, include/Graniitti/MProcess.h:84) and the internal uniforms are drawn by the engine -- the
 * reference draws them from its per-thread generator inside EventWeight (MFactorized.cc:374, MKinematics.h:405-409), here
 * point i takes the draws dim, dim+1, ... of the Philox4x32-10 stream (key = seed, counter = (iteration, first_index + i)),
 * the same stream grb_vegas_iterate uses.  Halves the host-to-device traffic of grb_eval for the BASELINE configurations
 * (48 instead of 80 bytes per point).  Reproducible on the host with tests/philox_ref.py.                                */
int grb_eval_rng(grb_engine *e, int64_t n, const double *u, int dim, uint64_t seed, uint64_t iteration,
                 int64_t first_index, double *w, uint8_t *flags, double *dump);

enum {
  GRB_DUMP_T1 = 0, GRB_DUMP_T2, GRB_DUMP_S1, GRB_DUMP_S2, GRB_DUMP_M2, GRB_DUMP_Y, GRB_DUMP_PT,
  GRB_DUMP_THAT, GRB_DUMP_UHAT, GRB_DUMP_AMP2, GRB_DUMP_HAMP0_RE, GRB_DUMP_HAMP0_IM,
  GRB_DUMP_LEN
};

/* ---- VEGAS iteration on the device (hot loop 1) ------------------------------------------------------ */

enum { GRB_MODE_BURNIN = 0, GRB_MODE_INTEGRATE = 1, GRB_MODE_GENERATE = 2 };

/* One VEGAS iteration of `ncalls` points for this rank: points [first_call, first_call+ncalls) of the
 * iteration are generated on the device by Philox4x32-10 (key = seed, counter = (iteration, global point
 * index, draw index)), mapped through the grid xmat[bins][dim] exactly as VEGASMultiThread does
 * (src/MGraniitti.cc:1357-1375), evaluated, and accumulated into *out (host).  Replaces the thread
 * fan-out + mutex accumulation of src/MGraniitti.cc:1208-1214,1341-1443.  In GRB_MODE_GENERATE accepted
 * events (hit-or-miss vs maxf, :1657-1668) are compacted into the engine's event buffer.                 */
int grb_vegas_iterate(grb_engine *e, const double *xmat, int bins, int dim, int64_t first_call,
                      int64_t ncalls, uint64_t seed, uint64_t iteration, int mode, double maxf,
                      grb_accum *out);

/* Same, but leaves the accumulators on the device and returns the device pointer (layout = grb_accum) so
 * that a collective can run on it without a host round trip.  Completed by grb_vegas_fetch.              */
int grb_vegas_iterate_async(grb_engine *e, const double *xmat, int bins, int dim, int64_t first_call,
                            int64_t ncalls, uint64_t seed, uint64_t iteration, int mode, double maxf,
                            void **accum_device_ptr);
int grb_vegas_fetch(grb_engine *e, grb_accum *out);

/* Grid refinement, reference VEGASData::OptimizeGrid/Rebin (include/Graniitti/MVEGAS.h:87-142). Pure host
 * function, O(bins*dim), run redundantly on every rank from the all-reduced f2mat.                        */
int grb_vegas_optimize_grid(double *xmat, const double *f2mat, int bins, int dim, double lambda);

/* Accepted events of the last GRB_MODE_GENERATE iteration: 4-momenta [n_particles][4][n_events] SoA with
 * n_events = the return value (particle 0 = central system, 1,2 = forward protons = lts.pfinal[0..2], then
 * one per decay-tree node in grb_process_desc.nodes order) and weights [2][n_events] = (W, f = W*vegasweight).
 * Host pointers; returns the number of events copied (<= max_events) or a negative error.  A point is
 * accepted iff U < f/maxf with U = draw #63 of its Philox stream (reference SaveEvent, MGraniitti.cc:
 * 1657-1668); grb_accum.n_accepted / n_overflow report the totals of the iteration.                        */
int64_t grb_get_events(grb_engine *e, int64_t max_events, double *momenta, double *weights,
                       int *n_particles);

/* Accepted events of the last GRB_MODE_GENERATE iteration that did not fit the engine's event buffer (GRB_EVENT_CAP
 * events, default 2^22): grb_accum.n_accepted counts every accepted event, grb_get_events returns at most the capacity. */
int64_t grb_events_dropped(const grb_engine *e);

/* ---- event record + writer ------------------------------------------------------------------ */

/* One particle / vertex of the event record the reference builds in MProcess::CommonRecord (src/MProcess.cc:1588-1706):
 * ids count from 1 in the order HepMC3::GenEvent assigns them (beam 1, outgoing proton 1, propagator 1, beam 2, outgoing
 * proton 2, propagator 2, central system, then the decay tree depth first); vertex ids count down from -1.
 * parent: 0 for the beams, else the (negative) id of the production vertex.                                        */
typedef struct {
  int32_t id, pid, status, parent;
  double  px, py, pz, e, m;
} grb_rec_particle;
typedef struct {
  int32_t id, status, n_in, n_out;
  int32_t in[4];       /* ids of the incoming particles                                                                */
  double  x, y, z, t;  /* position [mm]: decay vertices of particles with a width (exponential lifetime), else 0        */
} grb_rec_vertex;
#define GRB_MAX_REC_PARTICLES (7 + GRB_MAX_NODES)
#define GRB_MAX_REC_VERTICES (4 + GRB_MAX_NODES)

/* The record of ONE event from its momenta [3 + n_nodes][4] = (px, py, pz, E) of the central system, the two outgoing
 * forward systems and every decay-tree node in grb_process_desc.nodes order (the layout of grb_get_events, transposed).
 * beam pdg 0 = 2212.  lifetime_u: one uniform per node for the decay-vertex positions, or NULL (all vertices at 0).
 * Status codes and ids: PDG_BEAM 4, PDG_STABLE 1, PDG_DECAY 2, PDG_INTERMEDIATE 81; propagators 99, system 90, excited
 * forward system 90210 (include/Graniitti/MPDG.h:41-50).  Host-only.                                                */
int grb_event_record(const grb_process_desc *d, const double *momenta, int beam1_pdg, int beam2_pdg,
                     const double *lifetime_u, grb_rec_particle *particles, int *n_particles,
                     grb_rec_vertex *vertices, int *n_vertices);

typedef struct {
  const char *format;            /* "hepmc3" | "hepmc2" | "hepevt" (OUTPUT FORMAT of the steering card)                 */
  const char *generator_name;    /* run header tools (src/MGraniitti.cc:190-200); NULL = "GRANIITTI (TUNE0)"           */
  const char *generator_version; /* NULL = "1.051"                                                                      */
  const char *card_path;
  double      sigma_pb, sigma_err_pb; /* per-event GenCrossSection in picobarn (SaveEvent: stat.sigma * 1e12)          */
  int32_t     beam1_pdg, beam2_pdg;   /* 0 = 2212                                                                       */
  int32_t     weighted;          /* WEIGHTED: event weight = W, else 1                                                  */
  int32_t     nthreads;          /* formatting threads, 0 = all host cores                                              */
  int32_t     append;            /* append to an existing listing (no run header)                                       */
  int32_t     close_listing;     /* write the END_EVENT_LISTING footer                                                  */
  int32_t     number_events;     /* 0: event number 0 like the reference (it never sets one); 1: first_event + i       */
  int32_t     lifetimes;         /* 1: decay-vertex positions from exponential lifetimes (counter-based draws)          */
  int64_t     first_event;
  uint64_t    lifetime_seed;
} grb_writer_opts;

/* Record construction on several host threads + one ordered writer for n events given as SoA momenta [3 + n_nodes][4][n]
 * and weights [2][n] = (W, f) (the layout of grb_get_events); replaces the per-event CommonRecord + write_event under the
 * global mutex of MGraniitti::SaveEvent (src/MGraniitti.cc:1651-1726).  Returns the number of events written.  Host-only. */
int64_t grb_write_events(const grb_process_desc *d, const char *path, const grb_writer_opts *opts, int64_t n,
                         const double *momenta, const double *weights);

/* The same for the events accepted by the engine's last GRB_MODE_GENERATE iteration (device -> host -> file).        */
int64_t grb_write_engine_events(grb_engine *e, const char *path, const grb_writer_opts *opts);

/* ---- the integration loop (host C++, device iterations) ----------------------------------------------- */

/* INTEGRALPARAM.VEGAS of the steering card (reference VEGASPARAM, include/Graniitti/MVEGAS.h:18-32) */
typedef struct {
  int32_t  bins;        /* BINS (even, <= GRB_MAX_BINS)                                                     */
  int32_t  iter;        /* ITER: minimum iterations of the integration phase                               */
  int32_t  burnin_iter; /* burn-in iterations (reference SampleVegas: 3)                                    */
  int32_t  max_iter;    /* safety bound on the extended iteration count (0 = 1000)                         */
  int64_t  ncall;       /* NCALL: points per iteration, summed over ranks                                  */
  double   lambda;      /* LAMBDA                                                                          */
  double   chi2max;     /* CHI2MAX                                                                         */
  double   precision;   /* PRECISION: target relative error                                                */
  uint64_t seed;        /* RNDSEED: Philox key                                                             */
  int32_t  debug;       /* DEBUG >= 0 prints one line per iteration like the reference                     */
  int32_t  pad_;
} grb_vegas_param;

typedef struct {
  double  sigma, sigma_err, chi2; /* stat.sigma, stat.sigma_err, stat.chi2 (barn)                           */
  double  maxW, maxf;             /* for the hit-or-miss generation phase                                  */
  double  evaluations, n_kin_ok, n_fid_ok, n_veto_ok, n_amp_ok; /* integration-phase totals                */
  double  seconds;                /* host wall time of the whole call                                      */
  double  device_ms;              /* summed device time of all iterations on this rank                     */
  int64_t ncall_final;            /* NCALL after the chi2 > 50 => x10 restarts                             */
  int32_t iterations;             /* integration-phase iterations actually run                            */
  int32_t restarts;
  double  xmat[GRB_MAX_BINS * GRB_MAX_DIM]; /* converged grid [bins][dim]                                   */
} grb_vegas_result;

/* Called once per iteration with the DEVICE pointer of this rank's accumulators (layout grb_accum, on the
 * stream of grb_get_stream); must leave the all-rank sums / maxima in place.  NULL = single rank, or the
 * engine's own NCCL communicator when grb_comm_init was called.                                            */
typedef int (*grb_reduce_fn)(void *ctx, void *device_accum);

/* Burn-in + integration exactly as the reference drives them (MGraniitti::SampleVegas / VEGAS,
 * src/MGraniitti.cc:1074-1339): grid refinement every iteration, weighted average of the iteration
 * estimates, chi2, iteration extension until chi2 <= CHI2MAX and error <= PRECISION, x10 calls when
 * chi2 > 50.  Rank r evaluates calls [r*n/R, (r+1)*n/R) of every iteration (remainder to rank 0,
 * reference VEGASData::GetLocalCalls, MVEGAS.h:59-69).                                                      */
int grb_integrate(grb_engine *e, const grb_vegas_param *par, int rank, int nranks, grb_reduce_fn reduce,
                  void *ctx, grb_vegas_result *out);

/* INTEGRATOR "FLAT" (reference MGraniitti::SampleFlat, src/MGraniitti.cc:1446-1523; MCPARAM, MGraniitti.h:40-43): plain
 * Monte Carlo over the unit hypercube without a grid -- sigma = <W>, err = sqrt((<W^2> - <W>^2) / N) -- in device batches
 * of `batch` points (per batch all ranks together, sharded like grb_integrate), until the relative error is below
 * `precision` AND more than `min_events` points were evaluated (the reference tests after every point, here after every
 * batch).  out: sigma, sigma_err, maxW, the pass counters, evaluations; chi2 = 0, xmat = the uniform grid.            */
typedef struct {
  double   precision;  /* MCPARAM::PRECISION, default 0.05                                                             */
  double   min_events; /* MCPARAM::MIN_EVENTS, default 1e6                                                             */
  int64_t  batch;      /* points per device batch, 0 = 2^22                                                            */
  int64_t  max_events; /* safety bound, 0 = 2^40                                                                       */
  uint64_t seed;
} grb_flat_param;
int grb_sample_flat(grb_engine *e, const grb_flat_param *par, int rank, int nranks, grb_reduce_fn reduce, void *ctx,
                    grb_vegas_result *out);

/* The shard of an iteration's calls that rank `rank` of `nranks` evaluates (reference
 * VEGASData::GetLocalCalls, include/Graniitti/MVEGAS.h:59-69). Host-only.                                   */
int grb_local_calls(int64_t calls, int rank, int nranks, int64_t *first, int64_t *n);

/* ---- multi-GPU ------------------------------------------------------------------------------------- */

/* NCCL communicator bootstrap for hosts without torch.distributed (a C++ MGraniitti on each rank):
 * rank 0 calls grb_comm_unique_id, ships the 128 bytes to the other ranks by any means, then every rank
 * calls grb_comm_init.  grb_allreduce sums (fsum..n_overflow, fmat, f2mat) and max-reduces (maxW, maxf) the
 * device-resident accumulators of the last grb_vegas_iterate_async; no-op for a single rank.  This is
 * the B200 equivalent of the reference's global mutex (src/MGraniitti.cc:1389-1419).                     */
int grb_comm_unique_id(char id_out[128]);
int grb_comm_init(grb_engine *e, int rank, int nranks, const char id[128]);
int grb_allreduce(grb_engine *e);

/* ---- measurement helpers ---------------------------------------------------------------------------- */

/* Measured fp64 FMA peak of the device (dependent-free DFMA chains on all SMs), in TFLOP/s; this is the
 * roofline denominator for the path (MEASURED_PEAKS.json has no fp64 figure).                            */
int grb_measure_dfma_peak(int device, double *tflops_out);

/* ---- callers that loop the integrator ---------------------------------------------------- */

/* Re-parameterise an existing engine instead of re-creating it: the reference's energy scan and parameter fits build a
 * new MGraniitti per step (src/Program/xscan.cc:111-140, src/Program/Analysis/fitcentral.cc); here the device buffers,
 * streams and communicator stay, only the POD description is digested again.  The decay-tree topology (n_nodes, n_legs)
 * and the phase-space class must not change.  A changed s invalidates the installed eikonal table.              */
int grb_update_desc(grb_engine *e, const grb_process_desc *desc);

/* MProcess::SetInitialState / SetBeamEnergies for the beams of the description (src/MProcess.cc:291-334): new beam
 * energies (GeV), beam momenta and s.  Invalidates the eikonal table (rebuild with grb_build_eikonal).          */
int grb_set_beam_energies(grb_engine *e, double ebeam1, double ebeam2);

/* One row of the reference's `xscan` table (scan.csv: sqrts, xstot, xsin, xsel, xs_k): total / elastic / inelastic
 * cross sections of the eikonal model (mb) and the integrated process cross section (barn).                    */
typedef struct {
  double sqrt_s;
  double xs_tot_mb, xs_el_mb, xs_inel_mb; /* MEikonal::GetTotXS                                                */
  double sigma, sigma_err, chi2;          /* MGraniitti::GetXS                                                 */
  double evaluations, seconds;
} grb_xscan_row;

/* The loop of src/Program/xscan.cc:111-140 for this engine's process: for each sqrt(s) set symmetric beams of
 * sqrt(s)/2, build the eikonal tables on the device (grb_build_eikonal), integrate (grb_integrate, one rank).
 * The engine is left at the last energy.                                                                        */
int grb_xscan(grb_engine *e, const double *sqrt_s, int n, const grb_vegas_param *par, grb_xscan_row *rows);

/* Parameter fits (reference src/Program/Analysis/fitcentral.cc: every fit step re-creates the generator and re-integrates
 * for each parameter set): K descriptions of the same topology integrated back to back on ONE engine -- buffers, streams
 * and the communicator stay, each set costs one description digest and one constant upload (grb_update_desc).  The eikonal
 * table is kept when s and the loop discretisation do not change.  results[k] = grb_integrate of descs[k]; bit-identical
 * to K freshly created engines.  (The sets run one after the other: the kernels read their parameters from constant
 * memory, one set at a time.)                                                                                         */
int grb_integrate_batch(grb_engine *e, const grb_process_desc *descs, int k_sets, const grb_vegas_param *par, int rank,
                        int nranks, grb_reduce_fn reduce, void *ctx, grb_vegas_result *results);

/* Unit-test hook for graniitti_b200/csrc/grb_fastmath.cuh on n host values: kind 0 table-driven log, 1 table-driven
 * exp, 2 branch-free exact sqrt, 3 sqrt(), 4 branch-free exact x[i] / x[i+1], 5 operator/ (3 and 5 are the device's
 * own IEEE results the exact forms must reproduce bit for bit), 6 / 7 table-driven sin / cos,
 * 8 / 10 SolvePz3_A / _B in the reference's order, 9 / 11 their power-of-two-folded forms (must agree bit for bit). */
int grb_test_fastmath(int device, int kind, int64_t n, const double *x, double *y);

/* Device time in ms of the kernels launched by the last grb_eval / grb_vegas_iterate call (CUDA events on
 * the engine's stream), and how many kernels that call launched.                                         */
int grb_last_timing(const grb_engine *e, double *kernel_ms, int *n_launches);

/* Same, split by kernel: total device time of the call and the summed device time of its screening-loop
 * launches (k_screen, the dominant kernel of screened processes), each bracketed by CUDA events.          */
int grb_last_timing_kernels(const grb_engine *e, double *ms_total, double *ms_screen, int *n_screen_launches);

/* The engine's CUDA stream (cudaStream_t) so that a caller can order its own work (collectives, event
 * timing) after the engine's kernels without a host synchronisation.                                      */
void *grb_get_stream(grb_engine *e);

/* Block until the engine's stream is idle (after grb_vegas_iterate_async).                               */
int grb_sync(grb_engine *e);

/* sizeof() of the ABI structs as compiled (0 desc, 1 accum, 2 resonance, 3 node): layout check for FFI.   */
int grb_abi_sizeof(int what);

#ifdef __cplusplus
}
#endif
#endif /* GRB_H */
