"""ctypes mirror of include/grb.h (the C ABI). Plumbing only: no physics lives in Python.

The layout is checked against the compiled library by `grb_abi_sizeof` (tests/test_abi.py)."""
import ctypes as C

GRB_ABI_VERSION = 3
GRB_MAX_NODES = 16
GRB_MAX_RES = 12
GRB_MAX_VETO = 8
GRB_MAX_DIM = 20
GRB_MAX_TAPE = 32
GRB_MAX_BINS = 128
GRB_RHO_DIM = 5
GRB_MAX_HAMP = 36

GRB_OK, GRB_ERR_ARG, GRB_ERR_CUDA, GRB_ERR_UNSUPPORTED, GRB_ERR_PARSE, GRB_ERR_STATE, GRB_ERR_NCCL = (
    0, -1, -2, -3, -4, -5, -6)

GRB_PROC_YY_CON, GRB_PROC_YY_QED, GRB_PROC_PP_RES = 3, 4, 9
GRB_PROC_PP_CONTENSOR, GRB_PROC_PP_CON, GRB_PROC_PP_CON_MINUS, GRB_PROC_PP_RES_CON = 12, 15, 16, 17
GRB_PROC_PP_RESTENSOR, GRB_PROC_PP_RES_CONTENSOR = 11, 14

GRB_FLAG_KIN, GRB_FLAG_FID, GRB_FLAG_VETO, GRB_FLAG_AMP = 1, 2, 4, 8
GRB_FLAGS_VALID = 7

GRB_MODE_BURNIN, GRB_MODE_INTEGRATE, GRB_MODE_GENERATE = 0, 1, 2
GRB_CLASS_F, GRB_CLASS_C = 0, 1

DUMP_NAMES = ["t1", "t2", "s1", "s2", "m2", "Y", "Pt", "t_hat", "u_hat", "amp2", "hamp0_re", "hamp0_im"]
GRB_DUMP_LEN = len(DUMP_NAMES)

i32, f64 = C.c_int32, C.c_double


class Node(C.Structure):
    _fields_ = [("pdg", i32), ("spinX2", i32), ("parent", i32), ("first_child", i32), ("n_children", i32),
                ("pad_", i32), ("mass", f64), ("width", f64)]


class GenCut(C.Structure):
    _fields_ = [(k, f64) for k in ("Y_min", "Y_max", "M_min", "M_max", "forward_pt_min", "forward_pt_max",
                                   "XI_min", "XI_max", "kt_min", "kt_max", "rap_min", "rap_max")]


class FidCut(C.Structure):
    _fields_ = [("active", i32), ("usercuts", i32)] + [(k, f64) for k in (
        "eta_min", "eta_max", "rap_min", "rap_max", "pt_min", "pt_max", "Et_min", "Et_max",
        "M_min", "M_max", "Y_min", "Y_max", "Pt_min", "Pt_max",
        "forward_t_min", "forward_t_max", "forward_M_min", "forward_M_max")]


class VetoDomain(C.Structure):
    _fields_ = [(k, f64) for k in ("eta_min", "eta_max", "pt_min", "pt_max")]


class ParamSoft(C.Structure):
    _fields_ = [(k, f64) for k in ("DELTA_P", "ALPHA_P", "gN_P", "gN_O", "g3P", "gamma", "fc1", "fc2", "fc3")] + [
        ("odderon_on", i32), ("f2_model", i32)]


class ParamRegge(C.Structure):
    _fields_ = [("a0", f64 * 3), ("ap", f64 * 3), ("sgn", f64 * 3), ("s0", f64), ("offshellFF", i32),
                ("reggeize", i32), ("b_EXP", f64), ("a_OREAR", f64), ("b_OREAR", f64), ("b_POW", f64),
                ("omega", f64), ("c", f64 * 3), ("n", i32 * 3), ("pad_", i32)]


class ParamTensorPom(C.Structure):
    _fields_ = [(k, f64) for k in ("delta_P", "delta_O", "delta_1R", "delta_2R", "ap_P", "ap_O", "ap_1R",
                                   "ap_2R", "eta_O", "gPNN", "gPpipi", "gPKK")] + [
        ("gPrhorho", f64 * 2), ("gPphiphi", f64 * 2), ("M_O", f64), ("M_1R", f64)]


class Resonance(C.Structure):
    _fields_ = [("pdg", i32), ("spinX2", i32), ("P", i32), ("BW", i32), ("spingen", i32), ("spindec", i32),
                ("frame", i32), ("jmax", i32), ("mass", f64), ("width", f64), ("g_re", f64), ("g_im", f64),
                ("g_FF", f64), ("g_decay", f64), ("BR", f64), ("g_Tensor", f64 * 8), ("g_decay_tensor", f64 * 2),
                ("rho_re", f64 * (GRB_RHO_DIM * GRB_RHO_DIM)), ("rho_im", f64 * (GRB_RHO_DIM * GRB_RHO_DIM)),
                ("T_re", f64 * 9), ("T_im", f64 * 9), ("T_rows", i32), ("T_cols", i32)]


class ProcessDesc(C.Structure):
    _fields_ = [(k, i32) for k in (
        "abi_version", "ps_class", "proc", "screening", "excitation", "flatamp", "ps_active", "flatmass2",
        "lipsdim", "n_tape", "n_nodes", "n_legs", "n_final", "n_res", "n_veto", "veto_active", "em_ff",
        "qed_alpha", "loop_nkt", "loop_nphi")] + [(k, f64) for k in (
            "loop_min_kt", "loop_max_kt", "s", "sqrt_s", "pbeam1_pz", "pbeam1_E", "pbeam2_pz", "pbeam2_E",
            "beam1_mass", "beam2_mass", "S_factor", "offshell", "flat_b")] + [
        ("nodes", Node * GRB_MAX_NODES), ("gcuts", GenCut), ("fcuts", FidCut),
        ("veto", VetoDomain * GRB_MAX_VETO), ("soft", ParamSoft), ("regge", ParamRegge),
        ("tpom", ParamTensorPom), ("res", Resonance * GRB_MAX_RES)]


class Accum(C.Structure):
    _fields_ = [(k, f64) for k in ("fsum", "f2sum", "evaluations", "n_kin_ok", "n_fid_ok", "n_veto_ok",
                                   "n_amp_ok", "n_nan_inf", "n_accepted", "n_overflow", "maxW", "maxf")] + [
        ("fmat", f64 * (GRB_MAX_BINS * GRB_MAX_DIM)), ("f2mat", f64 * (GRB_MAX_BINS * GRB_MAX_DIM))]


ACCUM_LEN = C.sizeof(Accum) // 8   # doubles of one grb_accum; the engine's device buffer has 128 more (per-rank maxima slots)
ACCUM_NSUM = 10  # leading doubles that are sum-reduced; then maxW, maxf (max-reduced); then fmat, f2mat (sum)


class VegasParam(C.Structure):
    _fields_ = [("bins", i32), ("iter", i32), ("burnin_iter", i32), ("max_iter", i32), ("ncall", C.c_int64),
                ("lambda_", f64), ("chi2max", f64), ("precision", f64), ("seed", C.c_uint64), ("debug", i32),
                ("pad_", i32)]


class VegasResult(C.Structure):
    _fields_ = [(k, f64) for k in ("sigma", "sigma_err", "chi2", "maxW", "maxf", "evaluations", "n_kin_ok",
                                   "n_fid_ok", "n_veto_ok", "n_amp_ok", "seconds", "device_ms")] + [
        ("ncall_final", C.c_int64), ("iterations", i32), ("restarts", i32),
        ("xmat", f64 * (GRB_MAX_BINS * GRB_MAX_DIM))]


REDUCE_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_void_p)


class XscanRow(C.Structure):
    _fields_ = [(k, f64) for k in ("sqrt_s", "xs_tot_mb", "xs_el_mb", "xs_inel_mb", "sigma", "sigma_err", "chi2",
                                   "evaluations", "seconds")]


class RecParticle(C.Structure):
    _fields_ = [("id", i32), ("pid", i32), ("status", i32), ("parent", i32)] + [(k, f64) for k in ("px", "py", "pz", "e", "m")]


class RecVertex(C.Structure):
    _fields_ = [("id", i32), ("status", i32), ("n_in", i32), ("n_out", i32), ("in_", i32 * 4)] + [(k, f64) for k in ("x", "y", "z", "t")]


GRB_MAX_REC_PARTICLES = 7 + GRB_MAX_NODES
GRB_MAX_REC_VERTICES = 4 + GRB_MAX_NODES


class WriterOpts(C.Structure):
    _fields_ = [("format", C.c_char_p), ("generator_name", C.c_char_p), ("generator_version", C.c_char_p), ("card_path", C.c_char_p),
                ("sigma_pb", f64), ("sigma_err_pb", f64), ("beam1_pdg", i32), ("beam2_pdg", i32), ("weighted", i32), ("nthreads", i32),
                ("append", i32), ("close_listing", i32), ("number_events", i32), ("lifetimes", i32), ("first_event", C.c_int64),
                ("lifetime_seed", C.c_uint64)]


class FlatParam(C.Structure):
    _fields_ = [("precision", f64), ("min_events", f64), ("batch", C.c_int64), ("max_events", C.c_int64), ("seed", C.c_uint64)]
