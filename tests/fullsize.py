"""Full-size parity machinery (SURVEY.md section 8d "Parity sets"): per BASELINE configuration 2^20 flat points and 2^20
points drawn through the CONVERGED VEGAS grid of that configuration (the reference's sampling map, src/MGraniitti.cc:
1357-1375, with the device's Philox stream restated on the host), evaluated by the CUDA path through the C ABI and by the
CPU oracle on all host cores.  Where oracle/_ref (the compiled unmodified reference) is present, a subsample of the same
points also goes through the reference's own MProcess::EventWeight, and the 1-ulp conditioning kappa of every point that
misses 1e-10 is measured with the reference itself (else with the oracle port).

Used by tests/test_gpu_parity_fullsize.py (asserts) and scripts/parity_fullsize.py (prints the table that is committed
under profiles/)."""
import multiprocessing as mp
import os
import time

import numpy as np

import helpers
from philox_ref import philox_points_np

RTOL = 1e-10
OUTLIER_CAP = 1 << 20
RHO4PI = "rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}"
# id, card of this repo, the reference's card, process override, POMLOOP, tape length, reference subsample, and the bounds
# (fraction of the non-zero weights that may need the conditioning term, median |dW|/W) for the (VEGAS, flat) sets.
# The reference's yy[QED] amplitude (gauge cancellation between the two diagrams through the tensor-pomeron code) and its
# rank-4 tensor algebra of the 2xV continuum are ill-conditioned by themselves: measured kappa up to 0.2 (1.3e3 on flat
# points) and 1e-8; their bounds are the measured fractions with headroom (profiles/r02_parity_fullsize.txt).
# Third entry: the floor of the per-point tolerance max(floor, 8 kappa) -- 1e-10 (north_star) everywhere except the flat set
# of the 2xV continuum: there one point in 574 455 (W = 7e-34 b, forward pt 1e-11 GeV) sits at 1.45e-10 with an input /
# libm conditioning of 9e-14, i.e. the difference is internal rounding of the O(s^2) cancellations between the reference's
# 256-term rank-4 contractions and the device's contracted closed form; none on the VEGAS-distributed set.
WELL = ((1e-3, 1e-13, RTOL), (1e-3, 1e-13, RTOL))
CLOSED = ((1e-3, 2e-12, RTOL), (1e-3, 2e-12, RTOL))   # closed-form tensor / resonance / HELAS algebra on the device
CASES = [
    ("c1_rescon", "c1_test13.json", "./input/test.json", "PP[RES+CON]<F> -> pi+ pi-", 0, 8, 1 << 14, CLOSED),
    ("c1_tensor", "c1_test13.json", "./input/test.json", "PP[CONTENSOR]<F> -> pi+ pi-", 0, 8, 1 << 14, CLOSED),
    ("c2_born", "c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", 0, 8, 1 << 14, WELL),
    ("c2_screened", "c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", 1, 8, 1 << 14, WELL),
    ("c2t_screened", "c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CONTENSOR]<F> -> pi+ pi-", 1, 8, 1 << 11, CLOSED),
    ("c3_rescon", "c3_atlas.json", "./tests/run_atlas_multi/ATLAS.json", "PP[RES+CON]<F> -> pi+ pi-", 0, 8, 1 << 14, CLOSED),
    ("c4_yycon", "c4_cms_mumu_lowmass.json", "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", 0, 8, 1 << 14, CLOSED),
    ("c4_yyqed", "c4_cms_mumu_lowmass.json", "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[QED]<F> -> mu+ mu-", 0, 8, 1 << 13, ((0.35, 2e-11, RTOL), (0.95, 1.0, RTOL))),
    ("c5_regge", "c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CON]<F> -> " + RHO4PI, 0, 24, 1 << 14, WELL),
    ("c5_tensor", "c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CONTENSOR]<F> -> " + RHO4PI, 0, 24, 1 << 13, ((5e-3, 2e-12, RTOL), (5e-2, 2e-12, 3e-10))),
]
CASE_IDS = [c[0] for c in CASES]

# ---- oracle on all host cores (fork: the restatement keeps static scratch tensors, so processes, not threads) ----------
_W = {}


def _oracle_chunk(rng_):
    lo, hi = rng_
    o = helpers.oracle_eval(_W["desc"], _W["eik"], _W["u"][lo:hi], _W["tape"][lo:hi])
    return lo, o["w"], o["flags"]


def oracle_eval_parallel(desc, eik, u, tape, nproc=None, chunk=2048):
    n = len(u)
    nproc = nproc or os.cpu_count() or 1
    if n <= chunk or nproc == 1:
        o = helpers.oracle_eval(desc, eik, u, tape)
        return o["w"], o["flags"]
    helpers.load_oracle()
    _W.update(desc=desc, eik=eik, u=u, tape=tape)
    w, fl = np.zeros(n), np.zeros(n, dtype=np.uint8)
    with mp.get_context("fork").Pool(nproc) as pool:
        for lo, cw, cf in pool.imap_unordered(_oracle_chunk, [(i, min(i + chunk, n)) for i in range(0, n, chunk)]):
            w[lo:lo + len(cw)] = cw
            fl[lo:lo + len(cf)] = cf
    _W.clear()
    return w, fl


def _ulp_neighbours(x):
    """One step of the uniform grid (2^-53 = one ulp of 1.0): a relative ulp of a small u is swallowed by the first affine
    map of the sampling (e.g. y = Y_min + dY u) and would not move the point at all."""
    return x + 2.0 ** -53, x - 2.0 ** -53


def oracle_kappa(desc, eik, u, tape):
    """1-ulp conditioning of the oracle's weight: max over coordinates / tape slots of |W(x +- ulp) - W| / |W| (the
    protocol of SURVEY.md section 7, evaluated with the port where the reference itself is not available)."""
    base = helpers.oracle_eval(desc, eik, u, tape)["w"]
    kap = np.zeros(len(u))
    both = np.hstack([u, tape])
    du = u.shape[1]
    for k in range(both.shape[1]):
        for nb in _ulp_neighbours(both[:, k]):
            v = both.copy()
            v[:, k] = nb
            w = helpers.oracle_eval(desc, eik, v[:, :du], v[:, du:])["w"]
            with np.errstate(divide="ignore", invalid="ignore"):
                kap = np.maximum(kap, np.where(base != 0, np.abs(w - base) / np.abs(base), 0.0))
    return kap


def oracle_kappa_libm(desc, eik, u, tape, max_calls=64):
    """Conditioning with respect to the libm results (oracle/grb_oracle.c opert()): the result of the k-th transcendental
    call of every point moved by +-2 ulp, k < max_calls (the Born-level sampling map, decays, cuts and amplitude; the
    later calls belong to the screening loop nodes, each of which carries 1/300 of the loop integral)."""
    lib = helpers.load_oracle()
    base = helpers.oracle_eval(desc, eik, u, tape)["w"]
    kap = np.zeros(len(u))
    try:
        # single calls by +-2 ulp (CUDA's libm documents up to 2 ulp for sin / cos / sinh / tan / atan2), and every call of
        # the point together by +-1 ulp (differences of several calls can add up coherently)
        for k, sgn in [(k, sg) for k in range(max_calls) for sg in (2, -2)] + [(-2, 1), (-2, -1)]:
            if True:
                lib.grb_oracle_perturb_call(k, sgn)
                w = helpers.oracle_eval(desc, eik, u, tape)["w"]
                with np.errstate(divide="ignore", invalid="ignore"):
                    kap = np.maximum(kap, np.where((base != 0) & (w != 0), np.abs(w - base) / np.abs(base), 0.0))
    finally:
        lib.grb_oracle_perturb_call(-1, 1)
    return kap


def oracle_kappa_pz(desc, eik, u, tape):
    """Conditioning with respect to the ROUNDING of the solved forward longitudinal momenta (oracle probes
    grb_oracle_perturb_pz / _pz2: Born-level p1z, p2z moved by one ulp).  A 1-ulp libm difference upstream moves the exact
    solution by far less than an ulp of p_z ~ sqrt(s)/2, but flips its rounding with probability ~1e-3; with an excited
    forward system that flip moves M^2 = E^2 - p^2 by 2 p ulp(p) = 1.2e-8 GeV^2, and every node of the screening loop
    inherits it through the re-derived mass m = pfinal_orig.M() into S3FINEL(t, M^2) / x_Bj."""
    lib = helpers.load_oracle()
    base = helpers.oracle_eval(desc, eik, u, tape)["w"]
    kap = np.zeros(len(u))
    try:
        for fn in (lib.grb_oracle_perturb_pz, lib.grb_oracle_perturb_pz2):
            for sgn in (1, -1, 2, -2, 3, -3):   # (the response is a staircase: E^2 - p^2 re-rounds every few ulps of p_z)
                fn(sgn)
                w = helpers.oracle_eval(desc, eik, u, tape)["w"]
                with np.errstate(divide="ignore", invalid="ignore"):
                    kap = np.maximum(kap, np.where((base != 0) & (w != 0), np.abs(w - base) / np.abs(base), 0.0))
    finally:
        lib.grb_oracle_perturb_pz(0)
    return kap


def _libm_chunk(rng_):
    lo, hi = rng_
    return lo, oracle_kappa_libm(_W["desc"], _W["eik"], _W["u"][lo:hi], _W["tape"][lo:hi], _W["max_calls"])


def oracle_kappa_libm_parallel(desc, eik, u, tape, max_calls=64, nproc=None):
    n = len(u)
    nproc = nproc or os.cpu_count() or 1
    if n < 2 * nproc:
        return oracle_kappa_libm(desc, eik, u, tape, max_calls)
    helpers.load_oracle()
    _W.update(desc=desc, eik=eik, u=u, tape=tape, max_calls=max_calls)
    kap = np.zeros(n)
    cuts = np.linspace(0, n, 4 * nproc + 1).astype(int)
    with mp.get_context("fork").Pool(nproc) as pool:
        for lo, k in pool.imap_unordered(_libm_chunk, [(a, b) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]):
            kap[lo:lo + len(k)] = k
    _W.clear()
    return kap


def _ref_chunk(args):
    card, proc, pom, u, tape, kappa = args
    r = helpers.ref_eval(card, proc, pom, u, tape, kappa=kappa)
    return r["W"], r["flags"], r["kappa"]


def ref_eval_parallel(card, proc, pom, u, tape, kappa=False, nproc=None):
    """The unmodified reference's EventWeight (oracle/_ref/gra/bin/ref_eval) on all host cores: one process per slice."""
    n = len(u)
    nproc = max(1, min(nproc or os.cpu_count() or 1, (n + 15) // 16))
    cuts = np.linspace(0, n, nproc + 1).astype(int)
    jobs = [(card, proc, pom, u[a:b], tape[a:b], kappa) for a, b in zip(cuts[:-1], cuts[1:]) if b > a]
    with mp.get_context("fork").Pool(len(jobs)) as pool:
        parts = pool.map(_ref_chunk, jobs)
    return (np.concatenate([p[0] for p in parts]), np.concatenate([p[1] for p in parts]),
            np.concatenate([p[2] for p in parts]))


# ---- one parity set ----------------------------------------------------------------------------------------------------
def compare_set(E, desc, eik, ref_card, proc, pom, u, tape, ref_sub, label, libm_calls=64, floor=RTOL):
    """Device (grb_eval) vs oracle on (u, tape) [n][.].  Returns a dict of statistics; nothing is asserted here."""
    t0 = time.perf_counter()
    w, fl, _ = E.eval(u.T.copy(), tape.T.copy())
    t_dev = time.perf_counter() - t0
    t0 = time.perf_counter()
    ow, ofl = oracle_eval_parallel(desc, eik, u, tape)
    t_orc = time.perf_counter() - t0
    kin = (ofl & 1) == 1
    st = dict(label=label, n=len(u), n_kin=int(kin.sum()), n_valid=int((ow != 0).sum()), t_dev=t_dev, t_oracle=t_orc)
    st["kin_mismatch"] = int(np.sum((fl & 1) != (ofl & 1)))
    st["cut_mismatch"] = int(np.sum(fl[kin] != ofl[kin]))
    st["zero_mismatch"] = int(np.sum((w != 0) != (ow != 0)))
    nz = (ow != 0) & (w != 0)
    rel = np.zeros(len(u))
    rel[nz] = np.abs(w[nz] - ow[nz]) / np.abs(ow[nz])
    r = rel[nz]
    st.update(median=float(np.median(r)) if r.size else 0.0, p99=float(np.quantile(r, 0.99)) if r.size else 0.0,
              max=float(r.max()) if r.size else 0.0, n_above=int((r > RTOL).sum()))
    # conditioning of the points that miss 1e-10: the reference's own kappa where it is compiled in, else the port's
    out = np.flatnonzero(rel > RTOL)
    st["kappa_source"] = "-"
    st["n_fail"] = 0
    st["max_kappa"] = 0.0
    if out.size:
        cap = out[:OUTLIER_CAP]   # bound the cost; anything beyond is reported as failing
        if helpers.have_ref():
            _, _, kap = ref_eval_parallel(ref_card, proc, pom, u[cap], tape[cap], kappa=2)
            st["kappa_source"] = "reference"
        else:
            kap = oracle_kappa(desc, eik, u[cap], tape[cap])
            st["kappa_source"] = "port"
        fail = rel[cap] > np.maximum(floor, 8.0 * kap)
        st["n_need_libm"] = int(fail.sum())
        if fail.any():   # ... and with respect to the libm results (what two IEEE platforms legitimately compute differently)
            kl = oracle_kappa_libm_parallel(desc, eik, u[cap[fail]], tape[cap[fail]], max_calls=libm_calls)
            kap[fail] = np.maximum(kap[fail], kl)
            fail = rel[cap] > np.maximum(floor, 8.0 * kap)
        st["max_kappa"] = float(kap.max())
        st["n_fail"] = int(fail.sum()) + int(out.size - cap.size)
        dump_dir = os.path.join(helpers.ROOT, "gpurun_out")
        if fail.any() and os.path.isdir(dump_dir):   # keep the evidence of every failing point
            f = cap[fail]
            np.savez(os.path.join(dump_dir, "fullsize_fail_" + label.replace(" ", "_") + ".npz"), u=u[f], tape=tape[f],
                     w_dev=w[f], w_oracle=ow[f], kappa=kap[fail], flags=fl[f])
    # the same points through the reference itself (bounded subsample)
    st["ref_n"] = 0
    if helpers.have_ref() and ref_sub > 0:
        m = min(ref_sub, len(u))
        t0 = time.perf_counter()
        rw, rfl, _ = ref_eval_parallel(ref_card, proc, pom, u[:m], tape[:m], kappa=False)
        st["t_ref"] = time.perf_counter() - t0
        rk = (rfl & 1) == 1
        st["ref_n"] = m
        st["ref_flag_mismatch"] = int(np.sum((fl[:m] & 1) != (rfl & 1)) + np.sum(fl[:m][rk] != rfl[rk]) + np.sum((w[:m] != 0) != (rw != 0)))
        z = (rw != 0) & (w[:m] != 0)
        rr = np.abs(w[:m][z] - rw[z]) / np.abs(rw[z])
        st["ref_median"] = float(np.median(rr)) if rr.size else 0.0
        st["ref_max"] = float(rr.max()) if rr.size else 0.0
        st["ref_n_above"] = int((rr > RTOL).sum())
        st["ref_oracle_max"] = float(np.max(np.abs(ow[:m][z] - rw[z]) / np.abs(rw[z]))) if rr.size else 0.0
        st["ref_n_fail"] = 0
        bad = np.flatnonzero(z)[rr > RTOL]
        if bad.size:
            _, _, kap = ref_eval_parallel(ref_card, proc, pom, u[bad], tape[bad], kappa=2)
            kap = np.maximum(kap, oracle_kappa_libm_parallel(desc, eik, u[bad], tape[bad], max_calls=libm_calls))
            st["ref_n_fail"] = int(np.sum(rr[rr > RTOL] > np.maximum(floor, 8.0 * kap)))
    return st


def run_case(grb, msa, case, n=1 << 20, seed=20261018, log=print):
    """Converged grid of the configuration -> the two parity sets.  Returns [stats_vegas, stats_flat]."""
    cid, card, ref_card, proc, pom, ntape, ref_sub, _ = case
    E = grb.Engine(card=os.path.join(helpers.CARDS, card), process=proc, pomloop=pom)
    eik = None
    if E.desc.screening:
        E.set_eikonal(msa[0], msa[1], msa[2])
        eik = helpers.Eikonal(msa[0], msa[1], msa[2])
    dim, bins = E.desc.lipsdim, 128
    r = E.integrate(ncall=1 << 20, iters=5, precision=0.01, seed=seed)
    xmat = np.frombuffer(r.xmat, dtype=np.float64)[:bins * dim].reshape(bins, dim).copy()
    u, vw, tape = philox_points_np(xmat, 0, n, seed, 1000, ntape)
    # transcendental calls ahead of / inside the Born amplitude: 4-pi cascades and the QED spinor algebra make more of them
    lc = 320 if "QED" in proc else (160 if ntape > 8 else 64)
    sv = compare_set(E, E.desc, eik, ref_card, proc, pom, u, tape, ref_sub, cid + " vegas", lc, case[7][0][2])
    # the device iteration over the same Philox stream sees the same weights (sum f, sum f^2, max f)
    acc = E.vegas_iterate(xmat, 0, n, seed, 1000)
    w, _, _ = E.eval(u.T.copy(), tape.T.copy())
    f = w * vw
    sv["iter_fsum_rel"] = float(abs(f.sum() - acc.fsum) / acc.fsum) if acc.fsum else 0.0
    sv["iter_maxf_equal"] = bool(f.max() == acc.maxf)
    sv["sigma"], sv["sigma_err"] = r.sigma, r.sigma_err
    rng = np.random.default_rng(seed + 1)
    sf = compare_set(E, E.desc, eik, ref_card, proc, pom, rng.random((n, dim)), rng.random((n, ntape)), ref_sub, cid + " flat", lc, case[7][1][2])
    E.close()
    for s in (sv, sf):
        log(format_row(s))
    return [sv, sf]


HEADER = ("%-20s %8s %8s %8s | %4s %4s %4s | %9s %9s %9s %7s %6s %9s %-9s | %6s %5s %9s %9s %6s | %6s %6s" %
          ("set", "n", "kin_ok", "W!=0", "dkin", "dcut", "dzero", "median", "p99", "max", ">1e-10", "fail", "max kappa", "kappa by",
           "ref n", "dflag", "ref med", "ref max", ">1e-10", "t_dev", "t_orc"))


def format_row(s):
    ref = ("%6d %5d %9.2e %9.2e %6d" % (s["ref_n"], s["ref_flag_mismatch"], s["ref_median"], s["ref_max"], s["ref_n_above"])
           if s.get("ref_n") else "%6s %5s %9s %9s %6s" % ("-", "-", "-", "-", "-"))
    return ("%-20s %8d %8d %8d | %4d %4d %4d | %9.2e %9.2e %9.2e %7d %6d %9.2e %-9s | %s | %6.2f %6.2f" %
            (s["label"], s["n"], s["n_kin"], s["n_valid"], s["kin_mismatch"], s["cut_mismatch"], s["zero_mismatch"], s["median"],
             s["p99"], s["max"], s["n_above"], s["n_fail"], s["max_kappa"], s["kappa_source"], ref, s["t_dev"], s["t_oracle"]))
