"""Tensor-pomeron resonances (reference MTensorPomeron::ME3, PROC_11 "PP[RESTENSOR]" / PROC_14 "PP[RES+CONTENSOR]").

CPU side of the parity chain for this amplitude:
  reference (explicit rank-6 tensors, FTensor)  ==  oracle/grb_oracle_tres.inc (same tensors, same contraction order)
      -> tests/test_oracle_golden.py on the c*_restensor_* / *_rescontensor_* goldens
  oracle  ==  graniitti_b200/csrc/grb_tres.cuh (the contracted closed form the CUDA kernels run), compiled here by the HOST
      compiler through tests/host_tres.cc -> this file; the same header compiled by nvcc is checked on the GPU by
      tests/test_gpu_parity.py::test_device_vs_reference_golden on the same goldens.
Tolerance as everywhere: relative max(1e-10, 8 kappa) with kappa the reference's own 1-ulp input conditioning of the point."""
import ctypes as C
import glob
import os
import subprocess

import numpy as np
import pytest

from helpers import CARDS, GOLDEN, ROOT, load_oracle, oracle_eval
from test_oracle_golden import load_case

FILES = sorted(glob.glob(os.path.join(GOLDEN, "c*_restensor_*born.npz")) + glob.glob(os.path.join(GOLDEN, "late", "c*_restensor_*born.npz")))


class ResPar(C.Structure):   # grb::tres::ResPar
    _fields_ = [("J", C.c_int), ("on", C.c_int), ("M0sq", C.c_double), ("M0G", C.c_double), ("g", C.c_double * 7), ("gd", C.c_double)]


@pytest.fixture(scope="module")
def host_tres():
    so = os.path.join(ROOT, "tests", "_host_tres.so")
    deps = [os.path.join(ROOT, "tests", "host_tres.cc"), os.path.join(ROOT, "graniitti_b200", "csrc", "grb_tres.cuh")]
    if not os.path.exists(so) or any(os.path.getmtime(d) > os.path.getmtime(so) for d in deps):
        subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-shared", "-fPIC", deps[0], "-o", so])
    H = C.CDLL(so)
    assert H.tres_sizeof_respar() == C.sizeof(ResPar)
    return H


def proton_F1(d, t):   # reference src/MForm.cc:542-548,607-645
    Q2 = np.abs(t)
    tau = Q2 / (2 * 0.938272081) ** 2
    if d.em_ff:
        GE = (1 - 0.24 * tau) / (1 + 10.98 * tau + 12.82 * tau ** 2 + 21.97 * tau ** 3)
        GM = 2.792847337 * (1 + 0.12 * tau) / (1 + 10.97 * tau + 18.86 * tau ** 2 + 6.55 * tau ** 3)
    else:
        GM = 2.792847337 / (1 + Q2 / 0.71) ** 2
        GE = GM / 2.792847337
    return GE / (tau + 1) + tau / (tau + 1) * GM


def respars(d):
    rp = (ResPar * max(1, d.n_res))()
    for i in range(d.n_res):
        r = d.res[i]
        rp[i].J, rp[i].M0sq, rp[i].M0G, rp[i].gd = r.spinX2 // 2, r.mass ** 2, r.mass * r.width, r.g_decay_tensor[0]
        on = 0
        for m in range(7):
            rp[i].g[m] = r.g_Tensor[m]
            if m >= 1 and abs(r.g_Tensor[m]) > 1e-6:
                on |= 1 << m
        rp[i].on = on
    return rp


def closed_form(H, d, cap, extended=0):
    """K x resonance_sum of grb_tres.cuh on the momenta the oracle captured."""
    out = np.zeros((len(cap), 2))
    mom = np.ascontiguousarray(cap)
    H.tres_eval(respars(d), d.n_res, C.c_long(len(cap)), mom.ctypes.data_as(C.c_void_p), cap.shape[1], out.ctypes.data_as(C.c_void_p), extended)
    s1, s2, t1, t2 = cap[:, 28], cap[:, 29], cap[:, 30], cap[:, 31]
    T = d.tpom
    a1, a2 = T.delta_P + T.ap_P * t1, T.delta_P + T.ap_P * t2
    K = (9 * T.gPNN ** 2 * proton_F1(d, t1) * proton_F1(d, t2) / (1 - t1) / (1 - t2) *
         np.exp(a1 * np.log(s1 * T.ap_P) + a2 * np.log(s2 * T.ap_P)) / (16 * s1 * s2) * np.exp(-0.5j * np.pi * (a1 + a2)))
    return K * (out[:, 0] + 1j * out[:, 1])


def captured(d, u, tape):
    lib = load_oracle()
    lib.grb_oracle_capture.argtypes = [C.c_void_p]
    cap = np.zeros((len(u), 34))
    lib.grb_oracle_capture(cap.ctypes.data)
    try:
        o = oracle_eval(d, None, u, tape)
    finally:
        lib.grb_oracle_capture(None)
    return o, cap


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_closed_form_matches_reference_amplitudes(grb, msa, host_tres, path):
    z, d = load_case(grb, path)
    o, cap = captured(d, z["u"], z["tape"])
    nz = z["W"] != 0
    assert nz.sum() > 100
    ref = z["hamp"][nz, 0]                      # the reference's own first helicity amplitude (lts.hamp[0])
    A = closed_form(host_tres, d, cap[nz])
    rel = np.abs(A - ref) / np.abs(ref)
    # the explicit rank-6 contraction of the reference is the noisy side: its 1-ulp input conditioning kappa reaches 1e-7 on
    # the points where the closed form (double vs long double: <= 3e-11) and the reference disagree by 1e-8
    assert np.all(rel <= np.maximum(1e-10, 8.0 * z["kappa"][nz])), rel.max()
    assert np.median(rel) < 2e-13
    # and the oracle's restatement of the same explicit tensors agrees with the reference's amplitude to rounding
    oa = cap[nz, 32] + 1j * cap[nz, 33]
    assert np.max(np.abs(oa - ref) / np.abs(ref)) < 5e-12


def test_closed_form_is_well_conditioned(grb, msa, host_tres):
    """double vs long double evaluation of the closed form on the all-couplings golden: the contracted algebra loses nothing
    the reference's explicit contraction keeps."""
    z, d = load_case(grb, os.path.join(GOLDEN, "c5_cms_pipi_restensor_f2_allg_born.npz"))
    o, cap = captured(d, z["u"], z["tape"])
    nz = z["W"] != 0
    a, b = closed_form(host_tres, d, cap[nz]), closed_form(host_tres, d, cap[nz], extended=1)
    rel = np.abs(a - b) / np.abs(b)
    assert rel.max() < 1e-10 and np.median(rel) < 1e-13, (rel.max(), np.median(rel))


def test_unsupported_resonance_classes_are_flagged(grb):
    """0- resonances (-> gamma gamma only in the reference) and unknown gamma-V couplings have no amplitude for two pseudoscalars:
    the oracle reports NaN -> amplitude_ok false; the engine refuses them (and the 1- block, which only the oracle restates) at
    create time (tests/test_gpu_parity.py::test_tensor_resonance_scope).  Nothing is silently evaluated as something else."""
    from helpers import make_points
    d = grb.parse_card(os.path.join(CARDS, "c5_cms_4pi.json"), "PP[RESTENSOR]<F> -> pi+ pi- @RES{f0_980:1}", 0)
    d.res[0].P = -1                                 # a 0- state as MTensorPomeron::ME3 would see it
    u, tape = make_points(512, 6, 8, 7)
    u[:, 5] *= 0.1
    w = oracle_eval(d, None, u, tape)
    passed = (w["flags"] & 7) == 7
    assert passed.sum() > 20 and np.all(w["w"] == 0) and np.all((w["flags"][passed] & 8) == 0)

def _raw_sum(H, rp, n_res, mom, extended=0):
    out = np.zeros((len(mom), 2))
    mom = np.ascontiguousarray(mom)
    H.tres_eval(rp, n_res, C.c_long(len(mom)), mom.ctypes.data_as(C.c_void_p), mom.shape[1], out.ctypes.data_as(C.c_void_p), extended)
    return out[:, 0] + 1j * out[:, 1]


def test_closed_form_properties(grb, msa, host_tres):
    """Size-independent properties of the contracted vertex: (1) linear in the seven couplings -- the sum of the single-coupling
    amplitudes is the all-couplings amplitude; (2) a Lorentz scalar -- unchanged by a common boost along the beam axis and a
    rotation about it applied to every momentum; (3) symmetric under the exchange of the two proton sides (all seven
    structures are symmetric in pomeron 1 <-> pomeron 2)."""
    z, d = load_case(grb, os.path.join(GOLDEN, "c5_cms_pipi_restensor_f2_allg_born.npz"))
    o, cap = captured(d, z["u"], z["tape"])
    mom = cap[z["W"] != 0][:, :28].copy()
    rp = respars(d)
    full = _raw_sum(host_tres, rp, 1, mom, extended=1)
    # (1) linearity
    parts, scale = np.zeros_like(full), np.zeros(len(full))
    for m in range(7):
        one = respars(d)
        for k in range(7):
            one[0].g[k] = rp[0].g[k] if k == m else 0.0
        one[0].on = (1 << m) if m >= 1 else 0
        a = _raw_sum(host_tres, one, 1, mom, extended=1)
        parts += a
        scale += np.abs(a)
    assert np.max(np.abs(parts - full) / scale) < 1e-14   # (each part is returned rounded to double; the structures partly cancel)
    # (2) boost along z (rapidity 0.7) and rotation about z (1.1 rad) of all seven 4-vectors
    v = mom.reshape(-1, 7, 4).copy()
    ch, sh, c, s = np.cosh(0.7), np.sinh(0.7), np.cos(1.1), np.sin(1.1)
    x, y, zz, e = v[..., 0].copy(), v[..., 1].copy(), v[..., 2].copy(), v[..., 3].copy()
    v[..., 0], v[..., 1] = c * x - s * y, s * x + c * y
    v[..., 2], v[..., 3] = ch * zz + sh * e, sh * zz + ch * e
    moved = _raw_sum(host_tres, rp, 1, v.reshape(-1, 28), extended=1)
    rel = np.abs(moved - full) / np.abs(full)
    # the transformed momenta are rounded to double again: t_k = (p_beam - p_k)^2 of O(6500 GeV) components moves by ~1e-11
    # relative per ulp, and the exponentials amplify it -- a non-invariant formula would be off by O(1)
    assert np.median(rel) < 1e-9 and rel.max() < 1e-5, (np.median(rel), rel.max())
    # (3) side exchange: (pa, p1) <-> (pb, p2)
    w = mom.reshape(-1, 7, 4)[:, [1, 0, 3, 2, 4, 5, 6], :].reshape(-1, 28)
    swapped = _raw_sum(host_tres, rp, 1, w, extended=1)
    assert np.max(np.abs(swapped - full) / np.abs(full)) < 1e-12


@pytest.mark.skipif(not os.path.exists(os.path.join(ROOT, "oracle", "_ref", "gra", "bin", "ref_eval")), reason="compiled reference (oracle/_ref) not present")
def test_oracle_vs_live_reference_screened_C_class(grb, msa):
    """PP[RES+CONTENSOR] in the <C> phase-space class with the screening loop on: the oracle against the live reference on fresh
    points (the reference spends ~0.5 s per screened weight here, hence 40 points; measured max 7e-15 on 151)."""
    from helpers import Eikonal, make_points, ref_eval
    proc = "PP[RES+CONTENSOR]<C> -> pi+ pi- @RES{f0_980:1,f2_1270:1}"
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"), proc, 1)
    u, tape = make_points(40, d.lipsdim, 8, 77)
    r = ref_eval("./tests/processes/ALICECUTS.json", proc, 1, u, tape)
    o = oracle_eval(d, Eikonal(msa[0], msa[1], msa[2]), u, tape)
    nz = r["W"] != 0
    assert nz.sum() > 20 and np.array_equal(o["flags"], r["flags"]) and np.array_equal(o["w"] != 0, nz)
    assert np.max(np.abs(o["w"][nz] - r["W"][nz]) / np.abs(r["W"][nz])) < 1e-12
