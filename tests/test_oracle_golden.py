"""Pins the CPU oracle (oracle/grb_oracle.c) against golden vectors produced by the UNMODIFIED reference
compiled in the build container (tests/golden/make_golden.py).  Same libm, same operation order:
the restatement is required to be bit-exact on W and on the flags (HELAS yy[CON]: 1e-13)."""
import glob
import os

import numpy as np
import pytest

from helpers import CARDS, GOLDEN, Eikonal, card_variant, have_ref, make_points, oracle_eval, ref_eval

CARD_OF = {"./tests/processes/ALICECUTS.json": "c2_alice.json", "./input/test.json": "c1_test13.json",
           "./tests/processes/CMSCUTS.json": "c5_cms_4pi.json", "./tests/run_atlas_multi/ATLAS.json": "c3_atlas.json",
           "./tests/run_atlas_multi/ATLAS_NEG.json": "c3_atlas_neg.json",
           "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json": "c4_cms_mumu_lowmass.json"}
FILES = sorted(glob.glob(os.path.join(GOLDEN, "c*.npz")))  # c1..c5: <F> class per BASELINE config, cc_*: <C> class
FILES += sorted(glob.glob(os.path.join(GOLDEN, "late", "c*.npz")))  # generated after the round's GPU budget: tests/test_zz_gpu_late.py


def load_case(grb, path):
    z = np.load(path, allow_pickle=False)
    nstars = int(z["nstars"]) if "nstars" in z.files else 0
    d = grb.parse_card(card_variant(os.path.join(CARDS, CARD_OF[str(z["card"])]), nstars), str(z["process"]), int(z["pomloop"]))
    if "qed_alpha" in z.files:   # PARAM_STRUCTURE.QED_alpha of the tune the golden was generated with
        d.qed_alpha = int(z["qed_alpha"])
    return z, d


@pytest.mark.parametrize("path", FILES, ids=[os.path.basename(f)[:-4] for f in FILES])
def test_oracle_reproduces_reference_bit_exactly(grb, msa, path):
    z, d = load_case(grb, path)
    eik = Eikonal(msa[0], msa[1], msa[2]) if d.screening else None
    o = oracle_eval(d, eik, z["u"], z["tape"])
    kin = (z["flags"] & 1) == 1
    assert np.array_equal(o["flags"] & 1, z["flags"] & 1)
    assert np.array_equal(o["flags"][kin], z["flags"][kin])
    if "yy[CON]" in str(z["process"]):
        # HELAS algebra restated in light-cone components (not in the generated code's operation order): ulp-level differences
        nz = z["W"] != 0
        assert np.array_equal(o["w"] != 0, nz)
        assert np.max(np.abs(o["w"][nz] - z["W"][nz]) / np.abs(z["W"][nz])) < 1e-13
    else:
        # everything else bit for bit -- Regge amplitudes, resonances with the spin machinery, the ladders, and the whole
        # tensor-pomeron code (2xS, 2xV, 2xF, the QED branch, the resonance amplitude ME3 with its rank-6 vertex): explicit tensors
        # and Dirac matrices contracted in FTensor's own order (expression templates evaluate left to right, double contractions
        # run over the contracted pair in natural order from zero, T2 * T2 with the first index fastest), scalar factors applied
        # in the reference's association (FACTOR * (P % u) * (P % v); MMatrix / double is a division)
        assert np.array_equal(o["w"], z["W"]), "oracle differs from the reference"
    assert np.array_equal(o["dump"][kin, 0], z["t1"][kin]) and np.array_equal(o["dump"][kin, 1], z["t2"][kin])
    assert np.array_equal(o["tape_used"], z["tape_used"].astype(np.int32))
    assert (z["W"] != 0).sum() > 50


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("proc,pom,n", [("PP[CON]<F> -> pi+ pi-", 0, 20000), ("PP[CON]<F> -> pi+ pi-", 1, 1500),
                                        ("PP[CON-]<F> -> K+ K-", 0, 5000), ("PP[CON]<F> -> p+ p-", 0, 5000),
                                        ("PP[CON]<C> -> pi+ pi-", 0, 20000), ("PP[CON]<C> -> K+ K-", 1, 300)])
def test_oracle_vs_live_reference(grb, msa, proc, pom, n):
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"), proc, pom)
    u, tape = make_points(n, d.lipsdim, 8, 4242 + n, adapted=True)
    r = ref_eval("./tests/processes/ALICECUTS.json", proc, pom, u, tape)
    o = oracle_eval(d, Eikonal(msa[0], msa[1], msa[2]) if pom else None, u, tape)
    kin = (r["flags"] & 1) == 1
    assert np.array_equal(o["flags"] & 1, r["flags"] & 1) and np.array_equal(o["flags"][kin], r["flags"][kin])
    assert np.array_equal(o["w"], r["W"])


def test_edge_cases(grb, msa):
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"), None, 0)
    # empty batch
    o = oracle_eval(d, None, np.zeros((0, 6)), np.zeros((0, 8)))
    assert len(o["w"]) == 0
    # corners of the unit hypercube: never NaN, flags consistent
    u = np.array([[0.0] * 6, [1.0 - 2 ** -53] * 6, [0.5] * 6, [1.0 - 2 ** -53, 0, 0, 0, 0, 0]])
    o = oracle_eval(d, None, u, np.full((4, 8), 0.5))
    assert np.all(np.isfinite(o["w"]))
    # exhausted tape -> flagged, not crashing
    o = oracle_eval(d, None, u, np.zeros((4, 0)))
    assert np.all(o["tape_used"] < 0)


ORACLE_ONLY = sorted(glob.glob(os.path.join(GOLDEN, "oracle_only", "c*.npz")))


@pytest.mark.parametrize("path", ORACLE_ONLY, ids=[os.path.basename(f)[:-4] for f in ORACLE_ONLY])
def test_oracle_vector_photoproduction_block(grb, msa, path):
    """The 1- block of MTensorPomeron::ME3 (gamma-pomeron fusion with the full proton spinor vertices; reference
    src/MTensorPomeron.cc:273-361) exists in the oracle only -- the engine refuses these resonances (no device amplitude yet).
    The restatement contracts the nine-factor FTensor product pairwise in natural index order and reproduces the reference
    BIT-EXACTLY where the 1- block is alone (all 1116 non-zero weights; the Pauli term of iG_ypp divides by 2 m_p as
    MMatrix::operator/ does -- multiplying by the reciprocal instead left a 1e-8 tail through the photon pole), and inside the
    shipped card's eight-resonance list (PP[RES+CONTENSOR] on input/test.json: 0+, 1-, 2+ blocks and the continuum, helicity
    amplitudes accumulated resonance by resonance in std::map order)."""
    z, d = load_case(grb, path)
    o = oracle_eval(d, None, z["u"], z["tape"])
    assert np.array_equal(o["flags"], z["flags"])
    nz = z["W"] != 0
    assert nz.sum() > 500 and np.array_equal(o["w"] != 0, nz)
    assert np.array_equal(o["w"], z["W"]), "oracle differs from the reference"
