"""Host-side steering: card + process string + tune -> grb_process_desc (mirrors `gr -i card -p proc -l b`)."""
import os

import pytest

from helpers import CARDS, ROOT, card_variant, have_ref, ref_info
from graniitti_b200 import _abi

REF = "/root/reference"


def test_c2_card_flattening(grb):
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"))
    assert d.proc == _abi.GRB_PROC_PP_CON and d.ps_class == 0 and d.screening == 1
    assert d.lipsdim == 6 and d.n_legs == 2 and d.n_nodes == 2 and d.n_tape == 4
    assert (d.loop_nkt, d.loop_nphi) == (18, 15)          # ND = 1 -> 15+3, 12+3 (MEikonal.h:44-51)
    assert d.s == 169000000.0 and d.sqrt_s == 13000.0
    assert d.nodes[0].pdg == 211 and d.nodes[1].pdg == -211
    assert d.gcuts.forward_pt_max == 2.5 and d.gcuts.Y_max == 0.9
    assert d.fcuts.active == 1 and d.fcuts.pt_min == 0.15 and d.fcuts.usercuts == 0
    assert d.regge.c[0] == 35.00184 and d.regge.n[0] == 1 and d.regge.n[1] == 0
    assert abs(d.soft.g3P - 0.15 * 8.2) < 1e-15
    assert d.S_factor == 1.0


def test_overrides_like_gr_cli(grb):
    card = os.path.join(CARDS, "c2_alice.json")
    d = grb.parse_card(card, "PP[CON-]<F> -> K+ K-", 0)
    assert d.proc == _abi.GRB_PROC_PP_CON_MINUS and d.screening == 0
    assert d.nodes[0].pdg == 321 and d.regge.c[0] == 30.35376
    d = grb.parse_card(card, "PP[CON]<F> -> pi+ pi- @PDG[211]{M:0.2} @OFFSHELL:3")
    assert d.nodes[0].mass == 0.2 and d.nodes[1].mass == 0.2 and d.offshell == 3.0
    d = grb.parse_card(card, "PP[CON]<F> &> pi+ pi-")
    assert d.ps_active == 0


def test_continuum_class_flattening(grb):
    """<C> (MContinuum): daughters' rapidity range from GENCUTS.<C>.Rap, default kt range [0, 10 / legs],
    3 Nf - 4 = 8 integration dimensions (reference src/MGraniitti.cc:688-745, src/MContinuum.cc:64-91)."""
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"), "PP[CON]<C> -> pi+ pi-", 1)
    assert d.ps_class == _abi.GRB_CLASS_C and d.lipsdim == 8 and d.n_legs == 2 and d.screening == 1
    assert (d.gcuts.rap_min, d.gcuts.rap_max) == (-0.9, 0.9) and (d.gcuts.kt_min, d.gcuts.kt_max) == (0.0, 5.0)
    assert d.gcuts.forward_pt_max == 2.5
    d = grb.parse_card(os.path.join(CARDS, "c4_cms_mumu_lowmass.json"), "yy[CON]<C> -> mu+ mu-", 0)
    assert d.gcuts.kt_max == d.sqrt_s / 2.0      # QED / EW processes: kinematic maximum


def test_forward_excitation_flattening(grb):
    """NSTARS = 1, 2: one / two more integration dimensions, the technical forward pt bound of SetTechnicalBoundaries
    (50 / 100 GeV) and gra::GENCUT's default Xi range [0, 1] when the <F> block has no "Xi"."""
    for ns, ptmax in ((1, 50.0), (2, 100.0)):
        d = grb.parse_card(card_variant(os.path.join(CARDS, "c2_alice.json"), ns), "PP[CON]<F> -> pi+ pi-", 0)
        assert d.excitation == ns and d.lipsdim == 6 + ns and d.gcuts.forward_pt_max == ptmax
        assert (d.gcuts.XI_min, d.gcuts.XI_max) == (0.0, 1.0)


def test_cascade_tree_flattening(grb):
    d = grb.parse_card(os.path.join(CARDS, "c5_cms_4pi.json"))
    assert d.n_legs == 2 and d.n_nodes == 6 and d.n_final == 4 and d.S_factor == 2.0 and d.n_tape == 16
    assert [d.nodes[i].pdg for i in range(6)] == [113, 113, 211, -211, 211, -211]
    assert (d.nodes[0].first_child, d.nodes[1].first_child) == (2, 4)
    assert d.regge.c[0] == 25.0  # PARAM_RHO


def test_error_behaviour(grb, tmp_path):
    card = os.path.join(CARDS, "c2_alice.json")
    with pytest.raises(grb.GrbError) as e:
        grb.parse_card(card, "gg[CON]<F> -> pi+ pi-")
    assert e.value.rc == _abi.GRB_ERR_UNSUPPORTED
    with pytest.raises(grb.GrbError) as e:
        grb.parse_card(card, "PP[CON]<F> -> pi+ nosuchparticle")
    assert e.value.rc == _abi.GRB_ERR_PARSE
    with pytest.raises(grb.GrbError) as e:
        grb.parse_card(card, "PP[CON]<F> -> rho(770)0 > {pi+ pi-")
    assert e.value.rc == _abi.GRB_ERR_PARSE
    bad = tmp_path / "bad.json"
    bad.write_text('{"GENERALPARAM": {"MODELPARAM": "TUNE0",}, "PROCESSPARAM": {')
    with pytest.raises(grb.GrbError) as e:
        grb.parse_card(str(bad))
    assert e.value.rc == _abi.GRB_ERR_PARSE


@pytest.mark.skipif(not os.path.isdir(REF), reason="/root/reference not present")
def test_builtin_data_equals_reference_files(grb):
    """Built-in TUNE0 + PDG subset flatten to the same bytes as the reference's own modeldata files."""
    pairs = [("c2_alice.json", "tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", 1),
             ("c2_alice.json", "tests/processes/ALICECUTS.json", "PP[CON]<F> -> K+ K-", 0),
             ("c5_cms_4pi.json", "tests/processes/CMSCUTS.json",
              "PP[CON]<F> -> rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}", 0),
             ("c4_cms_mumu_lowmass.json", "tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", None, -1)]
    for mine, theirs, proc, pom in pairs:
        a = grb.parse_card(os.path.join(CARDS, mine), proc, pom)
        b = grb.parse_card(os.path.join(REF, theirs), proc, pom, REF)
        assert bytes(a) == bytes(b), (mine, proc)


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
def test_flattening_agrees_with_reference_objects(grb):
    """Facts printed by the reference's own parsed MProcess object equal the flattened description."""
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"))
    info = {ln.split()[0]: ln.split()[1:] for ln in ref_info("./tests/processes/ALICECUTS.json",
                                                             "PP[CON]<F> -> pi+ pi-", 1) if not ln.startswith("leg")}
    assert float(info["s"][0]) == d.s and float(info["sqrt_s"][0]) == d.sqrt_s
    assert float(info["pbeam1"][0]) == d.pbeam1_pz and float(info["pbeam2"][0]) == d.pbeam2_pz
    assert float(info["beam_mass"][0]) == d.beam1_mass
    assert int(info["lipsdim"][0]) == d.lipsdim
    assert [int(info["loop"][1]), int(info["loop"][3])] == [d.loop_nkt, d.loop_nphi]


def test_resonance_overrides_before_the_decay_coupling(grb):
    """@R[name]{M, W, g<n>, JZ<n>}: applied to the freshly read resonance before the decay-side computation, so that
    g_decay = sqrt(BR W / PS(M)) uses the overridden values (reference src/MGraniitti.cc:481-578, MProcess.cc:676-712).
    Facts of the reference's own objects: tests/golden/res_override_ref.json (make_res_override_ref.py)."""
    import json
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "res_override_ref.json")))
    d = grb.parse_card(os.path.join(CARDS, "c1_test13.json"), ref["process"], 0)
    by_pdg = {d.res[i].pdg: d.res[i] for i in range(d.n_res)}
    assert len(by_pdg) == len(ref["res"]) == 8
    for name, r in ref["res"].items():
        o = by_pdg[r["pdg"]]
        assert o.mass == r["mass"] and o.width == r["width"] and o.BR == r["BR"], name
        assert abs(o.g_decay - r["gdecay"]) <= 4e-16 * r["gdecay"], (name, o.g_decay, r["gdecay"])
        n = o.spinX2 + 1
        if o.spinX2:
            for a in range(n):
                for b in range(n):
                    assert abs(o.rho_re[a * 5 + b] - r["rho_re"][a * n + b]) < 1e-16 and o.rho_im[a * 5 + b] == r["rho_im"][a * n + b], name
    assert by_pdg[9010221].mass == 1.02 and by_pdg[225].width == 0.2 and by_pdg[113].g_Tensor[0] == 1.5
    # target errors are the reference's
    for bad in ("@R[nosuch]{M:1.0}", "@R[]{M:1.0}", "@R[f0_980,f2_1270]{M:1.0}"):
        with pytest.raises(grb.GrbError):
            grb.parse_card(os.path.join(CARDS, "c1_test13.json"), "PP[RES+CON]<F> -> pi+ pi- " + bad, 0)
