"""INTEGRATION.md route A, compiled (oracle/ref_harness/gr_b200.cc): the UNMODIFIED reference parses the steering card and
the binding flattens ITS objects into grb_process_desc.  The bytes must equal what the product's own card parser
(grb_parse_card) produces; on a GPU the integration through that binding equals the Python route bit for bit."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

from helpers import CARDS, REF_TREE
from graniitti_b200 import _abi

GR_B200 = os.path.join(REF_TREE, "bin", "gr_b200")
RHO4PI = "rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}"
CASES = [  # card of this repo, the reference's card, process, pomloop
    ("c1_test13.json", "./input/test.json", "null", -1),
    ("c1_test13.json", "./input/test.json", "PP[CONTENSOR]<F> -> pi+ pi-", -1),
    ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", 1),
    ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON-]<F> -> K+ K-", 0),
    ("c3_atlas.json", "./tests/run_atlas_multi/ATLAS.json", "PP[RES+CON]<F> -> pi+ pi-", 0),
    ("c3_atlas_neg.json", "./tests/run_atlas_multi/ATLAS_NEG.json", "PP[RES+CON]<F> -> pi+ pi-", 0),
    ("c4_cms_mumu_lowmass.json", "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", 0),
    ("c4_cms_mumu_lowmass.json", "./tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[QED]<F> -> mu+ mu-", 0),
    ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CON]<F> -> " + RHO4PI, 0),
    ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CONTENSOR]<F> -> " + RHO4PI, 0),
    ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON]<C> -> pi+ pi-", 1),
    ("c1_test13.json", "./input/test.json", "PP[RES+CON]<F> -> pi+ pi- @R[f0_980]{M:1.02,W:0.09} @R[f2_1270]{JZ0:0.2,JZ2:0.8,W:0.2}", 0),
    # tensor-pomeron resonances: g_Tensor overrides and the decay couplings hel.g_decay_tensor (GDecay; off-shell K+ K- branch for f0(980))
    ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[RESTENSOR]<F> -> pi+ pi- @RES{f2_1270:1} @R[f2_1270]{g0:0.3,g1:1.0,g2:0.7,g3:-0.6,g4:0.5,g5:0.4,g6:-0.2}", 0),
    ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[RES+CONTENSOR]<F> -> K+ K- @RES{f0_980:1,f2_1525:1,f0_1710:1}", 1),
]
pytestmark = pytest.mark.skipif(not os.path.exists(GR_B200), reason="compiled boundary (oracle/_ref/gra/bin/gr_b200) not present")


def ref_desc(card, proc, pom, tmp_path):
    out = str(tmp_path / "desc.bin")
    r = subprocess.run([GR_B200, "desc", card, proc, str(pom), out], cwd=REF_TREE, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    return _abi.ProcessDesc.from_buffer_copy(open(out, "rb").read())


def diff_struct(a, b, path=""):
    """Field-by-field differences of two ctypes structures (for the assertion message)."""
    out = []
    for name, typ in a._fields_:
        x, y = getattr(a, name), getattr(b, name)
        if isinstance(x, C.Structure):
            out += diff_struct(x, y, path + name + ".")
        elif isinstance(x, C.Array):
            for i in range(len(x)):
                if isinstance(x[i], C.Structure):
                    out += diff_struct(x[i], y[i], f"{path}{name}[{i}].")
                elif x[i] != y[i] and not (x[i] != x[i] and y[i] != y[i]):
                    out.append(f"{path}{name}[{i}]: {x[i]!r} != {y[i]!r}")
        elif x != y:
            out.append(f"{path}{name}: {x!r} != {y!r}")
    return out


@pytest.mark.parametrize("card,ref_card,proc,pom", CASES, ids=[f"{c[0][:2]}-{i}" for i, c in enumerate(CASES)])
def test_desc_from_reference_objects_equals_card_parser(grb, tmp_path, card, ref_card, proc, pom):
    d_ref = ref_desc(ref_card, proc, pom, tmp_path)
    d_own = grb.parse_card(os.path.join(CARDS, card), None if proc == "null" else proc, pom)
    diffs = diff_struct(d_ref, d_own)
    assert not diffs, "\n".join(diffs[:40])
    assert bytes(d_ref) == bytes(d_own)


@pytest.mark.gpu
@pytest.mark.parametrize("card,ref_card,proc,pom", [CASES[2], CASES[0], CASES[6]], ids=["c2", "c1", "c4"])
def test_integration_through_the_compiled_boundary(grb, msa, card, ref_card, proc, pom):
    """reference ReadInput -> desc from its objects -> grb_set_eikonal(proc->Eikonal.MSA) -> grb_integrate, against the
    same integration driven through the product's own parser: identical sigma (same Philox stream, same grids)."""
    r = subprocess.run([GR_B200, "run", ref_card, proc, str(pom), str(1 << 20), "4", "0.01", "4711"], cwd=REF_TREE,
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    line = [ln for ln in r.stdout.splitlines() if ln.startswith("RESULT")][-1].split()
    sigma, err = float(line[2]), float(line[4])
    E = grb.Engine(card=os.path.join(CARDS, card), process=None if proc == "null" else proc, pomloop=pom)
    if E.desc.screening:
        E.set_eikonal(msa[0], msa[1], msa[2])
    res = E.integrate(ncall=1 << 20, iters=4, precision=0.01, seed=4711)
    E.close()
    # the accumulators are summed with atomics (order varies run to run): agreement to the summation noise, far below the MC error
    assert abs(res.sigma - sigma) <= 1e-9 * sigma and abs(res.sigma_err - err) <= 1e-6 * err, (res.sigma, sigma, res.sigma_err, err)
