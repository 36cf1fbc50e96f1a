"""Reference cross sections for the sigma-parity tests: runs the UNMODIFIED reference CLI (oracle/_ref/gra/bin/gr,
built by oracle/build_ref.sh) on copies of its own cards with only INTEGRALPARAM.VEGAS NCALL/PRECISION raised,
and stores sigma +- err in tests/golden/sigma_ref.json.   python tests/golden/make_sigma_ref.py"""
import json
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from helpers import GOLDEN, REF_GR, REF_TREE  # noqa: E402

# name, card, process, pomloop, NCALL, PRECISION
CASES = [
    ("c2_alice_pipi_con_screened", "tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", "true", 100000, 0.003),
    ("c2_alice_pipi_con_born", "tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", "false", 300000, 0.001),
    ("c1_test_pipi_con_born", "input/test.json", "PP[CON]<F> -> pi+ pi-", "false", 300000, 0.001),
    ("c1_test_pipi_tensor_born", "input/test.json", "PP[CONTENSOR]<F> -> pi+ pi-", "false", 300000, 0.001),
    ("c2_alice_pipi_tensor_screened", "tests/processes/ALICECUTS.json", "PP[CONTENSOR]<F> -> pi+ pi-", "true", 20000, 0.01),
    ("c1_test_pipi_rescon_born", "input/test.json", "PP[RES+CON]<F> -> pi+ pi-", "false", 300000, 0.001),
    ("c3_atlas_pipi_rescon_born", "tests/run_atlas_multi/ATLAS.json", "PP[RES+CON]<F> -> pi+ pi-", "false", 300000, 0.002),
    ("c4_cms_mumu_yycon_born", "tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", "false", 300000, 0.001),
    ("c5_cms_rhorho_4pi_con_born", "tests/processes/CMSCUTS.json",
     "PP[CON]<F> -> rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}", "false", 300000, 0.002),
    ("cc_alice_pipi_con_born", "tests/processes/ALICECUTS.json", "PP[CON]<C> -> pi+ pi-", "false", 300000, 0.001),
    ("cc_alice_pipi_con_screened", "tests/processes/ALICECUTS.json", "PP[CON]<C> -> pi+ pi-", "true", 100000, 0.003),
    # forward N* excitation (NSTARS replaced in the card copy)
    ("c2_alice_pipi_con_nstar1_born", "tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", "false", 300000, 0.002, "NSTARS=1"),
    # (the reference needs > 25 min on 8 cores for this one at PRECISION 0.002: not stored yet, the test skips it)
    ("c4_cms_mumu_yycon_nstar2_born", "tests/run_cms_mumu_low_high/CMS_mumu_lowmass.json", "yy[CON]<F> -> mu+ mu-", "false", 100000, 0.01, "NSTARS=2"),
    # tensor-pomeron resonances (MTensorPomeron::ME3; the reference needs 0.8 ms per amplitude)
    ("c5_cms_pipi_restensor_f2_born", "tests/processes/CMSCUTS.json", "PP[RESTENSOR]<F> -> pi+ pi- @RES{f2_1270:1}", "false", 200000, 0.005),
    ("c5_cms_pipi_rescontensor_born", "tests/processes/CMSCUTS.json", "PP[RES+CONTENSOR]<F> -> pi+ pi- @RES{f0_980:1,f2_1270:1}", "false", 200000, 0.005),
    # another energy (beam energies replaced in the card copy): the energy-scan caller grb_xscan
    ("c2_alice_pipi_con_screened_7tev", "tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", "true", 100000, 0.003, 3500.0),
]


def main():
    out_path = os.path.join(GOLDEN, "sigma_ref.json")
    res = json.load(open(out_path)) if os.path.exists(out_path) else {}
    only = sys.argv[1:]
    for case in CASES:
        name, card, proc, pom, ncall, prec = case[:6]
        if only and name not in only:
            continue
        txt = open(os.path.join(REF_TREE, card)).read()
        if len(case) > 6 and isinstance(case[6], str):      # "KEY=int": an integer key of the card replaced
            key, val = case[6].split("=")
            txt = re.sub(r'("%s"\s*:\s*)\d+' % key, r"\g<1>%s" % val, txt)
        elif len(case) > 6:
            txt = re.sub(r'("ENERGY"\s*:\s*)\[[^\]]*\]', r"\g<1>[%g, %g]" % (case[6], case[6]), txt)
        txt = re.sub(r'("NCALL"\s*:\s*)\d+', r"\g<1>%d" % ncall, txt)
        txt = re.sub(r'("VEGAS"\s*:\s*\{[^}]*?"PRECISION"\s*:\s*)[0-9.eE+-]+', r"\g<1>%g" % prec, txt, flags=re.S)
        tmp = os.path.join(REF_TREE, "output", f"sigma_{name}.json")
        open(tmp, "w").write(txt)
        r = subprocess.run([REF_GR, "-i", tmp, "-p", proc, "-l", pom, "-n", "0", "-h", "0", "-c", str(os.cpu_count())],
                           cwd=REF_TREE, capture_output=True, text=True)
        clean = re.sub(r"\x1b\[[0-9;]*m", "", r.stdout)
        m = re.search(r"cross section\}:\s*\[([0-9.eE+-]+) \+- ([0-9.eE+-]+)\] barn", clean)
        f = re.search(r"Integrand sampling frequency:\s*([0-9.eE+-]+) Hz", clean)
        p = re.search(r"Fiducial cuts passing rate:\s*([0-9.eE+-]+)", clean)
        if not m:
            print(name, "FAILED", clean[-800:], r.stderr[-400:])
            continue
        res[name] = {"card": card, "process": proc, "pomloop": pom == "true", "sigma_barn": float(m.group(1)),
                     "err_barn": float(m.group(2)), "ncall": ncall, "precision": prec,
                     "ref_points_per_s": float(f.group(1)) if f else None, "fid_pass": float(p.group(1)) if p else None,
                     "cores": os.cpu_count()}
        if len(case) > 6 and isinstance(case[6], str):
            res[name]["nstars"] = int(case[6].split("=")[1])
        elif len(case) > 6:
            res[name]["ebeam"] = case[6]
            mt = re.search(r"Total:\s*([0-9.]+) mb.*?Elastic:\s*([0-9.]+) mb.*?Inelastic:\s*([0-9.]+) mb", clean, re.S)
            if mt:
                res[name]["xs_mb"] = [float(mt.group(k)) for k in (1, 2, 3)]
        print(name, res[name], flush=True)
        json.dump(res, open(out_path, "w"), indent=1)


if __name__ == "__main__":
    main()
