"""Event record + writer: the record the product builds from SoA momenta against the reference's
own MProcess::CommonRecord (src/MProcess.cc:1588-1706) on the same events, and the three output formats."""
import os
import re
import struct
import subprocess

import numpy as np
import pytest

from helpers import CARDS, REF_BIN, REF_TREE, have_ref, make_points

RHO4PI = "rho(770)0 > {pi+ pi-} rho(770)0 > {pi+ pi-}"


def ref_records(card, proc, pom, u, tape, tmp_path):
    fin, fout = str(tmp_path / "in.bin"), str(tmp_path / "rec.txt")
    with open(fin, "wb") as f:
        f.write(struct.pack("<qqq", len(u), u.shape[1], tape.shape[1]))
        f.write(np.hstack([u, tape]).tobytes())
    r = subprocess.run([REF_BIN, "record", card, proc, str(pom), fin, fout], cwd=REF_TREE, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    events, cur = [], None
    for ln in open(fout):
        f = ln.split()
        if f[0] == "EVENT":
            cur = dict(k=int(f[1]), P=[], V=[], W=float(f[4]))
            events.append(cur)
        elif f[0] == "P":
            cur["P"].append(dict(id=int(f[1]), pid=int(f[2]), status=int(f[3]), parent=int(f[4]), p=tuple(float(x) for x in f[5:9])))
        elif f[0] == "V":
            nin = int(f[2])
            cur["V"].append(dict(id=int(f[1]), inc=[int(x) for x in f[3:3 + nin]], n_out=int(f[3 + nin])))
    return events


@pytest.mark.skipif(not have_ref(), reason="compiled reference (oracle/_ref) not present")
@pytest.mark.parametrize("card,ref_card,proc,ntape", [
    ("c2_alice.json", "./tests/processes/ALICECUTS.json", "PP[CON]<F> -> pi+ pi-", 8),
    ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CON]<F> -> " + RHO4PI, 24),
    ("c1_test13.json", "./input/test.json", "PP[RES+CON]<F> -> pi+ pi-", 8),
    ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[CON]<C> -> pi+ pi- K+ K- pi+ pi-", 16),
    ("c5_cms_4pi.json", "./tests/processes/CMSCUTS.json", "PP[RES+CONTENSOR]<F> -> pi+ pi- @RES{f0_980:1,f2_1270:1}", 8)],
    ids=["pipi", "rhorho_4pi", "rescon", "ladder_6", "rescontensor"])
def test_record_equals_reference_common_record(grb, tmp_path, card, ref_card, proc, ntape):
    d = grb.parse_card(os.path.join(CARDS, card), proc, 0)
    u, tape = make_points(400 if d.n_legs < 6 else 1500, d.lipsdim, ntape, 5, adapted=True)
    if d.n_legs >= 4:
        u[:, 4:3 + d.n_legs] *= 0.3      # intermediate kt of the ladder below ~0.75 GeV
    ev = ref_records(ref_card, proc, 0, u, tape, tmp_path)
    assert len(ev) > 50
    for e in ev:
        P = e["P"]
        # the product's input: central system, the two forward systems, then the decay-tree nodes in grb_process_desc order
        # (first-level legs, then breadth first) -- taken from the reference's record of this event
        by_parent = {}
        for q in P:
            by_parent.setdefault(q["parent"], []).append(q)
        dec_vertex = {v["inc"][0]: v["id"] for v in e["V"][4:]}          # mother particle id -> its decay vertex
        legs = by_parent[-4]
        nodes, queue = list(legs), list(legs)
        while queue:
            m = queue.pop(0)
            kids = by_parent.get(dec_vertex.get(m["id"], 0), []) if m["id"] in dec_vertex else []
            nodes += kids
            queue += kids
        assert [n["pid"] for n in nodes] == [d.nodes[i].pdg for i in range(d.n_nodes)]
        mom = np.array([P[6]["p"], P[1]["p"], P[4]["p"]] + [n["p"] for n in nodes])
        parts, verts = grb.event_record(d, mom)
        assert len(parts) == len(P) and len(verts) == len(e["V"])
        for a, b in zip(parts, P):
            assert (a["id"], a["pid"], a["status"], a["parent"]) == (b["id"], b["pid"], b["status"], b["parent"]), (a, b)
            # beams, forward systems, central system and decay products: the momenta handed in; propagators: beam - forward
            assert np.allclose(a["p"], b["p"], rtol=0, atol=1e-9 * 6500.0), (a, b)
        for a, b in zip(verts, e["V"]):
            assert a["id"] == b["id"] and a["inc"] == b["inc"] and a["n_out"] == b["n_out"], (a, b)
        assert [p["status"] for p in parts[:7]] == [4, 1, 81, 4, 1, 81, 81] and [p["pid"] for p in parts[:7]] == [2212, 2212, 99, 2212, 2212, 99, 90]


def synthetic_events(grb, n=3000):
    d = grb.parse_card(os.path.join(CARDS, "c5_cms_4pi.json"), "PP[CON]<F> -> " + RHO4PI, 0)
    rng = np.random.default_rng(3)
    npart = 3 + d.n_nodes
    mom = rng.normal(size=(npart, 4, n))
    mom[:, 3] = np.sqrt((mom[:, :3] ** 2).sum(axis=1) + 0.14 ** 2)
    mom[1, 2] += 6499.0; mom[1, 3] = np.sqrt((mom[1, :3] ** 2).sum(axis=0) + 0.938272081 ** 2)
    mom[2, 2] -= 6499.0; mom[2, 3] = np.sqrt((mom[2, :3] ** 2).sum(axis=0) + 0.938272081 ** 2)
    return d, mom, np.abs(rng.normal(size=(2, n)))


@pytest.mark.parametrize("fmt", ["hepmc3", "hepmc2", "hepevt"])
def test_writer_formats(grb, tmp_path, fmt):
    d, mom, wts = synthetic_events(grb)
    n = mom.shape[2]
    path = tmp_path / ("ev." + fmt)
    assert grb.write_events(d, path, mom, wts, fmt=fmt, sigma_pb=2.3e6, sigma_err_pb=1.1e3, nthreads=4) == n
    txt = open(path).read()
    ev = [ln for ln in txt.splitlines() if ln.startswith("E ")]
    assert len(ev) == n
    if fmt == "hepmc3":
        assert txt.startswith("HepMC::Version 3.02.00\nHepMC::Asciiv3-START_EVENT_LISTING\nT GRANIITTI (TUNE0)\\|1.051\\|Generator\n")
        assert txt.rstrip().endswith("HepMC::Asciiv3-END_EVENT_LISTING")
        # first event: 13 particles (7 + 2 rho + 4 pi), 6 vertices; vertex lines only where several particles come in
        blk = txt.split("\nE ")[1].split("\n")
        assert blk[0] == "0 6 13" and blk[1] == "U GEV MM" and blk[2] == "W 1.0000000000000000e+00"
        assert blk[3] == "A 0 GenCrossSection 2.30000000e+06 1.10000000e+03 -1 -1"
        P = [ln.split() for ln in blk if ln.startswith("P ")]
        V = [ln for ln in blk if ln.startswith("V ")]
        assert V == ["V -3 0 [3,6]"] and len(P) == 13
        assert [(int(p[1]), int(p[2]), int(p[3]), int(p[9])) for p in P] == [
            (1, 0, 2212, 4), (2, 1, 2212, 1), (3, 1, 99, 81), (4, 0, 2212, 4), (5, 4, 2212, 1), (6, 4, 99, 81), (7, -3, 90, 81),
            (8, 7, 113, 2), (9, 8, 211, 1), (10, 8, -211, 1), (11, 7, 113, 2), (12, 11, 211, 1), (13, 11, -211, 1)]
        assert float(P[7][4]) == mom[3, 0, 0] and float(P[12][7]) == mom[8, 3, 0]          # %.16e round-trips a double
    if fmt == "hepmc2":
        assert txt.startswith("HepMC::Version 3.02.00\nHepMC::IO_GenEvent-START_EVENT_LISTING\nE 0 ")
        assert len([ln for ln in txt.splitlines() if ln.startswith("V ")]) == 6 * n and len([ln for ln in txt.splitlines() if ln.startswith("P ")]) == 13 * n
    if fmt == "hepevt":
        assert len(txt.splitlines()) == n * (1 + 2 * 13)
    # the thread count does not change the file
    p1 = tmp_path / ("ev1." + fmt)
    grb.write_events(d, p1, mom, wts, fmt=fmt, sigma_pb=2.3e6, sigma_err_pb=1.1e3, nthreads=1)
    assert open(p1).read() == txt
    with pytest.raises(grb.GrbError):
        grb.write_events(d, tmp_path / "x", mom, wts, fmt="root")


def test_decay_vertex_positions(grb):
    """rho(770) lifetime hbar / width = 4.4e-24 s: decay vertices sit O(1e-12 mm) from the origin along the rho's flight."""
    d, mom, _ = synthetic_events(grb, 4)
    parts, verts = grb.event_record(d, mom[:, :, 0], lifetime_u=np.full(16, 0.5))
    v5 = verts[4]
    assert v5["inc"] == [8] and 0 < v5["pos"][3] < 1e-9
    rho = parts[7]["p"]
    assert np.allclose(np.array(v5["pos"][:3]) / np.linalg.norm(v5["pos"][:3]), np.array(rho[:3]) / np.linalg.norm(rho[:3]))


@pytest.mark.gpu
def test_generated_events_to_file(grb, msa, tmp_path):
    """Device generation (GRB_MODE_GENERATE) -> grb_write_engine_events: every accepted event lands in the file, momentum
    is conserved in every record, and the write rate is reported."""
    import time
    E = grb.Engine(card=os.path.join(CARDS, "c2_alice.json"), pomloop=1)
    E.set_eikonal(msa[0], msa[1], msa[2])
    r = E.integrate(ncall=1 << 20, iters=3, precision=0.01, seed=5)
    xmat = np.frombuffer(r.xmat, dtype=np.float64)[:128 * 6].reshape(128, 6).copy()
    acc = E.vegas_iterate(xmat, 0, 1 << 22, 99, 7, mode=2, maxf=r.maxf)
    t0 = time.perf_counter()
    n = E.write_events(tmp_path / "gen.hepmc3", fmt="hepmc3", sigma_pb=r.sigma * 1e12, sigma_err_pb=r.sigma_err * 1e12)
    dt = time.perf_counter() - t0
    assert n == int(acc.n_accepted) and n > 1000 and E.events_dropped() == 0
    lines = open(tmp_path / "gen.hepmc3").read().splitlines()
    assert sum(ln.startswith("E ") for ln in lines) == n
    P = np.array([[float(x) for x in ln.split()[4:8]] for ln in lines if ln.startswith("P ")]).reshape(n, 9, 4)
    assert np.max(np.abs(P[:, 1] + P[:, 2] - P[:, 0])) < 1e-9 and np.max(np.abs(P[:, 7] + P[:, 8] - P[:, 6])) < 1e-9
    assert np.max(np.abs((P[:, 1] + P[:, 4] + P[:, 6])[:, :3])) < 1e-6
    print(f"wrote {n} events in {dt:.3f} s = {n / dt:.3e} events/s")
    E.close()
