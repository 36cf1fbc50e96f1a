"""Eikonal table construction (reference src/MEikonal.cc:129-436): proton opacity
MBT and screening amplitude MSA.

CPU: the C restatement (oracle/grb_oracle_eikonal.c) is pinned BIT-EXACTLY to the reference's own tables
(tests/golden/msa_pp_13000.npy, eikonal_ref.npz; made by make_golden.py / make_eikonal_golden.py from the
compiled reference) and to the reference's printed cross sections; the cache file names and the ASCII format of
the host library are pinned to the reference's files.
GPU: the device builder against the same goldens.  It sums the same 10 001 products per row in a different
order (strided partial sums + fma) with CUDA's sincos, so it is not bit-exact: tolerance 2e-13 of the table's
peak magnitude per entry (measured ~3e-15), and the screened weights computed with either table agree to 1e-10."""
import ctypes as C
import os

import numpy as np
import pytest

from helpers import CARDS, GOLDEN, Eikonal, build_oracle, make_points, oracle_eval

REF = np.load(os.path.join(GOLDEN, "eikonal_ref.npz"))


def oracle_tables(desc, s, nb=10000, nk=10000):
    lib = C.CDLL(build_oracle())
    mbt, msa, xs = np.zeros((3, nb + 1)), np.zeros((3, nk + 1)), np.zeros(3)
    rc = lib.grb_oracle_build_eikonal(C.byref(desc.soft), C.c_double(s), nb, nk,
                                      *[C.c_void_p(a.ctypes.data) for a in (*mbt, *msa, xs)])
    assert rc == 0
    return mbt, msa, xs * 1e3


@pytest.fixture(scope="module")
def desc(grb):
    return grb.parse_card(os.path.join(CARDS, "c2_alice.json"))


def test_oracle_tables_bit_exact_13tev(desc, msa):
    mbt, sa, xs = oracle_tables(desc, desc.s)
    assert np.array_equal(sa, msa), "MSA differs from the reference's table"
    assert np.array_equal(mbt[:, ::10], REF["mbt_13000"]), "MBT differs from the reference's cache file"
    assert np.array_equal(np.round(xs, 3), REF["xs_mb_13000"]), xs      # the reference prints %0.3f mb


def test_oracle_tables_bit_exact_7tev(desc):
    mbt, sa, xs = oracle_tables(desc, 7000.0 ** 2)
    assert np.array_equal(mbt[:, ::10], REF["mbt_7000"])
    assert np.array_equal(sa[:, ::10], REF["msa_7000"])
    assert np.array_equal(np.round(xs, 3), REF["xs_mb_7000"]), xs


def test_cache_names_match_reference(grb, desc):
    assert grb.eikonal_cache_names(desc) == (str(REF["mbt_name_13000"]), str(REF["msa_name_13000"]))
    d7 = grb.parse_card(os.path.join(CARDS, "c2_alice.json"))
    d7.s, d7.sqrt_s = 7000.0 ** 2, 7000.0
    assert grb.eikonal_cache_names(d7) == (str(REF["mbt_name_7000"]), str(REF["msa_name_7000"]))


def test_table_file_round_trip(grb, msa, tmp_path):
    """%.15g CSV rows as IArray1D::WriteArray writes them; the in-memory tables ARE 15-digit values, so the round
    trip is the identity, and a truncated file is refused like the reference's 'Corrupted file'."""
    p = tmp_path / "MSA_test"
    grb.write_eikonal_file(p, msa[0], msa[1], msa[2])
    assert np.array_equal(grb.read_eikonal_file(p, msa.shape[1]), msa)
    first = open(p).readline().strip().split(",")
    assert first[0] == "1e-06" and len(first) == 3
    lines = open(p).read().splitlines()
    open(p, "w").write("\n".join(lines[:-3]) + "\n")
    with pytest.raises(grb.GrbError):
        grb.read_eikonal_file(p, msa.shape[1])


@pytest.mark.gpu
def test_device_tables_vs_reference(grb, msa):
    E = grb.Engine(card=os.path.join(CARDS, "c2_alice.json"))
    t = E.build_eikonal()
    assert np.array_equal(t["msa"][0], msa[0])
    assert np.array_equal(t["mbt"][0, ::10], REF["mbt_13000"][0])
    for k in (1, 2):
        assert np.abs(t["msa"][k] - msa[k]).max() < 2e-13 * np.abs(msa[1:]).max()
        assert np.abs(t["mbt"][k, ::10] - REF["mbt_13000"][k]).max() < 2e-13 * np.abs(REF["mbt_13000"][1:]).max()
    assert np.array_equal(np.round(t["xs_mb"], 3), REF["xs_mb_13000"])
    # the engine now screens with its own table: weights vs the oracle screening with the reference's table
    u, tape = make_points(4096, 6, 8, 31, adapted=True)
    w, fl, _ = E.eval(u.T.copy(), tape.T.copy())
    o = oracle_eval(E.desc, Eikonal(msa[0], msa[1], msa[2]), u, tape)
    nz = o["w"] != 0
    assert np.array_equal(w != 0, nz) and nz.sum() > 500
    rel = np.abs(w[nz] - o["w"][nz]) / o["w"][nz]
    assert np.median(rel) < 1e-12 and np.mean(rel > 1e-10) < 2e-3
    E.close()


@pytest.mark.gpu
def test_device_tables_other_energy_and_sizes(grb):
    """sqrt(s) = 7 TeV against the reference's tables, on a coarser grid (every 10th row is the same integral)."""
    d = grb.parse_card(os.path.join(CARDS, "c2_alice.json"))
    d.s, d.sqrt_s = 7000.0 ** 2, 7000.0
    E = grb.Engine(desc=d)
    t = E.build_eikonal(10000, 1000)
    ref_b, ref_k = REF["mbt_7000"], REF["msa_7000"]
    for k in (1, 2):
        assert np.abs(t["mbt"][k, ::10] - ref_b[k]).max() < 2e-13 * np.abs(ref_b[1:]).max()
        assert np.abs(t["msa"][k] - ref_k[k]).max() < 1e-9 * np.abs(ref_k[1:]).max()
    assert np.array_equal(np.round(t["xs_mb"], 3), REF["xs_mb_7000"])
    E.close()
