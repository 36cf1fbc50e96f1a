"""Shared test plumbing: loads the CPU oracle (oracle/liboracle.so), drives the compiled reference
(oracle/_ref/gra/bin/ref_eval, only present where /root/reference was available at build time) and the
product C-ABI library.  The oracle is the CHECKER; nothing in graniitti_b200/ imports it."""
import ctypes as C
import os
import struct
import subprocess
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
import sys

sys.path.insert(0, ROOT)
from graniitti_b200 import _abi  # noqa: E402

ORACLE_DIR = os.path.join(ROOT, "oracle")
REF_BIN = os.path.join(ORACLE_DIR, "_ref", "gra", "bin", "ref_eval")
REF_GR = os.path.join(ORACLE_DIR, "_ref", "gra", "bin", "gr")
REF_TREE = os.path.join(ORACLE_DIR, "_ref", "gra")
GOLDEN = os.path.join(ROOT, "tests", "golden")
CARDS = os.path.join(ROOT, "tests", "cards")


class OracleEikonal(C.Structure):
    _fields_ = [("n", C.c_int32), ("pad_", C.c_int32), ("min", C.c_double), ("max", C.c_double),
                ("x", C.POINTER(C.c_double)), ("re", C.POINTER(C.c_double)), ("im", C.POINTER(C.c_double))]


def build_oracle(force=False):
    so = os.path.join(ORACLE_DIR, "liboracle.so")
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith(".c")]
    deps = srcs + [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".h", ".inc"))] + [os.path.join(ROOT, "include", "grb.h")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in deps):
        subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-std=gnu99", "-o", so]
                              + srcs + ["-lm"])
    return so


_oracle = None


def load_oracle():
    global _oracle
    if _oracle is None:
        lib = C.CDLL(build_oracle())
        lib.grb_oracle_eval.restype = C.c_int
        lib.grb_oracle_eval.argtypes = [C.POINTER(_abi.ProcessDesc), C.POINTER(OracleEikonal), C.c_int64,
                                        C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                        C.c_void_p, C.c_void_p]
        _oracle = lib
    return _oracle


class Eikonal:
    """Screening amplitude table (kt^2 grid, complex A_eik) + keeps the numpy buffers alive."""

    def __init__(self, x, re, im, xmin=1e-6, xmax=25.0):
        self.x = np.ascontiguousarray(x, dtype=np.float64)
        self.re = np.ascontiguousarray(re, dtype=np.float64)
        self.im = np.ascontiguousarray(im, dtype=np.float64)
        self.min, self.max = float(xmin), float(xmax)

    def cstruct(self):
        dp = C.POINTER(C.c_double)
        return OracleEikonal(len(self.x) - 1, 0, self.min, self.max, self.x.ctypes.data_as(dp),
                             self.re.ctypes.data_as(dp), self.im.ctypes.data_as(dp))


def oracle_eval(desc, eik, u, tape):
    """u [n][dim], tape [n][dtape] (point-major). Returns dict(w, flags, dump, tape_used)."""
    lib = load_oracle()
    u = np.ascontiguousarray(u, dtype=np.float64)
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    n, dim = u.shape
    dt = tape.shape[1]
    w = np.zeros(n)
    flags = np.zeros(n, dtype=np.uint8)
    dump = np.zeros((n, _abi.GRB_DUMP_LEN))
    used = np.zeros(n, dtype=np.int32)
    e = eik.cstruct() if eik is not None else None
    rc = lib.grb_oracle_eval(C.byref(desc), C.byref(e) if e is not None else None, n, u.ctypes.data, dim,
                             tape.ctypes.data, dt, w.ctypes.data, flags.ctypes.data, dump.ctypes.data,
                             used.ctypes.data)
    if rc != 0:
        raise RuntimeError(f"grb_oracle_eval rc={rc}")
    return dict(w=w, flags=flags, dump=dump, tape_used=used)


# ---- compiled reference (only where oracle/_ref exists) ------------------------------------------------

REC_FIELDS = ["W", "kin_ok", "fid_ok", "veto_ok", "amp_ok", "tape_used", "t1", "t2", "s1", "s2", "m2", "Y",
              "Pt", "t_hat", "u_hat", "nhamp", "kappa", "DW"]
REC_NHAMP, REC_NPART = 16, 9


def have_ref():
    return os.path.exists(REF_BIN)


def ref_eval(card, proc, pomloop, u, tape, kappa=False):
    """Run the UNMODIFIED reference's EventWeight on (u, tape). card is relative to the reference tree."""
    u = np.ascontiguousarray(u, dtype=np.float64)
    tape = np.ascontiguousarray(tape, dtype=np.float64)
    n, du = u.shape
    dt = tape.shape[1]
    with tempfile.TemporaryDirectory() as td:
        fin, fout = os.path.join(td, "in.bin"), os.path.join(td, "out.bin")
        with open(fin, "wb") as f:
            f.write(struct.pack("<qqq", n, du, dt))
            f.write(np.hstack([u, tape]).tobytes())
        r = subprocess.run([REF_BIN, "eval", card, proc or "null", str(int(pomloop)), fin, fout,
                            str(int(kappa))], cwd=REF_TREE, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_eval failed: " + r.stderr[-2000:])
        raw = open(fout, "rb").read()
    nn, rec = struct.unpack("<qq", raw[:16])
    a = np.frombuffer(raw[16:], dtype=np.float64).reshape(nn, rec)
    out = {k: a[:, i].copy() for i, k in enumerate(REC_FIELDS)}
    o = len(REC_FIELDS)
    out["hamp"] = (a[:, o:o + 2 * REC_NHAMP:2] + 1j * a[:, o + 1:o + 2 * REC_NHAMP:2]).copy()
    o += 2 * REC_NHAMP
    out["p"] = a[:, o:o + 4 * REC_NPART].reshape(nn, REC_NPART, 4).copy()
    out["flags"] = (out["kin_ok"].astype(np.uint8) * 1 + out["fid_ok"].astype(np.uint8) * 2 +
                    out["veto_ok"].astype(np.uint8) * 4 + out["amp_ok"].astype(np.uint8) * 8)
    return out


def ref_msa(card, proc):
    with tempfile.TemporaryDirectory() as td:
        fout = os.path.join(td, "msa.bin")
        r = subprocess.run([REF_BIN, "msa", card, proc or "null", fout], cwd=REF_TREE, capture_output=True,
                           text=True)
        if r.returncode != 0:
            raise RuntimeError("ref_eval msa failed: " + r.stderr[-2000:])
        raw = np.fromfile(fout, dtype=np.float64)
    hdr, tab = raw[:8], raw[8:].reshape(-1, 3)
    return hdr, Eikonal(tab[:, 0], tab[:, 1], tab[:, 2], hdr[1], hdr[2])


def ref_info(card, proc, pomloop):
    r = subprocess.run([REF_BIN, "info", card, proc or "null", str(int(pomloop))], cwd=REF_TREE,
                       capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("ref_eval info failed: " + r.stderr[-2000:])
    return [ln[5:] for ln in r.stdout.splitlines() if ln.startswith("INFO ")]


# ---- point generators ---------------------------------------------------------------------------------

def make_points(n, dim, dtape, seed, adapted=True):
    """Synthetic phase-space points. `adapted` concentrates the two log-pt coordinates where the proton
    pt is O(0.05-2.5 GeV) (u0,u1 in [0.85,1]): flat u is dominated by pt ~ 1e-6 GeV."""
    rng = np.random.default_rng(seed)
    u = rng.random((n, dim))
    if adapted:
        u[:, 0] = 0.85 + 0.15 * u[:, 0]
        u[:, 1] = 0.85 + 0.15 * u[:, 1]
    tape = rng.random((n, dtape))
    return u, tape


def card_variant(card_path, nstars):
    """A copy of a test card with PROCESSPARAM.NSTARS replaced (what editing the card does for `gr -i`)."""
    import re
    if not nstars:
        return card_path
    txt = re.sub(r'("NSTARS"\s*:\s*)\d+', r"\g<1>%d" % nstars, open(card_path).read())
    out = os.path.join(tempfile.gettempdir(), "grb_%s_nstars%d_%d.json" % (os.path.basename(card_path)[:-5], nstars, os.getpid()))
    open(out, "w").write(txt)
    return out
