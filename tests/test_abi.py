"""The C-ABI library loads without a GPU and exports every symbol include/grb.h declares."""
import ctypes as C
import os
import re

from helpers import ROOT
from graniitti_b200 import _abi


def declared_symbols():
    txt = open(os.path.join(ROOT, "include", "grb.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(grb_[a-z0-9_]+)\s*\(", txt)))


def test_exports_every_declared_symbol(grb):
    L = grb.lib()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f"libgrb.so does not export {s}"


def test_struct_layout_matches_header(grb):
    L = grb.lib()
    assert L.grb_abi_sizeof(0) == C.sizeof(_abi.ProcessDesc)
    assert L.grb_abi_sizeof(1) == C.sizeof(_abi.Accum)
    assert L.grb_abi_sizeof(2) == C.sizeof(_abi.Resonance)
    assert L.grb_abi_sizeof(3) == C.sizeof(_abi.Node)


def test_no_cpu_fallback_without_device(grb):
    """Without a CUDA device engine creation fails loudly (GRB_ERR_CUDA), never computes on the CPU."""
    import torch
    if torch.cuda.is_available():
        return
    d = grb.parse_card(os.path.join(ROOT, "tests", "cards", "c2_alice.json"))
    try:
        grb.Engine(desc=d)
    except grb.GrbError as e:
        assert e.rc == _abi.GRB_ERR_CUDA
    else:
        raise AssertionError("engine creation must fail without a GPU")


def test_product_does_not_import_oracle():
    """Nothing under graniitti_b200/ may reference oracle/ (the oracle is the checker only)."""
    pkg = os.path.join(ROOT, "graniitti_b200")
    for dp, _, fs in os.walk(pkg):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".cc", ".h")):
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert "oracle" not in txt.lower(), f"{f} mentions the oracle"
