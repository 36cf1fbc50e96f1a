"""Exports golden cases (tests/golden/*.npz) into the flat binary files scripts/gpu_check/gpu_check.c reads, plus the screening
table.  Usage: python scripts/gpu_check/export_cases.py <golden name> ...   (output: scripts/gpu_check/data/)"""
import os
import re
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests"))
from helpers import CARDS, GOLDEN  # noqa: E402
from test_oracle_golden import CARD_OF  # noqa: E402

OUT = os.path.join(ROOT, "scripts", "gpu_check", "data")
os.makedirs(OUT, exist_ok=True)
msa = np.load(os.path.join(GOLDEN, "msa_pp_13000.npy"))
with open(os.path.join(OUT, "msa.bin"), "wb") as f:
    f.write(struct.pack("<q", msa.shape[1]))
    f.write(np.ascontiguousarray(msa, dtype=np.float64).tobytes())
for name in sys.argv[1:]:
    z = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    card = os.path.join(CARDS, CARD_OF[str(z["card"])])
    nstars = int(z["nstars"]) if "nstars" in z.files else 0
    rel = os.path.relpath(card, ROOT)
    if nstars:   # NSTARS edited in the card, like the golden's run card
        txt = re.sub(r'("NSTARS"\s*:\s*)\d+', r"\g<1>%d" % nstars, open(card).read())
        rel = os.path.join("scripts", "gpu_check", "data", "%s_card.json" % name)
        open(os.path.join(ROOT, rel), "w").write(txt)
    u, tape = z["u"], z["tape"]
    with open(os.path.join(OUT, name + ".bin"), "wb") as f:
        f.write(struct.pack("<qqqq", u.shape[0], u.shape[1], tape.shape[1], int(z["pomloop"])))
        f.write(rel.encode().ljust(256, b"\0"))
        f.write(str(z["process"]).encode().ljust(512, b"\0"))
        for a in (u.T, tape.T, z["W"], z["kappa"]):
            f.write(np.ascontiguousarray(a, dtype=np.float64).tobytes())
        f.write(np.ascontiguousarray(z["flags"], dtype=np.uint8).tobytes())
    print(name, u.shape, rel)
