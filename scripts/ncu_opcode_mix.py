"""SASS opcode mix, warp-stall shares and the per-node instruction count of the screening loop from one ncu source
page (`ncu -i x.ncu-rep --page source --csv`, extracted on the GPU box by scripts/gpu_round.sh).
usage: python scripts/ncu_opcode_mix.py gpurun_out/prof_screen_TAG_source.csv > profiles/rNN_k_screen_opcode_mix.txt"""
import collections
import csv
import re
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    kernel, hdr = rows[0][1], rows[1]
    kernel = next((r[1] for r in rows if len(r) >= 2 and len(r) < 8 and ("k_points" in r[1] or "k_screen" in r[1])), kernel)
    data = [r for r in rows[2:] if len(r) == len(hdr) and r != hdr]   # (device functions called from the kernel: own header rows)
    ia, isrc, isamp = hdr.index("Instructions Executed"), hdr.index("Source"), hdr.index("# Samples")
    stall_cols = [(i, h[6:]) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    tot = sum(int(r[ia]) for r in data)
    mix, samp = collections.Counter(), collections.Counter()
    for r in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
        mix[op] += int(r[ia])
        samp[op] += int(r[isamp])
    ts = sum(samp.values())
    print(f"{kernel}: SASS opcode mix of one launch (ncu --set full source page, bench.py --ncalls 4194304)")
    print(f"warp-level instructions executed: {tot}")
    print("opcode   % of executed   % of stall samples")
    for k, v in mix.most_common(24):
        print(f"{k:8s} {100 * v / tot:8.2f} {100 * samp[k] / ts:8.2f}")
    st = collections.Counter()
    for r in data:
        for i, name in stall_cols:
            st[name] += int(r[i] or 0)
    tt = sum(st.values())
    print("stall reasons, % of all samples: " + ", ".join(f"{k} {100 * v / tt:.1f}" for k, v in st.most_common(8)))
    # the loop node: instructions executed once per node by (almost) every warp of a mass class
    mx = max(int(r[ia]) for r in data)
    for lo, hi, name in ((0.9, 1.1, "unequal-mass class (SolvePz3_B)"), (0.4, 0.7, "equal-mass class (SolvePz3_A)")):
        body = [r for r in data if lo * mx < int(r[ia]) <= hi * mx]
        h = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0] for r in body)
        f64 = sum(v for k, v in h.items() if k in ("DMUL", "DADD", "DFMA", "DSETP"))
        print(f"loop node, {name}: {len(body)} instructions = {f64} fp64 (DFMA {h['DFMA']}, DMUL {h['DMUL']}, DADD {h['DADD']}, "
              f"DSETP {h['DSETP']}) + {len(body) - f64} other; issue-cycle model 2*fp64 + other = {2 * f64 + len(body) - f64}")


if __name__ == "__main__":
    main(sys.argv[1])
