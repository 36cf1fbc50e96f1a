"""The final measurement script runs the bench lines first and the ncu captures afterwards, so a bench line carries
executed_* / fp64_slot_frac computed with the PREVIOUS capture's per-unit instruction mix.  This rescales those fields of
profiles/r02_bench_<w>.json to the mix captured in the same run (same formula as bench.py: units x per-unit mix / kernel time;
units and kernel time are the run's own).  usage: python scripts/refresh_mix.py <old r02_sass_mix.json>"""
import json, os, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
old = json.load(open(sys.argv[1]))
new = json.load(open(os.path.join(P, "r02_sass_mix.json")))
ex = lambda m: 2 * m["dfma"] + m["dmul"] + m["dadd"]
sl = lambda m: m["dfma"] + m["dmul"] + m["dadd"] + m.get("dsetp", 0) + m.get("other_fp64", 0)
for w in ["c1", "c2", "c2t", "c4", "c5"]:
    path = os.path.join(P, "r02_bench_%s.json" % w)
    d = json.loads(open(path).read().strip().splitlines()[-1])
    rf = d["roofline"]
    key = "%s:%s" % (w, rf["kernel"])
    if key not in new or "executed_flops_per_unit" not in rf or rf.get("mix_refreshed"):
        continue
    o = old.get(key)
    if not o or abs(ex(o) - rf["executed_flops_per_unit"]) > 1e-6 * ex(o):
        print(w, "old mix does not match the bench line, skipped"); continue
    n = new[key]
    rf["executed_tflops"] *= ex(n) / ex(o)
    rf["executed_frac"] *= ex(n) / ex(o)
    rf["fp64_slot_frac"] *= sl(n) / sl(o)
    rf["executed_flops_per_unit"] = ex(n)
    if rf.get("traffic") and o.get("dram_bytes_per_unit"):
        rf["traffic"] *= n["dram_bytes_per_unit"] / o["dram_bytes_per_unit"]
    rf["mix_refreshed"] = "executed_* and fp64_slot_frac rescaled to the ncu capture of the same run (scripts/refresh_mix.py)"
    open(path, "w").write(json.dumps(d) + "\n")
    print(w, key, "executed_frac %.3f fp64_slot_frac %.3f (%d -> %d fp64 instructions per unit)" % (rf["executed_frac"], rf["fp64_slot_frac"], sl(o), sl(n)))
