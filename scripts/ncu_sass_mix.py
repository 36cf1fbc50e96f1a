"""Executed fp64 work of one kernel per unit (loop node for k_screen, point for k_points) from one `ncu --set full`
capture: thread-level executed instruction counts of the source page summed per opcode, divided by the units of the
profiled launch (taken from the bench JSON line the profiled run printed), plus the DRAM bytes per unit of the raw page.
Writes / updates profiles/r02_sass_mix.json (read by bench.py for `executed_*` and `fp64_slot_frac`).
usage: python scripts/ncu_sass_mix.py <key e.g. c2:k_screen> <source.csv> <raw.csv> <ncu_run.log>"""
import collections
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main(key, src, raw, log):
    rows = list(csv.reader(open(src)))
    kernel, hdr = rows[0][1], rows[1]
    kernel = next((r[1] for r in rows if len(r) >= 2 and len(r) < 8 and ("k_points" in r[1] or "k_screen" in r[1])), kernel)
    data = [r for r in rows[2:] if len(r) == len(hdr) and r != hdr]   # (device functions called from the kernel: own header rows)
    it, isrc = hdr.index("Thread Instructions Executed"), hdr.index("Source")
    iw = hdr.index("Instructions Executed")
    th, wp = collections.Counter(), collections.Counter()
    for r in data:
        op = re.sub(r"^@!?U?P\d+\s+", "", r[isrc].strip()).split()[0].split(".")[0]
        th[op] += int(r[it])
        wp[op] += int(r[iw])
    line = [ln for ln in open(log) if ln.startswith('{"metric"')][-1]
    d = json.loads(line)
    pts = d["run"]["points_per_step_per_gpu"]
    rf = d["roofline"]
    if rf["kernel"] == "k_screen":
        units = pts * d["run"]["accepted_fraction"] * rf["executed_nodes_per_accepted_point"]
        per_launch_units = units / max(1, round(pts / 8388608 + 0.499))     # one launch per 2^23-point chunk
        unit = "loop node"
    else:
        units = per_launch_units = float(pts)
        unit = "point"
    rr = list(csv.reader(open(raw)))
    rh, rv = rr[0], rr[2]
    g = lambda k: float(rv[rh.index(k)]) if k in rh else None
    dram = (g("dram__bytes_read.sum") + g("dram__bytes_write.sum")) * 1e6
    tot_th, tot_wp = sum(th.values()), sum(wp.values())
    out = {"kernel": kernel, "unit": unit, "units_in_profiled_launch": per_launch_units,
           "dfma": th["DFMA"] / per_launch_units, "dmul": th["DMUL"] / per_launch_units, "dadd": th["DADD"] / per_launch_units,
           "dsetp": th["DSETP"] / per_launch_units, "other_fp64": (th["DMNMX"] + th["F2F"] * 0) / per_launch_units,
           "all_instructions": tot_th / per_launch_units,
           "dram_bytes_per_unit": dram / (per_launch_units if unit == "point" else per_launch_units / rf["executed_nodes_per_accepted_point"]),
           "dram_bytes_per": "point" if unit == "point" else "work-list item",
           "fp64_pipe_pct": g("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
           "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
           "registers": g("launch__registers_per_thread"), "duration_ms_under_ncu": g("gpu__time_duration.sum"),
           "source": "ncu --set full source page (thread instructions executed per opcode) of %s" % os.path.basename(src)}
    path = os.path.join(ROOT, "profiles", "r02_sass_mix.json")
    allm = json.load(open(path)) if os.path.exists(path) else {}
    allm[key] = out
    json.dump(allm, open(path, "w"), indent=1)
    ex = 2 * out["dfma"] + out["dmul"] + out["dadd"]
    print(f"{key}: {kernel}: per {unit}: DFMA {out['dfma']:.1f} DMUL {out['dmul']:.1f} DADD {out['dadd']:.1f} DSETP {out['dsetp']:.1f} "
          f"-> {ex:.0f} executed flop, {out['all_instructions']:.0f} instructions; fp64 pipe {out['fp64_pipe_pct']:.1f} %, issue active "
          f"{out['issue_active_pct']:.1f} %, {out['registers']:.0f} registers, DRAM {out['dram_bytes_per_unit']:.0f} B per {out['dram_bytes_per']}")


if __name__ == "__main__":
    main(*sys.argv[1:5])
