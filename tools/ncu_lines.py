"""Developer tool: per-source-line instruction / stall-sample shares of one kernel from an .ncu-rep (cuda,sass view).
Usage: python tools/ncu_lines.py report.ncu-rep [min_percent]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
thr = float(sys.argv[2]) if len(sys.argv) > 2 else 0.5
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, agg, hdr = None, {}, None
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]
        continue
    if r and r[0] == "Line No":
        hdr = r
        continue
    if len(r) < 8 or r[0] in ("Function Name", ""):
        continue
    try:
        ln, inst, samp = int(r[0]), int(r[7]), int(r[4])
    except ValueError:
        continue
    a = agg.setdefault((cur_file, ln), [0, 0, r[1]])
    a[0] += inst
    a[1] += samp
tot = sum(a[0] for a in agg.values()) or 1
tots = sum(a[1] for a in agg.values()) or 1
print("total warp instructions", tot, "samples", tots)
for k, a in sorted(agg.items()):
    if 100 * a[0] / tot >= thr or 100 * a[1] / tots >= thr:
        print(f"{k[0]:20s} {k[1]:5d} inst {100 * a[0] / tot:5.2f}% samp {100 * a[1] / tots:5.2f}%  {a[2].strip()[:100]}")
