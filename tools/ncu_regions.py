"""Developer tool: instruction / stall-sample shares of one kernel per source REGION (function-level line ranges found
by scanning the source for markers).  Usage: python tools/ncu_regions.py report.ncu-rep file.cu 'name:lo-hi,...'"""
import csv, subprocess, sys
rep, regions = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
cur_file, agg = None, {}
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if len(r) < 8 or r[0] in ("Function Name", "", "Line No"): continue
    try: ln, inst, samp = int(r[0]), int(r[7]), int(r[4])
    except ValueError: continue
    a = agg.setdefault((cur_file, ln), [0, 0]); a[0] += inst; a[1] += samp
tot = sum(a[0] for a in agg.values()) or 1; tots = sum(a[1] for a in agg.values()) or 1
regs = []
for spec in regions.split(","):
    name, rng = spec.split(":"); f, rr = (rng.split("@") + ["decode_flaglz.cu"])[:2] if "@" in rng else (rng, "decode_flaglz.cu")
    lo, hi = f.split("-"); regs.append((name, rr, int(lo), int(hi)))
used = set()
print(f"total warp instructions {tot}  samples {tots}")
for name, f, lo, hi in regs:
    i = sum(a[0] for k, a in agg.items() if k[0] == f and lo <= k[1] <= hi)
    s = sum(a[1] for k, a in agg.items() if k[0] == f and lo <= k[1] <= hi)
    used |= {k for k in agg if k[0] == f and lo <= k[1] <= hi}
    print(f"{name:28s} {f}:{lo}-{hi}  inst {100*i/tot:6.2f}%  samples {100*s/tots:6.2f}%")
oi = sum(a[0] for k, a in agg.items() if k not in used); os_ = sum(a[1] for k, a in agg.items() if k not in used)
print(f"{'(other files / lines)':28s} inst {100*oi/tot:6.2f}%  samples {100*os_/tots:6.2f}%")
by_file = {}
for k, a in agg.items():
    if k in used: continue
    b = by_file.setdefault(k[0], [0, 0]); b[0] += a[0]; b[1] += a[1]
for f, b in sorted(by_file.items(), key=lambda x: -x[1][0])[:6]:
    print(f"    other: {f:24s} inst {100*b[0]/tot:6.2f}%  samples {100*b[1]/tots:6.2f}%")
