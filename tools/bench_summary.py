"""Developer tool: one-screen summary of a bench.py JSON line.  Usage: python tools/bench_summary.py file.json"""
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print("value", d["value"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["value"], "cpu", (d.get("cpu_baseline") or {}).get("value"), "launches", d["gpu_launches"], "ms/step", d["ms_per_step"])
print("per_class", d.get("per_class"))
for k, v in (d.get("configs") or {}).items():
    if "error" in v:
        print(f"{k:16s} ERROR {v['error']}")
        continue
    print(f"{k:16s} {v['value']:9.2f} {v['unit']:12s} frac {v['roofline']['frac']:.4f} verified {v['verified']} ratio {v.get('compression_ratio')} enc {v.get('gpu_encode_gbs_raw_in')}")
