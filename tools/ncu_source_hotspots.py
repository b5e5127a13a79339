"""Aggregates `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` per CUDA source line: instructions executed, stall samples and
the dominant stall reasons.  usage: python tools/ncu_source_hotspots.py page.csv [top]"""
import csv
import sys
from collections import defaultdict

path = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
rows = list(csv.reader(open(path)))
cur_file, hdr = None, None
agg = defaultdict(lambda: defaultdict(float))
src = {}
for r in rows:
    if not r:
        continue
    if r[0] in ("File Name", "File Path"):
        cur_file = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        continue
    if r[0] == "Function Name" or hdr is None:
        continue
    d = dict(zip(hdr, r))
    if r[2] != "-":  # SASS rows are attributed through their own table; CUDA rows (Address "-") already hold the per-line sums
        continue
    key = (cur_file, int(r[0]))
    src[key] = r[1]
    for h in hdr[4:]:
        try:
            agg[key][h] += float(d[h])
        except (ValueError, KeyError):
            pass
tot_i = sum(v["Instructions Executed"] for v in agg.values())
tot_s = sum(v["# Samples"] for v in agg.values())
print(f"total warp instructions {tot_i:.0f}, samples {tot_s:.0f}")
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
print("stall totals:", {s: int(sum(v[s] for v in agg.values())) for s in stalls if sum(v[s] for v in agg.values()) > 0.01 * tot_s})
for title, metric in (("by samples", "# Samples"), ("by instructions", "Instructions Executed")):
    print(f"--- top {top} lines {title}")
    for key, v in sorted(agg.items(), key=lambda kv: -kv[1][metric])[:top]:
        st = sorted(((s, v[s]) for s in stalls), key=lambda x: -x[1])[:2]
        print(f"{key[0]}:{key[1]:5d} inst {100 * v['Instructions Executed'] / tot_i:5.1f}% smp {100 * v['# Samples'] / tot_s:5.1f}% thr/inst {v['Thread Instructions Executed'] / max(v['Instructions Executed'], 1):4.1f} "
              f"{st[0][0]}={st[0][1]:.0f} {st[1][0]}={st[1][1]:.0f} | {src[key].strip()[:110]}")
