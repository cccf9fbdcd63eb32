"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel: launches, total, share, us/launch."""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1], errors="ignore")))
hdr = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
H = rows[hdr]
kn, mv = H.index("Kernel Name"), H.index("Metric Value")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hdr + 1:]:
    if len(r) > mv:
        name = r[kn].split("(")[0]
        agg[name][0] += 1
        agg[name][1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
print("| kernel | launches | total us | share | us / launch |\n|---|---:|---:|---:|---:|")
for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f"| `{k}` | {v[0]} | {v[1] / 1e3:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / 1e3 / v[0]:.1f} |")
print(f"\ntotal {tot / 1e3:.1f} us over {sum(v[0] for v in agg.values())} launches")
