"""Condense an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals.
    python tools/launch_summary.py launches.csv "<command that was profiled>" > profiles/rNN/launches_X.txt"""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
body = rows[rows.index(hdr) + 1:]
kn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = defaultdict(lambda: [0, 0.0])
for r in body:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    v *= {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}.get(r[mu], 1e-6)
    name = r[kn].split("(")[0][:70]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("ncu --metrics gpu__time_duration.sum --clock-control none:", sys.argv[2] if len(sys.argv) > 2 else "")
print(f"total kernel time {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:70s} launches={c:6d} total_ms={t:10.2f} avg_us={t / c * 1e3:9.1f} share={100 * t / tot:5.1f}%")
