"""Condense an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch
list into per-kernel totals (time, launches and -- when the DRAM counters were collected -- bytes moved).
    python tools/launch_summary.py launches.csv "<command that was profiled>" [out.json] > profiles/rNN/launches_X.txt
The optional JSON gets {"total_ms", "launches", "dram_bytes", "kernels": {name: {...}}} for bench.py's roofline.traffic."""
import csv
import json
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
hdr = next(r for r in rows if "Kernel Name" in r)
body = rows[rows.index(hdr) + 1:]
kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
TIME = {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3, "nsecond": 1e-6, "usecond": 1e-3, "msecond": 1.0, "second": 1e3}
BYTES = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
agg = defaultdict(lambda: [0, 0.0, 0.0, 0.0])            # launches, ms, bytes read, bytes written
have_dram = False
for r in body:
    try:
        v = float(r[mv].replace(",", ""))
    except ValueError:
        continue
    name = r[kn].split("(")[0][:70]
    metric = r[mn]
    if metric.startswith("gpu__time_duration"):
        agg[name][0] += 1
        agg[name][1] += v * TIME.get(r[mu], 1e-6)
    elif metric.startswith("dram__bytes_read"):
        agg[name][2] += v * BYTES.get(r[mu], 1.0)
        have_dram = True
    elif metric.startswith("dram__bytes_write"):
        agg[name][3] += v * BYTES.get(r[mu], 1.0)
        have_dram = True
tot = sum(v[1] for v in agg.values())
tot_b = sum(v[2] + v[3] for v in agg.values())
metrics = "gpu__time_duration.sum" + (",dram__bytes_read.sum,dram__bytes_write.sum" if have_dram else "")
print(f"ncu --metrics {metrics} --clock-control none:", sys.argv[2] if len(sys.argv) > 2 else "")
print(f"total kernel time {tot:.1f} ms over {sum(v[0] for v in agg.values())} launches"
      + (f"; DRAM traffic {tot_b / 1e9:.2f} GB (read + write, all kernels)" if have_dram else ""))
for name, (c, t, br, bw) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    line = f"{name:70s} launches={c:6d} total_ms={t:10.2f} avg_us={t / max(c, 1) * 1e3:9.1f} share={100 * t / tot:5.1f}%"
    if have_dram:
        line += f" dram_read_GB={br / 1e9:9.3f} dram_write_GB={bw / 1e9:9.3f}"
    print(line)
if len(sys.argv) > 3:
    json.dump({"total_ms": tot, "launches": sum(v[0] for v in agg.values()), "dram_bytes": tot_b if have_dram else None,
               "kernels": {k: {"launches": v[0], "ms": v[1], "dram_read": v[2], "dram_write": v[3]}
                           for k, v in agg.items()}}, open(sys.argv[3], "w"), indent=1)
