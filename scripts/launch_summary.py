"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel (time, share, count).
   python scripts/launch_summary.py gpurun_out/launches.csv > profiles/x_launches.txt"""
import collections
import csv
import sys

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 5]
hdr = rows[0]
i_name, i_val, i_unit = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[i_val].replace(",", ""))
    v = v / 1e3 if r[i_unit] in ("ns", "nsecond") else v
    agg[r[i_name][:90]][0] += 1
    agg[r[i_name][:90]][1] += v
tot = sum(v[1] for v in agg.values())
print(f"# {len(rows) - 1} launches, {tot / 1e3:.2f} ms total (cold-cache, serialised under ncu: compare SHARES)")
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{t:12.1f} us {100 * t / tot:5.1f}%  n={n:4d}  avg {t / n:9.1f} us  {k}")
