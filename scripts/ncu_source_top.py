"""Top SASS instructions by warp-stall samples for each kernel in an .ncu-rep (needs --import-source on).
   python scripts/ncu_source_top.py rep.ncu-rep [N]"""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
topn = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
blocks, cur = [], None
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "Kernel Name":
        cur = {"name": row[1], "hdr": None, "rows": []}
        blocks.append(cur)
    elif row[0] == "Address":
        cur["hdr"] = row
    elif cur is not None and cur["hdr"] is not None:
        cur["rows"].append(row)
for b in blocks:
    h = b["hdr"]
    i_src, i_s, i_ex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
    tot = sum(int(r[i_s] or 0) for r in b["rows"])
    totex = sum(int(r[i_ex] or 0) for r in b["rows"])
    print("=" * 100)
    print(b["name"], "| samples", tot, "| warp-instructions executed", totex, "| SASS lines", len(b["rows"]))
    rows = sorted(enumerate(b["rows"]), key=lambda kv: -int(kv[1][i_s] or 0))[:topn]
    for idx, r in sorted(rows):
        print(f"  #{idx:5d} {100.0 * int(r[i_s] or 0) / max(tot, 1):6.2f}%  exec {int(r[i_ex] or 0):9d}  {r[i_src].strip()[:100]}")
