"""Per CUDA source line: warp instructions executed and stall samples, for each profiled launch whose
kernel name matches (needs --import-source on and -lineinfo).
   python scripts/ncu_source_lines.py rep.ncu-rep <kernel substring> [N]"""
import csv
import io
import subprocess
import sys

rep, pat = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                     capture_output=True, text=True).stdout
blocks, cur, path = [], None, ""
for row in csv.reader(io.StringIO(raw)):
    if not row:
        continue
    if row[0] == "File Path":
        path = row[1]
    elif row[0] == "Function Name":
        if cur is None or cur["name"] != row[1] or cur.get("closed"):
            cur = {"name": row[1], "lines": {}}
            blocks.append(cur)
        cur["path"] = path
    elif row[0] == "Line No":
        cur["hdr"] = row
    elif cur is not None and row[0] not in ("", "Line No") and "hdr" in cur:
        h = cur["hdr"]
        i_s, i_ex = h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        if not row[0].isdigit():
            continue
        key = (cur["path"].split("/")[-1], int(row[0]))
        try:
            ex, st = int(row[i_ex] or 0), int(row[i_s] or 0)
        except (ValueError, IndexError):
            continue  # a source line whose quotes broke the CSV row
        e = cur["lines"].setdefault(key, [0, 0, row[1]])
        e[0] += ex
        e[1] += st
for b in blocks:
    if pat not in b["name"]:
        continue
    tot_ex = sum(v[0] for v in b["lines"].values())
    tot_s = sum(v[1] for v in b["lines"].values())
    print("=" * 110)
    print(b["name"][:100], "| warp-instr", tot_ex, "| samples", tot_s)
    for (f, ln), v in sorted(b["lines"].items(), key=lambda kv: -kv[1][0])[:topn]:
        print(f"  {f:14s}:{ln:4d} exec {100.0 * v[0] / max(tot_ex, 1):5.1f}%  stall {100.0 * v[1] / max(tot_s, 1):5.1f}%  {v[2].strip()[:90]}")
