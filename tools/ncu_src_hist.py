"""Stall-sample histogram of an ncu report's source page, by source line (needs -lineinfo + --import-source on) or by SASS
region: python tools/ncu_src_hist.py report.ncu-rep [--sass]"""
import csv
import subprocess
import sys

rep = sys.argv[1]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"] + (["--print-source", "sass"] if "--sass" in sys.argv else ["--print-source", "cuda,sass"]),
                     capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
# find header row
hi = next(i for i, r in enumerate(rows) if "# Samples" in r or "Warp Stall Sampling (All Samples)" in r)
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) == len(hdr)]
col = {h: i for i, h in enumerate(hdr)}
isamp = col.get("# Samples", col.get("Warp Stall Sampling (All Samples)"))
isrc = col["Source"]
def I(x):
    try:
        return int(x)
    except ValueError:
        return 0


tot = sum(I(r[isamp]) for r in data)
print("rows", len(data), "samples", tot, "cols", [h for h in hdr[:6]])
stall = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
agg = {}
for r in data:
    for i, h in stall:
        agg[h] = agg.get(h, 0) + I(r[i])
for h, v in sorted(agg.items(), key=lambda x: -x[1])[:8]:
    print(f"  {h:24s} {100 * v / max(tot, 1):5.1f}%")
src_rows = [r for r in data if r[0] not in ("", "-")] if "Line No" in hdr else []
if src_rows and "--sass" not in sys.argv:
    data = src_rows
    isrc = 1
top = sorted(data, key=lambda r: -I(r[isamp]))[:45]
for r in top:
    st = sorted(((I(r[i]), h) for i, h in stall), reverse=True)[:2]
    print(f"{100 * I(r[isamp]) / tot:5.1f}%  {(r[0] + ": " + r[isrc].strip())[:120]:120s} {st[0][1]}:{st[0][0]} {st[1][1]}:{st[1][0]}")
