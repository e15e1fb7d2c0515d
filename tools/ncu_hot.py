"""top SASS lines of an `ncu --page source --csv` dump by warp-stall samples, with the dominant stall reason and the
source-level region they sit in (nearest preceding labelled instruction is shown as context)"""
import csv
import sys

path, topn = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 25
rows = list(csv.reader(open(path)))
hdr = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
body = rows[2:]
stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
tot = sum(int(r[idx["# Samples"]] or 0) for r in body)
print(f"total samples {tot}")
ranked = sorted(range(len(body)), key=lambda i: -int(body[i][idx["# Samples"]] or 0))[:topn]
for i in sorted(ranked):
    r = body[i]
    n = int(r[idx["# Samples"]] or 0)
    top = sorted(((int(r[idx[s]] or 0), s) for s in stalls), reverse=True)[:2]
    print(f"{i:5d} {100.0 * n / tot:5.1f}%  {r[idx['Source']].strip()[:70]:70s} {top[0][1]}={top[0][0]} {top[1][1]}={top[1][0]}")
