"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals (for profiles/)."""
import csv
import re
import sys
from collections import defaultdict

path = sys.argv[1]
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
rd = csv.DictReader(lines)
agg = defaultdict(lambda: [0, 0.0])
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"]).strip()
    name = re.sub(r"^void ", "", name)
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "nsecond": 1e-6}.get(unit, 1e-6)
    agg[name][0] += 1
    agg[name][1] += val * scale
tot = sum(v[1] for v in agg.values())
print(f"# {path}: {sum(v[0] for v in agg.values())} launches, {tot:.2f} ms total (ncu-serialised, cold-cache: compare SHARES)")
print(f"{'kernel':70s} {'launches':>8s} {'ms':>10s} {'share':>7s}")
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {n:8d} {ms:10.3f} {100 * ms / tot:6.1f}%")
