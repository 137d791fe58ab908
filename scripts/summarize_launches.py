"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (shares of the run)."""
import csv
import re
import sys
from collections import defaultdict

rows = []
with open(sys.argv[1], newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3, "s": 1e6}.get(unit, 1e-3)
    name = r["Kernel Name"]
    name = name.replace("(anonymous namespace)::", "")
    name = re.sub(r"\(.*", "", name)          # drop the argument list
    name = re.sub(r"^void ", "", name)
    base = re.sub(r"<.*", "", name)           # drop template arguments ...
    targs = re.findall(r"<(.*)>", name)
    name = base + ("<" + targs[0][:28] + ">" if targs else "")  # ... but keep a short hint of them
    rows.append((name, val * scale))
tot = sum(t for _, t in rows)
agg = defaultdict(lambda: [0, 0.0])
for n, t in rows:
    agg[n][0] += 1
    agg[n][1] += t
print(f"# {len(rows)} launches, total device time {tot/1e3:.2f} ms (serialised under ncu; shares are what matters)")
print(f"{'kernel':60s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{n[:60]:60s} {c:8d} {t:12.1f} {t/c:10.1f} {100*t/tot:6.2f}%")
