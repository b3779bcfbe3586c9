"""Summarises an ncu --csv launch list (gpu__time_duration.sum per launch) by kernel name."""
import csv, sys, collections, re
rows = collections.OrderedDict()
with open(sys.argv[1]) as f:
    lines = [l for l in f if l.startswith('"')]
rd = csv.DictReader(lines)
for r in rd:
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", r["Kernel Name"])
    v = float(r["Metric Value"].replace(",", ""))
    unit = r["Metric Unit"]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v if unit in ("us", "usecond") else v * 1e3)
    n, t = rows.get(name, (0, 0.0))
    rows[name] = (n + 1, t + us)
tot = sum(t for _, t in rows.values())
for name, (n, t) in sorted(rows.items(), key=lambda kv: -kv[1][1]):
    print(f"{name[:70]:70s} n={n:4d} total={t:10.1f} us share={t / tot:.3f} avg={t / n:8.1f} us")
