"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel launches / total time / share.
usage: summarize_launches.py launches.csv [last_n_launches_fraction]   (keeps the LAST 1/steps of the list = one step)"""
import csv, json, sys, collections

path = sys.argv[1]
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==")]
for r in csv.DictReader(lines):
    if r.get("Metric Name") != "gpu__time_duration.sum":
        continue
    v = float(r["Metric Value"].replace(",", ""))
    unit = r.get("Metric Unit", "ns")
    us = v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3
    rows.append((r["Kernel Name"], us))
n = len(rows) // steps
rows = rows[-n:] if steps > 1 else rows
agg = collections.OrderedDict()
for k, us in rows:
    k = k[:70]
    a = agg.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
total = sum(a[1] for a in agg.values())
out = {"launches": len(rows), "total_us": total,
       "kernels": [{"kernel": k, "launches": a[0], "total_us": round(a[1], 1), "share": round(a[1] / total, 4)}
                   for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1])]}
print(json.dumps(out, indent=1))
