"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name."""
import csv, sys, collections, re
path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
rows = []
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
r = csv.DictReader(lines)
agg = collections.OrderedDict()
for row in r:
    name = row.get("Kernel Name", "")
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    val = float(row["Metric Value"].replace(",", ""))
    unit = row.get("Metric Unit", "ns")
    if unit in ("us", "usecond"): val *= 1e3
    if unit in ("ms", "msecond"): val *= 1e6
    short = re.sub(r"\(.*", "", name)
    short = re.sub(r"^void ", "", short)
    d = agg.setdefault(short, [0, 0.0])
    d[0] += 1; d[1] += val
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':70s} {'launches/step':>13s} {'us/step':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:70]:70s} {v[0]/steps:13.1f} {v[1]/1e3/steps:10.1f} {v[1]/tot:7.3f}")
print(f"{'TOTAL':70s} {sum(v[0] for v in agg.values())/steps:13.1f} {tot/1e3/steps:10.1f}")
