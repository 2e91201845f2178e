"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name (share of the step)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
agg = collections.defaultdict(lambda: [0, 0.0])
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    k = row["Kernel Name"].split("(")[0][:70]
    v = float(row["Metric Value"].replace(",", ""))
    u = row["Metric Unit"]
    v = v / 1e3 if u in ("ns", "nsecond") else (v * 1e3 if u in ("ms", "msecond") else v)
    agg[k][0] += 1
    agg[k][1] += v
tot = sum(v[1] for v in agg.values())
n = sum(v[0] for v in agg.values())
print("total %.1f us in %d launches" % (tot, n))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 45
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:top]:
    print("%10.1f us %5d %5.1f%%  %s" % (v[1], v[0], 100 * v[1] / tot, k))
