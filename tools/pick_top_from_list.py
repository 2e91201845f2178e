"""Reads an `ncu --metrics gpu__time_duration.sum --csv` launch list and prints the indices (among the k_conv_tc / k_wgrad_tc
launches, in launch order) of the longest k_conv_tc and the longest k_wgrad_tc launch: `-s` values for a single-launch capture."""
import csv
import sys

lines = [l for l in open(sys.argv[1]) if l.startswith('"')]
best = {}
i = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = row["Kernel Name"]
    if "k_conv_tc" not in name and "k_wgrad_tc" not in name:
        continue
    v = float(row["Metric Value"].replace(",", ""))
    k = "wgrad" if "wgrad" in name else "conv"
    if k not in best or v > best[k][1]:
        best[k] = (i, v)
    i += 1
print("{} {}".format(best.get("conv", (0,))[0], best.get("wgrad", (0,))[0]))
