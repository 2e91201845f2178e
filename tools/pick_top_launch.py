"""Reads an `ncu --page raw --csv` export of the k_conv_tc / k_wgrad_tc launches of one step and prints, per kernel,
the launch with the largest gpu__time_duration plus (last line) the two launch indices to re-capture with source."""
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hdr = rows[0]
col = {h: i for i, h in enumerate(hdr)}
name_c = col["Kernel Name"]
dur_c = col["gpu__time_duration.sum"]
units = rows[1]
best = {}
for idx, r in enumerate(rows[2:]):
    if len(r) <= dur_c:
        continue
    try:
        d = float(r[dur_c].replace(",", ""))
    except ValueError:
        continue
    if units[dur_c] in ("ns", "nsecond"):
        d /= 1e3
    elif units[dur_c] in ("ms", "msecond"):
        d *= 1e3
    k = "wgrad" if "wgrad" in r[name_c] else "conv"
    if k not in best or d > best[k][1]:
        best[k] = (idx, d, r[col["Grid Size"]] if "Grid Size" in col else "")
for k, v in best.items():
    print("{}: launch #{} of the filtered list, {:.1f} us, grid {}".format(k, v[0], v[1], v[2]))
print("{} {}".format(best.get("conv", (0,))[0], best.get("wgrad", (0,))[0]))
