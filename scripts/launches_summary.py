"""Per-kernel totals of an `ncu --metrics gpu__time_duration.sum --csv` launch list (cold-cache, serialised times:
compare SHARES with the bench line, not absolutes).   python scripts/launches_summary.py launches.csv [steps] > summary.txt"""
import collections
import csv
import re
import sys

path = sys.argv[1]
steps = float(sys.argv[2]) if len(sys.argv) > 2 else None
with open(path) as f:
    lines = [l for l in f if not l.startswith("==")]
tot = collections.OrderedDict()
n_launch = 0
for row in csv.DictReader(lines):
    if row.get("Metric Name") != "gpu__time_duration.sum":
        continue
    name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("stemseg::<unnamed>::", "")
    v = float(row["Metric Value"].replace(",", ""))
    unit = row["Metric Unit"]
    us = v / 1e3 if unit == "ns" else v if unit == "us" else v * 1e3
    e = tot.setdefault(name, [0, 0.0])
    e[0] += 1
    e[1] += us
    n_launch += 1
total = sum(v for _, v in tot.values())
ours = sum(v for k, (_, v) in tot.items() if not k.startswith("at::"))
print("# %s: %d launches, %.1f us serialised device time (%.1f us in this repo's kernels)" % (path, n_launch, total, ours))
if steps:
    print("# per step (%g steps in the list): %.1f us" % (steps, total / steps))
print("%-64s %7s %12s %7s" % ("kernel", "n", "us", "share"))
for k, (n, v) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print("%-64s %7d %12.1f %6.1f%%" % (k[:64], n, v, 100 * v / total))
