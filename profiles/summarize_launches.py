#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel (share of the step)."""
import collections
import csv
import re
import sys


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0])
    for row in csv.DictReader(lines):
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e6 if unit in ("ns", "nsecond") else v / 1e3 if unit in ("us", "usecond") else v
        name = re.sub(r"^void |\(.*", "", row["Kernel Name"])
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        agg[name][0] += 1
        agg[name][1] += v
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.3f ms total (cold-cache, serialised: compare shares)" % (path, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s n=%4d %10.3f ms %6.1f%%  avg %8.1f us" % (k[:70], v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0]))


if __name__ == "__main__":
    main(sys.argv[1])
