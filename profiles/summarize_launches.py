#!/usr/bin/env python
"""Summarise an `ncu --metrics gpu__time_duration.sum[,dram__bytes_read.sum,dram__bytes_write.sum] --csv` launch list by
kernel: launches, total time, share of the step and, when the DRAM counters were collected, bytes and GB/s per launch."""
import collections
import csv
import re
import sys


def _scale(v, unit):
    v = float(v.replace(",", ""))
    if unit in ("ns", "nsecond"):
        return v / 1e6
    if unit in ("us", "usecond"):
        return v / 1e3
    if unit in ("ms", "msecond"):
        return v
    return v * {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)


def main(path):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    agg = collections.defaultdict(lambda: [0, 0.0, 0.0])      # launches, ms, dram bytes
    for row in csv.DictReader(lines):
        name = re.sub(r"^void |\(.*", "", row["Kernel Name"])
        name = re.sub(r"\(anonymous namespace\)::", "", name)
        metric = row.get("Metric Name", "gpu__time_duration.sum")
        v = _scale(row["Metric Value"], row["Metric Unit"])
        if metric.startswith("gpu__time_duration"):
            agg[name][0] += 1
            agg[name][1] += v
        elif metric.startswith("dram__bytes"):
            agg[name][2] += v
    tot = sum(v[1] for v in agg.values())
    print("# %s: %d launches, %.3f ms total (cold-cache, serialised: compare shares)" % (path, sum(v[0] for v in agg.values()), tot))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        extra = ""
        if v[2] > 0:
            extra = "  dram %8.1f MB/launch %7.0f GB/s" % (v[2] / v[0] / 1e6, v[2] / (v[1] * 1e-3) / 1e9)
        print("%-70s n=%4d %10.3f ms %6.1f%%  avg %8.1f us%s" % (k[:70], v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0], extra))


if __name__ == "__main__":
    main(sys.argv[1])
