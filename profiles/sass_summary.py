#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that prove tcgen05 / TMEM / TMA use (B200_PROFILING.md):
    python profiles/sass_summary.py [libmsx.so] > profiles/sass_r2_summary.txt
UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA tensor load / store /
reduce, UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, HMMA = mma.sync, SYNCS = mbarrier ops."""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                         "musicstyletransfer_b200", "libmsx.so")
MNEMONICS = ["UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "UTCBAR", "HMMA", "SYNCS", "FFMA"]
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        if op in MNEMONICS:
            counts[cur][op] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(counts), capture_output=True, text=True).stdout.splitlines()
print("%-110s " % "kernel" + " ".join("%8s" % m for m in MNEMONICS))
tot = collections.Counter()
for (k, c), name in sorted(zip(counts.items(), demangle), key=lambda kv: kv[1]):
    tot.update(c)
    name = re.sub(r"\(anonymous namespace\)::", "", name)
    name = name.split("(")[0]
    print("%-110s " % name[:110] + " ".join("%8d" % c[m] for m in MNEMONICS))
print("%-110s " % "TOTAL" + " ".join("%8d" % tot[m] for m in MNEMONICS))
