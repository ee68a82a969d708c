"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: python profiles/summarize_launches.py <csv> [first_launch last_launch]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
lo = int(sys.argv[2]) if len(sys.argv) > 2 else 0
hi = int(sys.argv[3]) if len(sys.argv) > 3 else 10 ** 9
hdr, agg, n = None, collections.defaultdict(lambda: [0, 0.0]), 0
for r in rows:
    if len(r) > 10 and r[0] == "ID":
        hdr = r
        continue
    if hdr and len(r) == len(hdr):
        try:
            v = float(r[hdr.index("Metric Value")].replace(",", ""))
        except ValueError:
            continue
        unit = r[hdr.index("Metric Unit")]
        v *= {"ns": 1.0, "us": 1e3, "ms": 1e6}.get(unit, 1.0)
        i = int(r[0])
        if lo <= i <= hi:
            name = re.sub(r"\(.*", "", r[hdr.index("Kernel Name")])[:72]
            agg[name][0] += 1
            agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("launches %d, total %.1f us" % (sum(v[0] for v in agg.values()), tot / 1e3))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-74s n=%4d %9.1f us %5.1f%%" % (k, v[0], v[1] / 1e3, 100 * v[1] / tot))
