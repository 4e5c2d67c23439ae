#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel name.

    python tools/agg_launches.py gpurun_out/launches.csv [skip_first_n]
"""
import collections
import csv
import re
import sys


def main():
    path = sys.argv[1]
    skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi = hdr.index("Grid Size")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows[1 + skip:]:
        n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("unnamed>::", "").replace("sj::", "")
        v = float(r[vi].replace(",", ""))
        if r[ui] == "ns":
            v /= 1000
        a = agg.setdefault(n, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    for n, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"{v:9.1f} us {100 * v / tot:5.1f}% {c:4d}  {n[-90:]}")
    print(f"{tot:9.1f} us total, {sum(c for _, c in agg.values())} launches")


if __name__ == "__main__":
    main()
