#!/usr/bin/env python
"""Format the LAST n launches of an `ncu --metrics gpu__time_duration.sum --csv` log (one forward, in launch order)
followed by the aggregation by kernel.

    python tools/last_step.py gpurun_out/launches.csv 90 > profiles/rNN_launches_one_step.txt
"""
import collections
import csv
import re
import sys


def main():
    path, n = sys.argv[1], int(sys.argv[2])
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    gi, bi = hdr.index("Grid Size"), hdr.index("Block Size")
    rows = rows[1:][-n:]
    print("# one forward step (bf16, batch 16) in launch order: us, grid, block, kernel   [ncu gpu__time_duration.sum, "
          "cold-cache, serialised]")
    agg, tot = collections.OrderedDict(), 0.0
    for r in rows:
        name = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("<unnamed>::", "").replace("unnamed>::", "").replace("sj::", "")
        v = float(r[vi].replace(",", "")) / (1000.0 if r[ui] == "ns" else 1.0)
        print(f"{v:8.1f} {r[gi]:>14s} {r[bi]:>13s} {name[-80:]}")
        a = agg.setdefault(name, [0.0, 0])
        a[0] += v
        a[1] += 1
        tot += v
    print(f"# total {tot:.1f} us in {len(rows)} launches")
    print("# by kernel:")
    for name, (v, c) in sorted(agg.items(), key=lambda x: -x[1][0]):
        print(f"# {v:8.1f} us {100 * v / tot:5.1f}% {c:3d}  {name[-80:]}")


if __name__ == "__main__":
    main()
