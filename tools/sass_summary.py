#!/usr/bin/env python
"""Count the Blackwell-specific SASS mnemonics per kernel of the built library (evidence that the hot path is tcgen05 /
TMA / TMEM code; the .so itself is git-ignored).

    python tools/sass_summary.py > profiles/rNN_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "strajnet_b200", "lib", "libstrajnet_b200.so")
PAT = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTCCP", "SYNCS", "HMMA.16816", "LDGSTS", "ELECT"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|sj::|void ", "", name)
            cur = name.split("(")[0].strip() or m.group(1)
            per.setdefault(cur, collections.Counter())
            continue
        if cur is None:
            continue
        for p in PAT:
            if re.search(r"\b" + re.escape(p), line):
                per[cur][p] += 1
    tot = collections.Counter()
    print(f"# {os.path.relpath(SO, ROOT)}: Blackwell-specific SASS mnemonics per kernel (cuobjdump -sass)")
    print("# UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor load, "
          "SYNCS = mbarrier, HMMA.16816 = mma.sync, LDGSTS = cp.async")
    for k, c in per.items():
        if not c:
            continue
        print(f"{k[:90]:90s} " + " ".join(f"{p}={c[p]}" for p in PAT if c[p]))
        tot.update(c)
    print("# total: " + " ".join(f"{p}={tot[p]}" for p in PAT if tot[p]))


if __name__ == "__main__":
    sys.exit(main())
