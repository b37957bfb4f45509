#!/usr/bin/env python
"""Appends the per-kernel mean time and share-of-one-buffer footer to an ncu launch list (profiles/rNN_launches_*.csv).

  python tools/launch_shares.py gpurun_out/r02_launches_c4.csv <workload> "<command>" > profiles/r02_launches_<workload>.csv
"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    path, workload, command = sys.argv[1], sys.argv[2], sys.argv[3]
    text = open(path).read()
    rows = [r for r in csv.reader(l for l in text.splitlines() if l.startswith('"'))]
    hdr, rows = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    per = OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*$", "", r[ik]).replace("void ", "").replace("(anonymous namespace)::", "").replace("unnamed>::", "").strip()
        per.setdefault(name, []).append(float(r[iv]))
    steady = max(len(v) for v in per.values())
    # a kernel that runs every buffer has (about) as many launches as the most frequent one
    every = {k: v for k, v in per.items() if len(v) >= steady - 1}
    total = sum(sum(v[1:]) / max(len(v) - 1, 1) for v in every.values())
    sys.stdout.write("\n".join(l for l in text.splitlines() if l.startswith('"')) + "\n")
    print("# per-launch device time (ncu, cold-cache, serialised): mean ns (first launch excluded) and SHARE of one steady-state buffer, workload %s" % workload)
    for k in sorted(per):
        v = per[k]
        m = sum(v[1:]) / max(len(v) - 1, 1) if len(v) > 1 else v[0]
        share = "share %.1f%%" % (100.0 * m / total) if k in every else "share (not every buffer)"
        print("#   %-44s launches %2d  mean %9.0f ns  %s" % (k, len(v), m, share))
    print("# one buffer = %.0f ns of kernel time under ncu" % total)
    print("# command: %s" % command)


if __name__ == "__main__":
    main()
