#!/usr/bin/env python
"""Per-kernel-family totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv`).
usage: tools/launch_summary.py launches.csv ["description line"]"""
import collections
import csv
import re
import sys


def main(path, desc=""):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.reader(lines))
    hdr, data = rows[0], rows[1:]
    ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
    tot, cnt = collections.Counter(), collections.Counter()
    for r in data:
        m = re.search(r"(k_[a-z0-9_]+)", r[ik])
        name = m.group(1) if m else r[ik][:28]
        tot[name] += float(r[iv].replace(",", "")) / 1e6
        cnt[name] += 1
    allms = sum(tot.values())
    if desc:
        print(desc)
    print("per-kernel-family totals over the whole run (cold-cache, serialised: compare SHARES with bench.py's roofline.kernels)")
    print("%-28s %8s %12s %8s" % ("kernel", "launches", "total ms", "share"))
    for k, v in tot.most_common():
        print("%-28s %8d %12.2f %7.1f%%" % (k, cnt[k], v, 100 * v / allms))


if __name__ == "__main__":
    main(*sys.argv[1:3])
