#!/usr/bin/env python
"""Aggregate an `ncu --page source --csv` (SASS view) export: dynamic instruction mix and stall samples per opcode."""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ix = {h: i for i, h in enumerate(hdr)}
ops = collections.Counter()
samp = collections.Counter()
stall = collections.Counter()
tot = 0
tots = 0
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
for r in rows[hi + 1:]:
    if len(r) < len(hdr):
        continue
    m = re.match(r"\s*(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", r[ix["Source"]])
    if not m:
        continue
    op = m.group(1)
    try:
        n = int(r[ix["Instructions Executed"]])
        s = int(r[ix["# Samples"]])
    except ValueError:
        continue
    ops[op] += n
    samp[op] += s
    tot += n
    tots += s
    for c in stall_cols:
        try:
            stall[c] += int(r[ix[c]])
        except ValueError:
            pass
print("total warp instructions %d, samples %d" % (tot, tots))
print("%-10s %12s %6s %8s %6s  samples/instr(rel)" % ("op", "executed", "%", "samples", "%"))
for op, n in ops.most_common(28):
    print("%-10s %12d %5.1f%% %8d %5.1f%%  %.2f" % (op, n, 100.0 * n / tot, samp[op], 100.0 * samp[op] / max(tots, 1),
                                                  (samp[op] / max(tots, 1)) / (n / tot) if n else 0))
print("stall samples:", ", ".join("%s %.1f%%" % (k[6:], 100.0 * v / max(tots, 1)) for k, v in stall.most_common(8)))
