#!/usr/bin/env python
"""Parse `nvcc -Xptxas -v` output (stdin): one line per kernel with registers, spills, static smem."""
import re
import subprocess
import sys

txt = sys.stdin.read()
pat = sys.argv[1] if len(sys.argv) > 1 else ""
cur = None
for line in txt.splitlines():
    m = re.search(r"Compiling entry function '(\S+)'", line)
    if m:
        cur = m.group(1)
        continue
    if cur is None:
        continue
    m = re.search(r"(\d+) bytes stack frame, (\d+) bytes spill stores, (\d+) bytes spill loads", line)
    if m:
        stack, ss, sl = m.groups()
        continue
    m = re.search(r"Used (\d+) registers", line)
    if m:
        try:
            name = subprocess.run(["c++filt", cur], capture_output=True, text=True).stdout.strip().split("(")[0]
        except Exception:
            name = cur
        if pat in name:
            print("%-48s regs %3s  stack %4s  spill st/ld %s/%s" % (name[-48:], m.group(1), stack, ss, sl))
        cur = None
