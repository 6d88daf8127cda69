#!/usr/bin/env python
"""Summarise an `ncu --page raw --csv` export: one line per kernel launch with the numbers DESIGN.md / profiles/ quote."""
import csv
import sys


def main(path):
    rows = list(csv.reader(open(path)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    stall = [h for h in hdr if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio")]

    def g(r, k, d="-"):
        return r[idx[k]] if k in idx and r[idx[k]] != "" else d

    def f(r, k):
        try:
            return float(g(r, k, "nan").replace(",", ""))
        except ValueError:
            return float("nan")
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0][-40:]
        t = f(r, "gpu__time_duration.sum")
        tu = units[idx["gpu__time_duration.sum"]]
        rd, wr = f(r, "dram__bytes_read.sum"), f(r, "dram__bytes_write.sum")
        ru = units[idx["dram__bytes_read.sum"]]
        sc = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(ru, 1.0)
        tsec = t * {"ms": 1e-3, "us": 1e-6, "s": 1.0, "ns": 1e-9}.get(tu, 1.0)
        issue = f(r, "smsp__issue_active.avg.pct_of_peak_sustained_active")
        fp64 = f(r, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active")
        st = sorted([(f(r, s), s.split("issue_stalled_")[1].split("_per")[0]) for s in stall if g(r, s) not in ("-", "n/a")], reverse=True)[:3]
        print("%-40s t=%8.3f ms  dram R %.2f W %.2f GB -> %.2f TB/s  dram%% %s  issue%% %.1f fp64pipe%% %.1f occ%% %s regs %s smem/blk %s  inst %.0fM  L2hit %s  stalls %s" % (
            name, tsec * 1e3, rd * sc / 1e9, wr * sc / 1e9, (rd + wr) * sc / tsec / 1e12,
            g(r, "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed")[:5], issue, fp64,
            g(r, "sm__warps_active.avg.pct_of_peak_sustained_active")[:5], g(r, "launch__registers_per_thread"),
            g(r, "launch__shared_mem_per_block_dynamic"), f(r, "smsp__inst_executed.sum") / 1e6, g(r, "lts__t_sector_hit_rate.pct")[:5],
            [(round(a, 2), b) for a, b in st]))


if __name__ == "__main__":
    main(sys.argv[1])
