#!/usr/bin/env python
"""profiles/traffic.json from the raw ncu table of ONE coarse step (the file bench.py's roofline.traffic comes from):

    ncu --profile-from-start off --clock-control none --csv --log-file profiles/r2_ncu_raw_512.csv --metrics \\
        gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,\\
lts__t_sector_hit_rate.pct,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,\\
smsp__inst_executed.sum  python tools/profile_step.py 512
    python tools/ncu_traffic.py profiles/r2_ncu_raw_512.csv 512 <commit> > profiles/traffic.json

Per kernel family (template arguments folded): launches, device time, DRAM bytes read + written, bytes per zone-update, L2 hit rate,
FP64-pipe busy time (pct of peak x duration), issue utilisation; and the totals of the hydro pipeline (construct_ctu_hydro_source)
and of the whole step.  ncu serialises the launches and replays them with cold caches, so the TIMES here are for shares only."""
import csv
import json
import re
import sys

PIPELINE = ("k_ctoprim", "k_src_to_prim", "k_flat_div", "k_trace", "k_riemann1", "k_trans", "k_final", "k_update", "k_store_flux")


def main(path, n, commit):
    zones = float(n) ** 3
    rows = list(csv.reader(l for l in open(path) if l.startswith('"')))
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    launches = {}
    order = []
    for r in rows[1:]:
        if len(r) < len(hdr):
            continue
        lid = r[ix["ID"]]
        if lid not in launches:
            launches[lid] = {"name": r[ix["Kernel Name"]]}
            order.append(lid)
        v = r[ix["Metric Value"]].replace(",", "")
        try:
            v = float(v)
        except ValueError:
            continue
        u = r[ix["Metric Unit"]]
        scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1e-3, "us": 1e-6, "ns": 1e-9, "s": 1.0,
                 "msecond": 1e-3, "usecond": 1e-6, "nsecond": 1e-9, "second": 1.0}.get(u, 1.0)
        launches[lid][r[ix["Metric Name"]]] = v * scale
    fam = {}
    for lid in order:
        L = launches[lid]
        name = re.sub(r"^void\s+", "", L["name"]).split("(")[0]
        name = re.sub(r"^ctu::", "", name)
        key = re.sub(r"<.*", "", name)
        f = fam.setdefault(key, {"launches": 0, "time_ms": 0.0, "dram_read_bytes": 0.0, "dram_write_bytes": 0.0, "fp64_busy_ms": 0.0,
                                 "l2_hit_w": 0.0, "issue_w": 0.0, "warp_inst": 0.0, "occ_w": 0.0})
        t = L.get("gpu__time_duration.sum", 0.0)
        f["launches"] += 1
        f["time_ms"] += t * 1e3
        f["dram_read_bytes"] += L.get("dram__bytes_read.sum", 0.0)
        f["dram_write_bytes"] += L.get("dram__bytes_write.sum", 0.0)
        f["fp64_busy_ms"] += t * 1e3 * L.get("sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active", 0.0) / 100.0
        f["l2_hit_w"] += t * L.get("lts__t_sector_hit_rate.pct", 0.0)
        f["issue_w"] += t * L.get("smsp__issue_active.avg.pct_of_peak_sustained_active", 0.0)
        f["occ_w"] += t * L.get("sm__warps_active.avg.pct_of_peak_sustained_active", 0.0)
        f["warp_inst"] += L.get("smsp__inst_executed.sum", 0.0)
    out = {"source": path, "commit": commit, "workload": "one coarse step of 3-D Sedov %d^3, HLLC (tools/profile_step.py)" % n, "zones": zones,
           "algorithmic_bytes_per_zone": 128, "kernels": {}}
    tot = {"pipeline": [0.0, 0.0, 0.0], "step": [0.0, 0.0, 0.0]}
    for k, f in fam.items():
        t = f["time_ms"]
        b = f["dram_read_bytes"] + f["dram_write_bytes"]
        out["kernels"][k] = {
            "launches": f["launches"], "time_ms_under_ncu": round(t, 3), "dram_read_GB": round(f["dram_read_bytes"] / 1e9, 3),
            "dram_write_GB": round(f["dram_write_bytes"] / 1e9, 3), "bytes_per_zone": round(b / zones, 1),
            "dram_TBps": round(b / (t * 1e-3) / 1e12, 3) if t else None, "l2_hit_pct": round(f["l2_hit_w"] / (t * 1e-3), 1) if t else None,
            "fp64_pipe_busy_ms": round(f["fp64_busy_ms"], 3), "fp64_pipe_pct": round(100 * f["fp64_busy_ms"] / t, 1) if t else None,
            "issue_pct": round(f["issue_w"] / (t * 1e-3), 1) if t else None, "occupancy_pct": round(f["occ_w"] / (t * 1e-3), 1) if t else None,
            "warp_instructions_M": round(f["warp_inst"] / 1e6, 1)}
        for grp in (("pipeline", "step") if k.startswith(PIPELINE) else ("step",)):
            tot[grp][0] += b
            tot[grp][1] += t
            tot[grp][2] += f["fp64_busy_ms"]
    out["pipeline_dram_bytes_per_zone"] = round(tot["pipeline"][0] / zones, 1)
    out["step_dram_bytes_per_zone"] = round(tot["step"][0] / zones, 1)
    out["fp64_pipe_busy_ns_per_zone"] = tot["pipeline"][2] * 1e6 / zones
    out["pipeline_time_ms_under_ncu"] = round(tot["pipeline"][1], 3)
    out["step_time_ms_under_ncu"] = round(tot["step"][1], 3)
    json.dump(out, sys.stdout, indent=1)
    print()


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]), sys.argv[3] if len(sys.argv) > 3 else "unknown")
