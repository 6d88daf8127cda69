"""Developer tool: one line per bench JSON file (ms/step and the per-kernel-family times)."""
import json
import sys
for f in sys.argv[1:]:
    try:
        d = None
        for ln in open(f):
            if ln.startswith("{"):
                d = json.loads(ln)
        k = d["roofline"]["kernels"]
        print("%-40s ms/step %.2f  pipeline %.2f | %s | clk %s %s replays/step %s" % (
            f.split("/")[-1], d["ms_per_step"], d["roofline"]["pipeline_ms_per_step"],
            " ".join("%s %.2f" % (n[2:], x["ms_per_step"]) for n, x in k.items()), d["clocks"]["sm_mhz"], d["clocks"]["reasons"],
            d.get("exact_replay_tiles_per_step")))
    except Exception as e:      # noqa: BLE001
        print(f, "FAILED", e)
