#!/usr/bin/env bash
# Developer A/B tool (run under gpurun): parity tests + a short 512^3 bench for each library variant given on the command line.
# usage: tools/ab_bench.sh base v1 v2 ...   (base = the product library)
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then lib="$PWD/castro_b200/libcastro_ctu.so"; else lib="$PWD/castro_b200/libcastro_ctu_$v.so"; fi
  echo "=== $v"
  CTU_LIB=$lib python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
  CTU_LIB=$lib python bench.py --steps 5 --warmup 3 --e2e-steps 0 --no-cpu > gpurun_out/ab_$v.json 2> gpurun_out/ab_$v.err
  python - "$v" <<'PY'
import json, sys
v = sys.argv[1]
try:
    d = json.load(open("gpurun_out/ab_%s.json" % v))
    k = d["roofline"]["kernels"]
    print(v, "ms/step %.2f" % d["ms_per_step"], " ".join("%s %.1f" % (n[2:], x["ms_per_step"]) for n, x in k.items()), d["clocks"]["sm_mhz"], d["clocks"]["reasons"], "replays/step", d.get("exact_replay_tiles_per_step"))
except Exception as e:
    print(v, "FAILED", e)
PY
done
