#!/usr/bin/env bash
# Developer A/B tool (run under gpurun): the gpu tests once, then a short 512^3 bench per value of one environment switch.
# usage: tools/ab_env.sh <tag> <ENVVAR> v1 v2 ...      e.g. tools/ab_env.sh r3a CTU_ZGROUP 1 2 4 8
tag=$1; var=$2; shift 2
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${tag}_tests.log 2>&1; tail -1 gpurun_out/${tag}_tests.log
for v in "$@"; do
  env $var=$v python bench.py --steps 8 --warmup 3 --e2e-steps 0 --no-cpu --no-ref-gpu --no-checks > gpurun_out/${tag}_${var}_$v.json 2> gpurun_out/${tag}_${var}_$v.err
  python - "$var=$v" gpurun_out/${tag}_${var}_$v.json <<'PY'
import json, sys
try:
    d = json.load(open(sys.argv[2]))
    k = d["roofline"]["kernels"]
    print(sys.argv[1], "ms/step %.2f" % d["ms_per_step"], " ".join("%s %.1f" % (n[2:], x["ms_per_step"]) for n, x in k.items()), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
