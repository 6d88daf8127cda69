#!/usr/bin/env bash
# Developer A/B tool (run under gpurun): a short bench per library variant (tools/build_variant.sh) with extra bench.py arguments.
# usage: tools/ab_lib.sh <tag> "<bench args>" base v1 v2 ...
tag=$1; bargs=$2; shift 2
mkdir -p gpurun_out
for v in "$@"; do
  if [ "$v" = base ]; then lib="$PWD/castro_b200/libcastro_ctu.so"; else lib="$PWD/castro_b200/libcastro_ctu_$v.so"; fi
  CTU_LIB=$lib python bench.py --steps 8 --warmup 3 --e2e-steps 0 --no-cpu --no-ref-gpu --no-checks $bargs > gpurun_out/${tag}_$v.json 2> gpurun_out/${tag}_$v.err
  python - "$tag $v" gpurun_out/${tag}_$v.json <<'PY'
import json, sys
try:
    d = json.loads([l for l in open(sys.argv[2]) if l.startswith("{")][-1])
    k = d["roofline"]["kernels"]
    print(sys.argv[1], "ms/step %.2f" % d["ms_per_step"], " ".join("%s %.1f" % (n[2:], x["ms_per_step"]) for n, x in k.items()), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
