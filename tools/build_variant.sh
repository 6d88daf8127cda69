#!/usr/bin/env bash
# Developer A/B tool: build castro_b200/libcastro_ctu_<name>.so with extra -D flags (same sources, same arithmetic);
# select it at run time with CTU_LIB=<path> (castro_b200/lib.py).   usage: tools/build_variant.sh <name> [-DFLAG=..]...
set -euo pipefail
name=$1; shift
here="$(cd "$(dirname "$0")/.." && pwd)"
src="$here/castro_b200/csrc"
obj="/tmp/ctu_variant_$name"; mkdir -p "$obj"
ARCH="-gencode arch=compute_100a,code=sm_100a"
for f in ctu_api ctu_sim ctu_dist ctu_io ctu_eos ctu_np3 ctu_np4 ctu_np5 ctu_rec ctu_rec3 ctu_rec4 ctu_rec5 ctu_ext ctu_ext3 ctu_ext4 ctu_ext5; do
  nvcc $ARCH -O3 -std=c++17 -lineinfo -fmad=false -Xcompiler -fPIC "$@" -c "$src/$f.cu" -o "$obj/$f.o" &
done
wait
nvcc $ARCH -shared -o "$here/castro_b200/libcastro_ctu_$name.so" "$obj/ctu_api.o" "$obj/ctu_sim.o" "$obj/ctu_dist.o" "$obj/ctu_io.o" "$obj/ctu_eos.o" "$obj/ctu_np3.o" "$obj/ctu_np4.o" "$obj/ctu_np5.o" "$obj/ctu_rec.o" "$obj/ctu_rec3.o" "$obj/ctu_rec4.o" "$obj/ctu_rec5.o" "$obj/ctu_ext.o" "$obj/ctu_ext3.o" "$obj/ctu_ext4.o" "$obj/ctu_ext5.o" -lcudart -ldl
echo "built castro_b200/libcastro_ctu_$name.so ($*)"
