#!/usr/bin/env bash
# Developer A/B tool: build castro_b200/libcastro_ctu_<name>.so with extra -D flags (same sources, same arithmetic);
# select it at run time with CTU_LIB=<path> (castro_b200/lib.py).   usage: tools/build_variant.sh <name> [-DFLAG=..]...
# The translation units are the SRCS of castro_b200/csrc/Makefile; objects go to /tmp/ctu_variant_<name>.
set -euo pipefail
name=$1; shift
here="$(cd "$(dirname "$0")/.." && pwd)"
obj="/tmp/ctu_variant_$name"; mkdir -p "$obj"
make -C "$here/castro_b200/csrc" -j"$(nproc)" OBJDIR="$obj" LIB="$here/castro_b200/libcastro_ctu_$name.so" EXTRA="$*"
echo "built castro_b200/libcastro_ctu_$name.so ($*)"
