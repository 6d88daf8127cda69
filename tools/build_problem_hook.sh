#!/usr/bin/env bash
# Builds a source-compatible problem hook library from the reference's UNMODIFIED problem directory:
#   tools/build_problem_hook.sh <problem dir> <name> [NUMSPEC [NUMADV [NUMAUX [DIM]]]]   ->  castro_b200/hooks/libctu_hook_<name>.so
# e.g. tools/build_problem_hook.sh /root/reference/Exec/hydro_tests/Sedov sedov
# The library is selected at run time with ctu_problem_init("hook:<path>", ...) / Sim.init_problem("hook:<path>", ...).
set -euo pipefail
dir=$1; name=$2; nspec=${3:-1}; nadv=${4:-0}; naux=${5:-0}; dim=${6:-3}
here="$(cd "$(dirname "$0")/.." && pwd)"
hooks="$here/castro_b200/hooks"
gen="/tmp/ctu_hook_gen_$name"; mkdir -p "$gen"
common="$(cd "$dir/../../.." && pwd)/Source/problems/_prob_params"
tables=(); [ -f "$dir/_prob_params" ] && tables+=("$dir/_prob_params")      # a problem without parameters of its own (Noh) has none
[ -f "$common" ] && tables+=("$common")
python "$hooks/gen_prob_params.py" "$gen" "${tables[@]}"
# -ffp-contract=off, no fast-math: the reference's CPU build contains no FMA (generic x86-64), and the initial data must match it bit for bit
g++ -O2 -std=c++17 -fPIC -shared -fopenmp -fno-gnu-unique -fvisibility=hidden -ffp-contract=off -DAMREX_SPACEDIM=$dim -DCTU_HOOK_NUMSPEC=$nspec -DCTU_HOOK_NUMADV=$nadv -DCTU_HOOK_NUMAUX=$naux \
    -I"$hooks/shim" -I"$gen" -I"$dir" -I"$here/include" "$hooks/hook_main.cpp" -o "$hooks/libctu_hook_$name.so"
echo "built castro_b200/hooks/libctu_hook_$name.so from $dir (NumSpec $nspec NumAdv $nadv NumAux $naux)"
