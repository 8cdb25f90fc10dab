#!/usr/bin/env bash
# registers / spill stack / static shared memory of every kernel instantiation, from the built objects
# (cuobjdump --dump-resource-usage = what `ptxas -v` prints at build time) -> profiles/<round>_ptxas/
set -eu
out=${1:-profiles/r02_ptxas}
mkdir -p "$out"
for u in pt_kernels_parity pt_kernels_parity_sorted pt_kernels_parity_v4sw pt_kernels_fast pt_kernels_fast_sorted pt_kernels_fast_v4sw pt_post b200pt_group; do
  cuobjdump --dump-resource-usage cpuperformanceraytracer_b200/build/$u.o 2>/dev/null | grep -E "Function|REG" | paste - - \
    | sed -E 's/ Function /\n/; s/^ *//' | grep -v "^$" \
    | sed -E 's/^(.*):\s+REG:([0-9]+) STACK:([0-9]+) SHARED:([0-9]+) LOCAL:([0-9]+).*/\1 REG=\2 STACK=\3 SHARED=\4 LOCAL=\5/' > "$out/$u.resource_usage.txt"
done
wc -l "$out"/*.txt
