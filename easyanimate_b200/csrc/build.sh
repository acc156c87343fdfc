#!/usr/bin/env bash
# Builds libea_b200.so (sm_100a only) in-tree next to the sources' package: easyanimate_b200/libea_b200.so
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libea_b200.so"
OBJ="$HERE/../_build"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr
       -Xptxas -v -I"$HERE/../../include")
pids=()
for src in "$HERE"/*.cu; do
  obj="$OBJ/$(basename "${src%.cu}").o"
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$HERE/common.cuh" -nt "$obj" || "$HERE/host.h" -nt "$obj" || "$HERE/../../include/ea_b200.h" -nt "$obj" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" > "$obj.log" 2>&1 || { cat "$obj.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "$OBJ"/*.o -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
