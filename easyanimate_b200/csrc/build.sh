#!/usr/bin/env bash
# Builds libea_b200.so (sm_100a only) in-tree next to the sources' package: easyanimate_b200/libea_b200.so
#   EA_ATTN_AB=1  also links the retired attention generations of tools/attn_ab/ (A/B measurements only)
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../libea_b200.so"
OBJ="$HERE/../_build"
mkdir -p "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr
       -Xptxas -v -I"$HERE/../../include")
SRCS=("$HERE"/*.cu)
AB="${EA_ATTN_AB:-0}"
if [[ "$AB" == "1" ]]; then
  FLAGS+=(-DEA_ATTN_AB)
  SRCS+=("$HERE"/../../tools/attn_ab/*.cu)
fi
# a change of the A/B setting invalidates the dispatcher object
if [[ ! -f "$OBJ/.ab" || "$(cat "$OBJ/.ab")" != "$AB" ]]; then rm -f "$OBJ/attn_api.o"; echo "$AB" > "$OBJ/.ab"; fi
pids=()
objs=()
for src in "${SRCS[@]}"; do
  obj="$OBJ/$(basename "${src%.cu}").o"
  objs+=("$obj")
  if [[ ! -f "$obj" || "$src" -nt "$obj" || "$HERE/common.cuh" -nt "$obj" || "$HERE/host.h" -nt "$obj" || "$HERE/../../include/ea_b200.h" -nt "$obj" ]]; then
    ( "$NVCC" "${FLAGS[@]}" -c "$src" -o "$obj" > "$obj.log" 2>&1 || { cat "$obj.log"; exit 1; } ) &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$OUT" "${objs[@]}" -lcudart_static -ldl -lrt -lpthread
echo "built $OUT"
