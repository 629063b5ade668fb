#!/bin/bash
# Builds liblob_b200.so for sm_100a (B200) in-tree.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
ROOT="$(cd "$HERE/../.." && pwd)"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
OUT="$HERE/liblob_b200.so"
OBJ="$HERE/build"
mkdir -p "$OBJ"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -I"$ROOT/include" -I"$HERE")
pids=()
objs=()
for src in "$HERE"/*.cu; do
  o="$OBJ/$(basename "${src%.cu}").o"
  objs+=("$o")
  if [[ ! -f "$o" || "$src" -nt "$o" || "$HERE/common.cuh" -nt "$o" || "$HERE/simt_tile.cuh" -nt "$o" || "$HERE/tcgen05_util.cuh" -nt "$o" || "$ROOT/include/lob_b200.h" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" ${LOB_PTXAS_V:+-Xptxas -v} -c "$src" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -o "$OUT" "${objs[@]}" -L/usr/local/cuda/lib64 -lcufft -Xlinker -rpath -Xlinker /usr/local/cuda/lib64
echo "built $OUT"
