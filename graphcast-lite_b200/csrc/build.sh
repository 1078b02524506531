#!/usr/bin/env bash
# Builds libgcl_b200.so (sm_100a only) next to the Python host package.  Used by __graft_entry__.build().
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
out="$here/../gcl_b200/libgcl_b200.so"
obj="$here/_obj"
mkdir -p "$obj"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -lineinfo -Xcompiler -fPIC ${GCL_NVCC_EXTRA:-})
pids=()
for f in "$here"/*.cu; do
  o="$obj/$(basename "${f%.cu}").o"
  if [[ ! -f "$o" || "$f" -nt "$o" || "$here/common.cuh" -nt "$o" || "$here/scan.cuh" -nt "$o" || "$here/tile.cuh" -nt "$o" || "$here/ws.cuh" -nt "$o" || "$here/../../include/gcl_b200.h" -nt "$o" ]]; then
    "$NVCC" "${FLAGS[@]}" -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
"$NVCC" -shared -cudart static -gencode arch=compute_100a,code=sm_100a -o "$out" "$obj"/*.o
echo "built $out"
