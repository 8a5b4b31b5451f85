#!/bin/bash
# Build libtn_b200.so for sm_100a (B200) in-tree.  Called by __graft_entry__.build().
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="$HERE/../lib"
OBJ="$HERE/../lib/obj"
mkdir -p "$OUT" "$OBJ"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
COMMON="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -I$HERE/../../include -Xcompiler -fPIC"
pids=()
build() {  # build <file> [extra flags]: skip when the object is newer than the source and headers
  local src="$HERE/$1"; shift
  local obj="$OBJ/$(basename "${src%.cu}").o"
  if [[ -f "$obj" && "$obj" -nt "$src" && "$obj" -nt "$HERE/tn_common.cuh" && "$obj" -nt "$HERE/tn_tc.cuh" && "$obj" -nt "$HERE/tn_encode_core.cuh" && "$obj" -nt "$HERE/tn_geometry.cuh" &&"$obj" -nt "$HERE/../../include/tn_b200.h" && "$obj" -nt "$HERE/build.sh" ]]; then return; fi
  $NVCC $COMMON "$@" -c "$src" -o "$obj" &
  pids+=($!)
}
# -fmad=false: these kernels must round every op like the reference's separate torch ops (bit-exact indices)
build tn_encode.cu -fmad=false
build tn_geometry.cu -fmad=false
build tn_ray.cu -fmad=false
build tn_prop.cu -fmad=false
build tn_level.cu -fmad=false
build tn_mlp.cu
build tn_mlp_tc.cu
build tn_fused.cu
build tn_model.cu
build tn_optim.cu
build tn_api.cu
build tn_probe.cu
for p in "${pids[@]:-}"; do [[ -n "$p" ]] && wait "$p"; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o "$OUT/libtn_b200.so" "$OBJ"/*.o -lcudart
echo "built $OUT/libtn_b200.so"
