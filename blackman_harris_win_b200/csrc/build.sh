#!/usr/bin/env bash
# Build blackman_harris_win_b200/libbhw.so for sm_100a (cross-compiles without a GPU).
# The translation units are compiled in parallel, then linked.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libbhw.so"
OBJ="$HERE/../build"
if [ "${1:-}" = "debug" ]; then   # bounds-checked build (device-side asserts), loaded with BHW_LIB=.../libbhw_debug.so
  OUT="$HERE/../libbhw_debug.so"
  OBJ="$HERE/../build/debug"
  BHW_NVCC_EXTRA="${BHW_NVCC_EXTRA:-} -DBHW_BOUNDS_CHECK"
fi
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden"
mkdir -p "$OBJ"
pids=()
for f in bhw_kernels.cu bhw_group.cu bhw_api.cu bhw_resolve.cpp bhw_plan.cpp; do
  "$NVCC" $FLAGS ${BHW_NVCC_EXTRA:-} -c "$HERE/$f" -o "$OBJ/${f%.*}.o" &
  pids+=($!)
done
for p in "${pids[@]}"; do wait "$p"; done
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared "$OBJ"/bhw_kernels.o "$OBJ"/bhw_group.o "$OBJ"/bhw_api.o "$OBJ"/bhw_resolve.o "$OBJ"/bhw_plan.o -o "$OUT"
echo "built $OUT"
