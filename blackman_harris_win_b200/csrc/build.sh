#!/usr/bin/env bash
# Build blackman_harris_win_b200/libbhw.so for sm_100a (cross-compiles without a GPU).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../libbhw.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC,-fvisibility=hidden -shared"
"$NVCC" $FLAGS ${BHW_NVCC_EXTRA:-} "$HERE/bhw_kernels.cu" "$HERE/bhw_api.cu" "$HERE/bhw_resolve.cpp" "$HERE/bhw_plan.cpp" -o "$OUT"
echo "built $OUT"
