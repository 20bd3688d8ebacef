#!/usr/bin/env bash
# Build the reference's own C++ models into oracle/_ref/ (git-ignored, NOT gpurun-ignored).
#
# TEST INFRASTRUCTURE ONLY.  Compiles the sources WHERE THEY LIE under $REF (default
# /root/reference); nothing is copied into the repo.  The only edit is the width #defines
# (NPHASE/NWIDTH in hls/*/…​.h, PHASE_WIDTH/DATA_WIDTH in cpp/cordic_sincos.cpp), which the
# reference itself sets at compile time only; the patched header lives in a mktemp dir that is
# removed afterwards and the .cpp streams through stdin.  A driver (oracle/ref_driver_*.cc)
# is appended to the translation unit to expose a C entry point.
#
# Usage: oracle/build_ref.sh [REF_DIR]
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-${BHW_REFERENCE:-/root/reference}}"
OUT="$HERE/_ref"
if [ ! -d "$REF/hls/windows" ]; then
  echo "build_ref: reference not present at $REF - keeping prebuilt $OUT" >&2
  exit 0
fi
mkdir -p "$OUT"
CXX="${CXX:-g++}"
CXXFLAGS="-O2 -fPIC -shared -std=c++14 -w"
TMP="$(mktemp -d)"
trap 'rm -rf "$TMP"' EXIT

# (NPHASE x NWIDTH) pairs: the BASELINE configs + width/phase edge cases (NP-1<NW and NP>=NW+1)
HLS_CFGS="${BHW_HLS_CFGS:-4x8 6x6 8x32 10x16 10x24 11x16 12x12 14x12 16x16 16x17 17x16 18x16 20x18 20x32 22x24}"
# (PHASE_WIDTH x DATA_WIDTH) pairs for cpp/cordic_sincos.cpp (14x12 is the file's own default)
CPP_CFGS="${BHW_CPP_CFGS:-14x12 10x16 16x24 20x32 12x12 13x12 8x20}"

build_one() { # $1=label
  if ! "$@" ; then echo "build_ref: FAILED: $*" >&2; return 1; fi
}

for cfg in $HLS_CFGS; do
  np="${cfg%x*}"; nw="${cfg#*x}"
  d="$TMP/hlsw_$cfg"; mkdir -p "$d"
  sed -e "s/^#define NPHASE .*/#define NPHASE $np/" -e "s/^#define NWIDTH .*/#define NWIDTH $nw/" \
      "$REF/hls/windows/win_function.h" > "$d/win_function.h"
  cat "$REF/hls/windows/win_function.cpp" "$HERE/ref_driver_hls_win.cc" | \
      $CXX $CXXFLAGS -x c++ -I"$d" -I"$HERE/shim" - -o "$OUT/hls_win_np${np}_nw${nw}.so"
  d="$TMP/hlsc_$cfg"; mkdir -p "$d"
  sed -e "s/^#define NPHASE .*/#define NPHASE $np/" -e "s/^#define NWIDTH .*/#define NWIDTH $nw/" \
      "$REF/hls/cordic/cordic.h" > "$d/cordic.h"
  cat "$REF/hls/cordic/cordic.cpp" "$HERE/ref_driver_hls_cordic.cc" | \
      $CXX $CXXFLAGS -x c++ -I"$d" -I"$HERE/shim" - -o "$OUT/hls_cordic_np${np}_nw${nw}.so"
done

for cfg in $CPP_CFGS; do
  pw="${cfg%x*}"; dw="${cfg#*x}"
  sed -e "s/^#define PHASE_WIDTH .*/#define PHASE_WIDTH $pw/" -e "s/^#define DATA_WIDTH .*/#define DATA_WIDTH $dw/" \
      "$REF/cpp/cordic_sincos.cpp" | cat - "$HERE/ref_driver_cpp.cc" | \
      $CXX $CXXFLAGS -x c++ -include "$HERE/shim/msvc_shim.h" -Dmain=ref_cpp_main_impl - \
      -o "$OUT/cpp_cordic_pw${pw}_dw${dw}.so"
done
ls "$OUT" | wc -l | xargs echo "build_ref: built objects:"
