#!/usr/bin/env bash
# Instruction stream of one kernel instantiation, normalised (no addresses, no parameter offsets), for
# diffing across commits: the bank kernel is sensitive to instruction scheduling, so a change that is
# not meant to touch the bench kernel should leave this output identical.
#   tools/sass_check.sh [mangled-name-substring] > /tmp/a.txt ; ... ; diff /tmp/a.txt /tmp/b.txt
set -euo pipefail
HERE="$(cd "$(dirname "$0")/.." && pwd)"
PAT="${1:-k_synth_bankILi4ELi1ELi1ELb0EE}"   # bench kernel: 4 terms, staged half table, paired, 32-bit tail
TMP="$(mktemp -d)"
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -cubin \
  -I"$HERE/include" -o "$TMP/k.cubin" "$HERE/blackman_harris_win_b200/csrc/bhw_kernels.cu"
cuobjdump -sass "$TMP/k.cubin" | awk -v pat="$PAT" '/Function : /{f = index($0, pat) > 0; next} f' \
  | grep -E "^\s+/\*[0-9a-f]{4}\*/" | sed -E 's/^\s+\/\*[0-9a-f]+\*\/\s+//; s/\s*\/\*.*$//; s/c\[0x0\]\[0x[0-9a-f]+\]/c[P]/g'
rm -rf "$TMP"
