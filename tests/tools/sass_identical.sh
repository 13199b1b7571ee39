#!/bin/sh
# Development tool: is the device code of the in-tree libgcmf.so identical to the one built from commit $1?
# Used when opt-in variants (GCMF_OPT_* switches, default off) are added without a GPU at hand: if the SASS of the
# default build does not change, the last GPU verification still stands.
#   sh tests/tools/sass_identical.sh <commit>
set -e
REF=${1:?commit}
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
TMP=$(mktemp -d)
git -C "$ROOT" archive "$REF" gcm_filters_b200/csrc include | tar x -C "$TMP"
(cd "$TMP" && nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -fmad=false -shared \
    -Xcompiler -fPIC -diag-suppress 128 gcm_filters_b200/csrc/gcmf.cu -o ref.so)
python -m gcm_filters_b200.build > /dev/null
filter() { cuobjdump -sass "$1" | grep -v "Fatbin\|===\|host =\|compile_size\|identifier"; }
filter "$TMP/ref.so" > "$TMP/a.sass"
filter "$ROOT/gcm_filters_b200/libgcmf.so" > "$TMP/b.sass"
if cmp -s "$TMP/a.sass" "$TMP/b.sass"; then echo "SASS identical to $REF"; rm -rf "$TMP"; else echo "SASS DIFFERS from $REF (see $TMP)"; exit 1; fi
