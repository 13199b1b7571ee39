#!/bin/sh
# TEST-ONLY: build the host emulator of libgcmf (same sources, every launch is a host loop).
# Never loaded by the product package.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="$HERE/../../gcm_filters_b200/csrc"
g++ -O2 -std=c++17 -fPIC -shared -DGCMF_HOSTEMU -ffp-contract=off -x c++ "$SRC/gcmf.cu" -o "$HERE/libgcmf_hostemu.so"
