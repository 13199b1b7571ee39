#!/bin/sh
# Development tool (CPU): build the host emulator of the CUDA sources with ASAN + UBSAN and run the randomised
# emulator-vs-oracle sweeps and the emulator test files under it.  Found (round 1) two out-of-bounds reads that a
# GPU silently tolerates: the C-grid tile wrapped indices only once (grids narrower than a tile, bands whose height
# is not a multiple of the tile), and load_mask of the fused REGULAR5 kernel looked one row beyond the ghost rows.
#   sh tests/tools/asan_fuzz.sh [cases]
set -e
ROOT="$(cd "$(dirname "$0")/../.." && pwd)"
N=${1:-500}
OUT=${TMPDIR:-/tmp}/libgcmf_hostemu_asan.so
g++ -O1 -g -std=c++17 -fPIC -shared -DGCMF_HOSTEMU -ffp-contract=off -fsanitize=address,undefined \
    -fno-omit-frame-pointer -x c++ "$ROOT/gcm_filters_b200/csrc/gcmf.cu" -o "$OUT"
export LD_PRELOAD="$(gcc -print-file-name=libasan.so) $(gcc -print-file-name=libubsan.so)"
export ASAN_OPTIONS=detect_leaks=0 GCMF_HOSTEMU_LIB="$OUT"
cd "$ROOT"
python tests/tools/fuzz_hostemu.py --cases "$N" --seed 3
python tests/tools/fuzz_pitch.py --cases "$N" --seed 8
python tests/tools/fuzz_bands.py --cases $((N / 10 + 10)) --seed 9
python -m pytest tests/test_hostemu.py tests/test_hostemu_fused.py tests/test_reference_suite.py \
    tests/test_scheduler_gloo.py -x -q -m "not gpu"
