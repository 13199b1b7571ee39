#!/bin/sh
# Development tool for one gpurun call: time the A/B variants under build/variants against the in-tree library,
# run the GPU test-suite on the in-tree build and, if a variant beats it by > 1.5 %, once more on that variant.
#   gpurun -- 'sh tests/tools/ab_and_verify.sh c00c offpf2 ...'
# VB_ARGS passes extra arguments to variant_bench.py, e.g. the cfg2-like REGULAR5 problem:
#   VB_ARGS="--op reg5 --dtype f32 --ny 720 --nx 1440 --nb 365 --steps 11" sh tests/tools/ab_and_verify.sh sm
V=build/variants
mkdir -p gpurun_out
ARGS="intree=gcm_filters_b200/libgcmf.so"
for n in "$@"; do ARGS="$ARGS $n=$V/libgcmf_$n.so"; done
# 3-second plain-C check of every variant against the host emulator first (bit-identical expected for all switches but
# CONTRACT): a variant that fails here is dropped from the timing
if [ ! -x tests/cabi/gpu_vs_emu ]; then
    gcc -std=c99 -O1 tests/cabi/gpu_vs_emu.c -I include -I /usr/local/cuda/include -L /usr/local/cuda/lib64 \
        -Wl,-rpath,/usr/local/cuda/lib64 -lcudart -ldl -lm -o tests/cabi/gpu_vs_emu
fi
[ -f tests/hostemu/libgcmf_hostemu.so ] || sh tests/hostemu/build.sh
if [ -x tests/cabi/gpu_vs_emu ]; then
    for n in "$@"; do
        ./tests/cabi/gpu_vs_emu "$V/libgcmf_$n.so" > "gpurun_out/gpu_vs_emu_$n.log" 2>&1 || echo "variant $n: differs from the emulator (see gpurun_out/gpu_vs_emu_$n.log)"
        tail -n 1 "gpurun_out/gpu_vs_emu_$n.log"
    done
fi
timeout 60 python tests/tools/variant_bench.py --reps 4 ${VB_ARGS:-} $ARGS > gpurun_out/variants_final.log 2>&1
cat gpurun_out/variants_final.log
timeout 70 python -m pytest tests -x -q -m gpu > gpurun_out/gpu_tests_intree.log 2>&1
echo "rc=$?" >> gpurun_out/gpu_tests_intree.log
tail -n 3 gpurun_out/gpu_tests_intree.log
best=$(python - <<'PY'
import json
rows = [json.loads(l) for l in open("gpurun_out/variants_final.log") if l.startswith("{")]
base = [r for r in rows if r["variant"] == "intree"][0]["ms"]
ok = [r for r in rows if r["maxdiff_vs_first"] < 1e-13 and r["nan_out"] == rows[0]["nan_out"]]
b = min(ok, key=lambda r: r["ms"])
print(b["variant"] if b["ms"] < 0.985 * base else "intree")
PY
)
echo "best=$best"
if [ "$best" != "intree" ] && [ -n "$best" ]; then
    cp "$V/libgcmf_$best.so" gcm_filters_b200/libgcmf.so
    timeout 45 python -m pytest tests/test_gpu_parity.py -x -q -m gpu > "gpurun_out/gpu_tests_$best.log" 2>&1
    echo "rc=$?" >> "gpurun_out/gpu_tests_$best.log"
    tail -n 3 "gpurun_out/gpu_tests_$best.log"
fi
