#!/bin/sh
# Development tool for one gpurun call: time A/B variants of libgcmf.so (build/variants, tests/tools/build_variant.py)
# on one bench.py workload.  The in-tree library is timed first, then each variant is copied over it (on the GPU box's
# scratch copy of the repo only).
#   gpurun -- 'sh tests/tools/ab_bench_workload.sh cfg5 cg16 cgmb4 ...'
W=$1
shift
mkdir -p gpurun_out
cp gcm_filters_b200/libgcmf.so /tmp/libgcmf_intree.so
OUT=gpurun_out/ab_$W.jsonl
: > "$OUT"
for n in intree "$@"; do
    if [ "$n" = intree ]; then cp /tmp/libgcmf_intree.so gcm_filters_b200/libgcmf.so; else cp "build/variants/libgcmf_$n.so" gcm_filters_b200/libgcmf.so; fi
    python bench.py --workload "$W" --no-cpu-baseline --steps 20 --warmup 3 --e2e-steps 1 ${BENCH_ARGS:-} 2> "gpurun_out/ab_${W}_$n.err" |
        python -c "
import json, sys
for l in sys.stdin:
    if l.startswith('{'):
        d = json.loads(l)
        print(json.dumps({'variant': '$n', 'workload': '$W', 'value_G': round(d['value'] / 1e9, 2), 'ms_per_step': round(d['ms_per_step'], 3),
                          'frac': round(d['roofline']['frac'], 3), 'launch_mix_ms': d['roofline'].get('launch_mix_ms')}))
" | tee -a "$OUT"
done
cp /tmp/libgcmf_intree.so gcm_filters_b200/libgcmf.so
