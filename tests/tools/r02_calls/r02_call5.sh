# GPU call 5: vec2 kernel rewrite (two producer warps, division-free ring bookkeeping, unrolled consumer loop): parity + A/B
mkdir -p gpurun_out
cp gcm_filters_b200/libgcmf.so /tmp/libgcmf_intree.so
B="python bench.py --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d.get('roofline',{})
        print('$1', 'ms_per_call', round(d['ms_per_step'],4), 'Gpts/s', round(d['value']/1e9,2), 'mix', r.get('launch_mix_ms'))
"; }
: > gpurun_out/c5_ab.log
for n in intree "$@"; do
    if [ "$n" = intree ]; then cp /tmp/libgcmf_intree.so gcm_filters_b200/libgcmf.so; else cp "build/variants/libgcmf_$n.so" gcm_filters_b200/libgcmf.so; fi
    timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_step or fused_banded" > gpurun_out/c5_tests_$n.log 2>&1
    echo "$n tests: $(tail -1 gpurun_out/c5_tests_$n.log)" >> gpurun_out/c5_ab.log
    $B --workload cfg5 | pick ${n}_cfg5 >> gpurun_out/c5_ab.log 2>&1
    $B --workload cfgb | pick ${n}_cfgb >> gpurun_out/c5_ab.log 2>&1
done
cp /tmp/libgcmf_intree.so gcm_filters_b200/libgcmf.so
for r in 30 44; do GCMF_CGRID_ROWS=$r $B --workload cfg5 | pick intree_rows$r >> gpurun_out/c5_ab.log 2>&1; done
cat gpurun_out/c5_ab.log
