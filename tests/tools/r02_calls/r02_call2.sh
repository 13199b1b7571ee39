# GPU call 2: the two-step C-grid kernel -- parity first, then timing (band-height sweep, one-step baseline)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "cgrid or peer_banded" > gpurun_out/c2_tests.log 2>&1; echo "exit $?" >> gpurun_out/c2_tests.log
tail -5 gpurun_out/c2_tests.log
timeout 600 python -m pytest tests/test_gpu_baseline_geometry.py -m gpu -x -q -k cfg5 >> gpurun_out/c2_tests.log 2>&1; echo "exit $?" >> gpurun_out/c2_tests.log
tail -3 gpurun_out/c2_tests.log
B="python bench.py --workload cfg5 --steps 10 --warmup 3 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
pick() { python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); r=d.get('roofline',{})
        print('$1', 'ms_per_call', round(d['ms_per_step'],4), 'Gpts/s', round(d['value']/1e9,2), 'mix', r.get('launch_mix_ms'), 'frac', r.get('frac'))
"; }
( $B --steps-per-block 1 | pick onestep
  $B | pick two_step_default
  for r in 24 36 48 59 72 90 120 180; do GCMF_CGRID_ROWS=$r $B | pick rows_$r; done ) > gpurun_out/c2_cfg5.log 2>&1
cat gpurun_out/c2_cfg5.log
