# GPU call 7 (1 GPU): exchange fused into the two-step kernel, single rank as its own neighbour; whole vector suite; N = 1 timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_step or fused_banded or peer_banded or cgrid" > gpurun_out/c7_tests.log 2>&1; echo "exit $?" >> gpurun_out/c7_tests.log
tail -4 gpurun_out/c7_tests.log
for v in "--fused" "--fused --push"; do
  echo "== banded $v"; timeout 300 python bench.py --workload cfg5 --banded $v --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('   G units/s', round(d['value']/1e9,1), 'ms/call', round(d['ms_per_step'],3), 'launches', d.get('gpu_launches'))"
done > gpurun_out/c7_banded_n1.log 2>&1
cat gpurun_out/c7_banded_n1.log
