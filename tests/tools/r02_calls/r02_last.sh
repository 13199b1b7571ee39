mkdir -p gpurun_out
timeout 100 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined_host_path or concurrent_calls" > gpurun_out/last_tests.log 2>&1; echo "exit $?" >> gpurun_out/last_tests.log; tail -2 gpurun_out/last_tests.log
timeout 200 python bench.py --no-secondary > gpurun_out/last_bench.json 2> gpurun_out/last_bench.err
python - <<PY
import json
for l in open('gpurun_out/last_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value G', round(d['value']/1e9,1), 'ms', round(d['ms_per_step'],2), 'e2e G', round(d['e2e']['value']/1e9,1), 'e2e_numpy G', round(d['e2e_numpy']['value']/1e9,1), 'frac', round(d['roofline']['frac'],3), 'traffic', d['roofline']['traffic'], d['clocks'])
PY
