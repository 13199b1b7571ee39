mkdir -p gpurun_out
for c in 5 7 13; do
  GCMF_PIPELINE_CHUNKS=$c timeout 120 python bench.py --no-secondary --no-cpu-baseline --no-e2e-numpy --steps 3 --warmup 3 2>/dev/null | python -c "
import json,sys
for l in sys.stdin:
    if l.startswith('{'):
        d=json.loads(l); print('target_chunks $c', 'e2e G', round(d['e2e']['value']/1e9,1), 'value G', round(d['value']/1e9,1))"
done > gpurun_out/last2_chunks.log 2>&1
cat gpurun_out/last2_chunks.log
