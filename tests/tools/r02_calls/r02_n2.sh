mkdir -p gpurun_out
timeout 330 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29544 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/final_bench_n2.json 2> gpurun_out/final_bench_n2.err; echo "exit $?"
python - <<PY
import json
for l in open('gpurun_out/final_bench_n2.json'):
    if l.startswith('{'):
        d=json.loads(l); print('value G', round(d['value']/1e9,1), 'ms', round(d['ms_per_step'],2), 'e2e', round(d['e2e']['value']/1e9,1) if d.get('e2e') else None)
        for s in d.get('secondary',[]): print('  SEC', s.get('workload','')[:50], round(s.get('value',0)/1e9,1) if 'value' in s else s.get('error'), s.get('sharding','')[:60])
PY
tail -3 gpurun_out/final_bench_n2.err | cut -c1-200
