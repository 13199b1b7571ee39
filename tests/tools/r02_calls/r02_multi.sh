# Multi-GPU call (N = $1 GPUs): band / shard tests at world N, then cfg5 strong scaling of the band decompositions
N=$1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_multi.py -m gpu -x -q ${MULTI_K:+-k "$MULTI_K"} > gpurun_out/multi_tests_${N}gpu.log 2>&1; echo "exit $?" >> gpurun_out/multi_tests_${N}gpu.log
tail -3 gpurun_out/multi_tests_${N}gpu.log
: > gpurun_out/multi_banded_${N}gpu.log
for v in "--peer" "--fused --peer" "--fused --push"; do
  echo "== cfg5 banded $v" >> gpurun_out/multi_banded_${N}gpu.log
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --workload cfg5 --banded $v --steps 10 --warmup 3 2>&1 | grep '^{' | tail -1 >> gpurun_out/multi_banded_${N}gpu.log
done
python - <<PY
import json
for l in open('gpurun_out/multi_banded_${N}gpu.log'):
    if l.startswith('=='): print(l.strip())
    elif l.startswith('{'):
        d=json.loads(l); print('   G units/s', round(d['value']/1e9,1), 'ms/call', round(d['ms_per_step'],3), 'launches', d.get('gpu_launches'))
PY
