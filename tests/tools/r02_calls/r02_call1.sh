mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/c1_smi.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_gpu_tests.log
for form in tile march; do for nb in 62 8; do
  echo "== form=$form nb=$nb"; GCMF_FUSED_FORM=$form timeout 300 python tests/tools/variant_bench.py --nb $nb --reps 3 head=gcm_filters_b200/libgcmf.so
done; done > gpurun_out/c1_form_ab.log 2>&1
export GCMF_FUSED_FORM=tile
timeout 1500 sh tests/tools/ncu_all.sh > gpurun_out/c1_ncu_all.log 2>&1
export GCMF_FUSED_FORM=march
COMMON="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 15 -c 1 -f -o gpurun_out/ncu_march_cfg3 python bench.py --workload cfg3 $COMMON > gpurun_out/ncu_march_cfg3.log 2>&1
ncu -i gpurun_out/ncu_march_cfg3.ncu-rep --page raw --csv > gpurun_out/ncu_march_cfg3.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_march_cfg3.ncu-rep --page details > gpurun_out/ncu_march_cfg3.details.txt 2>/dev/null
ncu -i gpurun_out/ncu_march_cfg3.ncu-rep --page source --csv > gpurun_out/ncu_march_cfg3.source.csv 2>/dev/null
rm -f gpurun_out/ncu_march_cfg3.ncu-rep
tail -3 gpurun_out/c1_gpu_tests.log; cat gpurun_out/c1_form_ab.log | grep -v "^$" | tail -12
