# Final 1-GPU validation of the round: whole GPU suite, the default bench line, the reference arm, ncu of the kernels changed last
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/final_smi.txt
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/final_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/final_gpu_tests.log
tail -3 gpurun_out/final_gpu_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1; tail -1 gpurun_out/final_smoke.log
timeout 900 python bench.py > gpurun_out/final_bench.json 2> gpurun_out/final_bench.err; tail -c 600 gpurun_out/final_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/final_bench_reference.json 2>> gpurun_out/final_bench.err
COMMON="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
cap() {  # tag kernel-regex skip "bench args"
    timeout 400 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/ncu_$1" python bench.py $4 $COMMON > "gpurun_out/ncu_$1.log" 2>&1
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page raw --csv > "gpurun_out/ncu_$1.raw.csv" 2>/dev/null
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page details > "gpurun_out/ncu_$1.details.txt" 2>/dev/null
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page source --csv > "gpurun_out/ncu_$1.source.csv" 2>/dev/null
    rm -f "gpurun_out/ncu_$1.ncu-rep"
}
cap march_final_cfg3 march_kernel 15 "--workload cfg3"
cap vec2_final_cgrid_cfg5 vec2_kernel 12 "--workload cfg5"
cap vec2_final_bgrid_cfgb vec2_kernel 12 "--workload cfgb"
cap vec2_halo_cgrid_cfg5 vec2_kernel 12 "--workload cfg5 --banded --fused --push"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_final_cfg3.csv python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy > /dev/null 2>&1
ls gpurun_out | head -50
