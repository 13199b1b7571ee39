# GPU call 4 (1 GPU): band-plan form of the two-step vector kernel (tests), ncu of vec2 (C-grid, B-grid), banded bench at N = 1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "two_step or fused_banded or peer_banded or cgrid" > gpurun_out/c4_tests.log 2>&1; echo "exit $?" >> gpurun_out/c4_tests.log
tail -4 gpurun_out/c4_tests.log
COMMON="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
cap() {  # tag kernel-regex skip "bench args"
    timeout 300 ncu --set full --clock-control none --import-source on -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/ncu_$1" python bench.py $4 $COMMON > "gpurun_out/ncu_$1.log" 2>&1
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page raw --csv > "gpurun_out/ncu_$1.raw.csv" 2>/dev/null
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page details > "gpurun_out/ncu_$1.details.txt" 2>/dev/null
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page source --csv > "gpurun_out/ncu_$1.source.csv" 2>/dev/null
    rm -f "gpurun_out/ncu_$1.ncu-rep"
}
cap vec2_cgrid_cfg5 vec2_kernel 12 "--workload cfg5"
cap vec2_bgrid_cfgb vec2_kernel 12 "--workload cfgb"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_cfg5.csv python bench.py --workload cfg5 --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy > /dev/null 2>&1
for v in "" "--peer" "--fused" ; do
  echo "== banded $v"; timeout 300 python bench.py --workload cfg5 --banded $v --steps 10 --warmup 3 2>&1 | tail -1 | cut -c1-400
done > gpurun_out/c4_banded_n1.log 2>&1
cat gpurun_out/c4_banded_n1.log
