#!/bin/sh
# Development tool for one gpurun call (1 GPU): the launch list of the default bench command and one `ncu --set full`
# capture per kernel family.  Outputs under gpurun_out/: launches_*.csv, ncu_<tag>.{raw.csv,details.txt} (the .ncu-rep
# files stay there too; copy the csv / txt summaries into profiles/).
#   gpurun --timeout 900 -- 'sh tests/tools/ncu_all.sh'
mkdir -p gpurun_out
COMMON="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg3.csv \
    python bench.py --steps 2 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy > /dev/null 2>&1
cap() {  # tag kernel-regex skip "bench args" [extra ncu args]
    ncu --set full --clock-control none $5 -k "regex:$2" -s "$3" -c 1 -f -o "gpurun_out/ncu_$1" python bench.py $4 $COMMON \
        > "gpurun_out/ncu_$1.log" 2>&1
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page raw --csv > "gpurun_out/ncu_$1.raw.csv" 2>/dev/null
    ncu -i "gpurun_out/ncu_$1.ncu-rep" --page details > "gpurun_out/ncu_$1.details.txt" 2>/dev/null
    grep -E "^  [a-z_]+.*Duration|DRAM Throughput|Registers Per" "gpurun_out/ncu_$1.details.txt" | head -3
    [ "$6" = keep ] || rm -f "gpurun_out/ncu_$1.ncu-rep"   # gpurun merges at most 64 MiB back
}
cap march_flux_cfg3 march_kernel 15 "--workload cfg3" "--import-source on" keep
ncu -i gpurun_out/ncu_march_flux_cfg3.ncu-rep --page source --csv > gpurun_out/ncu_march_flux_cfg3.source.csv 2>/dev/null
rm -f gpurun_out/ncu_march_flux_cfg3.ncu-rep
( export GCMF_FUSED_FORM=tile; cap fused_flux_tile_cfg3 fused_kernel 15 "--workload cfg3" )
cap fused_reg5_f32_cfg2 fused_kernel 4 "--workload cfg2"
cap fused_flux_tripolar_cfg4 fused_kernel 15 "--workload cfg4 --nb 8"
cap vec2_cgrid_cfg5 vec2_kernel 12 "--workload cfg5"
cap vec2_bgrid_cfgb vec2_kernel 12 "--workload cfgb"
cap vec2_halo_cgrid_cfg5 vec2_kernel 12 "--workload cfg5 --banded --fused --push"
cap cgrid_tma_cfg5 cgrid_tma 30 "--workload cfg5 --steps-per-block 1"
cap cgrid_tma_halo_cfg5 cgrid_tma 30 "--workload cfg5 --banded --peer"
cap halo_push_cfg5 halo_push 1 "--workload cfg5 --banded --peer"
cap step_vectorb step_kernel 30 "--workload cfgb --steps-per-block 1"
cap step_flux_onestep_cfg3 step_kernel 10 "--workload cfg3 --nb 8 --steps-per-block 1"
ls -la gpurun_out/ncu_*.raw.csv
