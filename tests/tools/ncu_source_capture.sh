#!/bin/sh
# Development tool for one gpurun call: source-level ncu capture of one mid-recurrence launch of a bench workload.
#   gpurun --timeout 600 -- 'sh tests/tools/ncu_source_capture.sh TAG KERNEL_REGEX SKIP "bench args"'
# e.g.  sh tests/tools/ncu_source_capture.sh fused_cfg3 fused_kernel 15 "--workload cfg3 --nb 16"
# Outputs (gpurun_out/): ncu_TAG.ncu-rep, ncu_TAG.raw.csv, ncu_TAG.source.csv, ncu_TAG.details.txt
TAG=$1; KRE=$2; SKIP=$3; BARGS=$4
mkdir -p gpurun_out
CMD="python bench.py $BARGS --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary"
ncu --set full --clock-control none --import-source on -k "regex:$KRE" -s "$SKIP" -c 1 -f -o "gpurun_out/ncu_$TAG" $CMD \
    > "gpurun_out/ncu_$TAG.log" 2>&1
ncu -i "gpurun_out/ncu_$TAG.ncu-rep" --page raw --csv > "gpurun_out/ncu_$TAG.raw.csv" 2>/dev/null
ncu -i "gpurun_out/ncu_$TAG.ncu-rep" --page source --csv > "gpurun_out/ncu_$TAG.source.csv" 2>/dev/null
ncu -i "gpurun_out/ncu_$TAG.ncu-rep" --page details > "gpurun_out/ncu_$TAG.details.txt" 2>/dev/null
ls -la gpurun_out/ncu_$TAG.*
