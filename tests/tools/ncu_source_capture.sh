#!/bin/sh
# Development tool for one gpurun call: source-level ncu capture of one mid-recurrence fused launch of the headline
# workload, to settle where the stall cycles of fused_kernel<FLUX> sit (profiles/variants_r01.md: the
# long_scoreboard share is NOT the synchronous bar load; the mbarrier waits are the suspects).
#   gpurun --timeout 600 -- 'sh tests/tools/ncu_source_capture.sh [nb]'
# Outputs (gpurun_out/): prof_src.ncu-rep, prof_src.raw.csv, prof_src.source.csv, launches_src.csv
NB=${1:-16}     # levels: enough for a steady state, short enough for ~40 replays
mkdir -p gpurun_out
CMD="python bench.py --workload cfg3 --nb $NB --steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline"
# launch 0..10 = warm-up call (11 fused launches for 44 steps), take the 5th launch of the timed call (a mid block)
ncu --set full --clock-control none --import-source on -k regex:fused_kernel -s 15 -c 1 -f -o gpurun_out/prof_src $CMD \
    > gpurun_out/ncu_src.log 2>&1
ncu -i gpurun_out/prof_src.ncu-rep --page raw --csv > gpurun_out/prof_src.raw.csv 2>/dev/null
ncu -i gpurun_out/prof_src.ncu-rep --page source --csv > gpurun_out/prof_src.source.csv 2>/dev/null
# the per-instruction stall samples around the barrier waits and the bar load
grep -n "SYNCS\|LDG\|UBLKCP" gpurun_out/prof_src.source.csv | head -40
ls -la gpurun_out/prof_src.*
