# GPU call 12: march form -- band-height sweep, fp32, ncu capture
mkdir -p gpurun_out
VB="python tests/tools/variant_bench.py --reps 3 head=gcm_filters_b200/libgcmf.so"
export GCMF_FUSED_FORM=march
( for r in 240 400 600 800 1200 2400; do echo "== march rows=$r nb=62"; GCMF_MARCH_ROWS=$r timeout 300 $VB --nb 62; done
  for r in 120 240 600 1200; do echo "== march rows=$r nb=8"; GCMF_MARCH_ROWS=$r timeout 300 $VB --nb 8; done
  echo "== march f32 nb=62 rows=240"; GCMF_MARCH_ROWS=240 timeout 300 $VB --nb 62 --dtype f32
  echo "== tile f32 nb=62"; GCMF_FUSED_FORM=tile timeout 300 $VB --nb 62 --dtype f32 ) > gpurun_out/c12_sweep.log 2>&1
grep -v "^$" gpurun_out/c12_sweep.log | cut -c1-120
export GCMF_MARCH_ROWS=240
COMMON="--steps 1 --warmup 1 --e2e-steps 1 --no-cpu-baseline --no-secondary --no-e2e-numpy"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:march_kernel -s 15 -c 1 -f -o gpurun_out/ncu_march8_cfg3 python bench.py --workload cfg3 $COMMON > gpurun_out/ncu_march8_cfg3.log 2>&1
ncu -i gpurun_out/ncu_march8_cfg3.ncu-rep --page raw --csv > gpurun_out/ncu_march8_cfg3.raw.csv 2>/dev/null
ncu -i gpurun_out/ncu_march8_cfg3.ncu-rep --page details > gpurun_out/ncu_march8_cfg3.details.txt 2>/dev/null
ncu -i gpurun_out/ncu_march8_cfg3.ncu-rep --page source --csv > gpurun_out/ncu_march8_cfg3.source.csv 2>/dev/null
rm -f gpurun_out/ncu_march8_cfg3.ncu-rep
