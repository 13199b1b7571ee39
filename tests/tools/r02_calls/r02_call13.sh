# GPU call 13: march form with the check-free steady-state iteration: parity, timing
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_baseline_geometry.py -m gpu -x -q -k "march_form or fused_level_slabs or neighbour_sync or cfg3 or smoke or captured_reference" > gpurun_out/c13_tests.log 2>&1; echo "exit $?" >> gpurun_out/c13_tests.log
tail -3 gpurun_out/c13_tests.log
VB="python tests/tools/variant_bench.py --reps 3 head=gcm_filters_b200/libgcmf.so"
( for nb in 62 8 1; do echo "== march default nb=$nb"; timeout 300 $VB --nb $nb; done
  for r in 200 800; do echo "== march rows=$r nb=62"; GCMF_MARCH_ROWS=$r timeout 300 $VB --nb 62; done ) > gpurun_out/c13_ab.log 2>&1
grep -v "^$" gpurun_out/c13_ab.log | cut -c1-120
