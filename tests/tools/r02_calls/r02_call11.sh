# GPU call 11: march form with row-keyed register slots (8 iterations per trip): parity, then A/B against the tile form
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "march_form or fused_level_slabs or neighbour_sync" > gpurun_out/c11_tests.log 2>&1; echo "exit $?" >> gpurun_out/c11_tests.log
tail -3 gpurun_out/c11_tests.log
for form in tile march; do for nb in 62 8; do
  echo "== form=$form nb=$nb"; GCMF_FUSED_FORM=$form timeout 300 python tests/tools/variant_bench.py --nb $nb --reps 3 head=gcm_filters_b200/libgcmf.so
done; done > gpurun_out/c11_form_ab.log 2>&1
for r in 60 240; do echo "== march rows=$r nb=62"; GCMF_MARCH_ROWS=$r GCMF_FUSED_FORM=march timeout 300 python tests/tools/variant_bench.py --nb 62 --reps 3 head=gcm_filters_b200/libgcmf.so; done >> gpurun_out/c11_form_ab.log 2>&1
grep -v "^$" gpurun_out/c11_form_ab.log | cut -c1-200
