# GPU call 14: march form on short batches (one GPU's share of an 8-GPU split): band-height sweep
mkdir -p gpurun_out
VB="python tests/tools/variant_bench.py --reps 3 head=gcm_filters_b200/libgcmf.so"
( for r in 50 75 100 150 200 300; do echo "== march rows=$r nb=8"; GCMF_MARCH_ROWS=$r timeout 300 $VB --nb 8; done
  for nb in 7 9 16 31; do echo "== march default nb=$nb"; timeout 300 $VB --nb $nb; done
  for nb in 7 9 16 31; do echo "== tile nb=$nb"; GCMF_FUSED_FORM=tile timeout 300 $VB --nb $nb; done ) > gpurun_out/c14_ab.log 2>&1
grep -v "^$" gpurun_out/c14_ab.log | cut -c1-110
