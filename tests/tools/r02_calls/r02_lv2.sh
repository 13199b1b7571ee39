# Last GPU seconds of round 2: march kernel with two levels per CTA and two CTAs per SM (16 consumer warps per SM, 96 registers) vs the in-tree form
mkdir -p gpurun_out
export GCMF_FUSED_FORM=march
( timeout 45 python tests/tools/variant_bench.py --nb 62 --reps 2 head=gcm_filters_b200/libgcmf.so lv2=build/variants/libgcmf_lv2.so
  timeout 30 python tests/tools/variant_bench.py --nb 8 --reps 2 head=gcm_filters_b200/libgcmf.so lv2=build/variants/libgcmf_lv2.so ) > gpurun_out/lv2_ab.log 2>&1
grep -v "^$" gpurun_out/lv2_ab.log | cut -c1-160
