#!/bin/sh
# Build every prepared A/B variant of libgcmf.so (profiles/variants_r01.md) in parallel; ~2 min on 8 cores.
#   sh tests/tools/build_all_variants.sh && gpurun --timeout 300 -- 'sh tests/tools/ab_and_verify.sh ss wo edge sswoedge rp rownan'
cd "$(dirname "$0")/../.."
B="python tests/tools/build_variant.py"
$B ss -DGCMF_OPT_SANSTATE=1 &
$B wo -DGCMF_OPT_WRAPONCE=1 &
$B edge -DGCMF_OPT_EDGEREFILL=1 &
$B rp -DGCMF_OPT_ROWPTR=1 &
wait
$B rownan -DGCMF_OPT_ROWNAN=1 &
$B sswoedge -DGCMF_OPT_SANSTATE=1 -DGCMF_OPT_WRAPONCE=1 -DGCMF_OPT_EDGEREFILL=1 -DGCMF_OPT_ROWPTR=1 &
$B sswoedgeskip -DGCMF_OPT_SANSTATE=1 -DGCMF_OPT_WRAPONCE=1 -DGCMF_OPT_EDGEREFILL=1 -DGCMF_OPT_ROWPTR=1 -DGCMF_OPT_SKIPLAST=1 &
$B sm -DGCMF_OPT_STATICMASK=1 &
wait
ls -la build/variants/*.so
for f in build/variants/*.ptxas.log; do
    printf '%s: ' "$(basename "$f" .so.ptxas.log)"
    grep -A3 "fused_kernelIdLi0ELi0" "$f" | grep -i "spill\|Used" | sed 's/ptxas info    : //' | tr '\n' ' '
    echo
done
