mkdir -p gpurun_out
timeout 62 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "march_form or pipelined_host_path or fused_level_slabs or neighbour_sync or fused_steps_equal" -p no:cacheprovider > gpurun_out/lv2_validate.log 2>&1; echo "exit $?" >> gpurun_out/lv2_validate.log
tail -3 gpurun_out/lv2_validate.log
