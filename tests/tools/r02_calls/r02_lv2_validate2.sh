mkdir -p gpurun_out
timeout 36 python -m pytest tests/test_reference_suite.py tests/test_zz_gpu_c_driver.py -m gpu -x -q -p no:cacheprovider > gpurun_out/lv2_validate2.log 2>&1; echo "exit $?" >> gpurun_out/lv2_validate2.log
tail -3 gpurun_out/lv2_validate2.log
