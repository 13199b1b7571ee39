mkdir -p gpurun_out
timeout 420 python -X faulthandler -m pytest tests -m gpu -q -x -o faulthandler_timeout=120 > gpurun_out/final2_gpu_tests.log 2>&1; echo "pytest exit $?" >> gpurun_out/final2_gpu_tests.log
tail -4 gpurun_out/final2_gpu_tests.log
