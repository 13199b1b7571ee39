mkdir -p gpurun_out
timeout 150 python -X faulthandler -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "pipelined_host_path" -o faulthandler_timeout=60 > gpurun_out/dbg_pipelined.log 2>&1; echo "exit $?" >> gpurun_out/dbg_pipelined.log
tail -40 gpurun_out/dbg_pipelined.log | cut -c1-200
