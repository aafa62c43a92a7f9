mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
python tools/time_gemm.py 2>&1 | tail -4 | tee gpurun_out/r2_time_gemm.log
