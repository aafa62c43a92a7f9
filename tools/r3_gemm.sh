mkdir -p gpurun_out
timeout 300 python -m pytest tests -m gpu -x -q -k "mel_gemm or mel_linear or bench_step or tacotron2 or host_pipeline_matches_oracle or inv_mel" 2>&1 | tail -6 > gpurun_out/r3_gemm_tests.log; cat gpurun_out/r3_gemm_tests.log
timeout 120 python tools/time_gemm.py 2>&1 | head -1 | sed 's/^/tc2  /' | tee gpurun_out/r3_gemm_time.log
TTSA_MEL_GEMM=tc96 timeout 120 python tools/time_gemm.py 2>&1 | head -1 | sed 's/^/tc96 /' | tee -a gpurun_out/r3_gemm_time.log
