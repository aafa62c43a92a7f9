mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:mel_to_linear_tc2 -s 10 -c 1 -o gpurun_out/r3_gemm python tools/time_gemm.py > gpurun_out/r3_gemm_ncu.log 2>&1
tail -2 gpurun_out/r3_gemm_ncu.log
