mkdir -p gpurun_out
timeout 120 tools/bin/ubench_dft_tc > gpurun_out/r2_ubench_dft_tc.log 2>&1; cat gpurun_out/r2_ubench_dft_tc.log
CMD="python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gl_stream -s 20 -c 1 -o gpurun_out/r2b_wps $CMD > gpurun_out/ncu_r2b.log 2>&1
tail -3 gpurun_out/ncu_r2b.log
