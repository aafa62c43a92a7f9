mkdir -p gpurun_out
timeout 120 tools/bin/ubench_tmem > gpurun_out/r2_ubench_tmem.log 2>&1
CMD="python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gl_stream -s 20 -c 2 -o gpurun_out/r2a_wps $CMD > gpurun_out/ncu_r2a.log 2>&1
tail -3 gpurun_out/ncu_r2a.log
