mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 --no-graph"
$CMD > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:gl_stream -s 20 -c 1 -o gpurun_out/r3a_gl $CMD > gpurun_out/ncu_r3a.log 2>&1
tail -2 gpurun_out/ncu_r3a.log
