# final-build evidence of round 2 (session 3): launch list + full capture of the GL kernel + bench + stress of the hand-over
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 --no-graph"
$CMD > gpurun_out/r3f_plain.log 2>&1 || { tail -5 gpurun_out/r3f_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r3f_launches_ncu.csv python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 3 --no-graph > gpurun_out/r3f_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gl_stream -s 20 -c 3 -o gpurun_out/r3f_gl $CMD > gpurun_out/r3f_ncu1.log 2>&1
tail -2 gpurun_out/r3f_ncu1.log
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r3f_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r3f_bench.json')); print({k: d[k] for k in ('value','ms_per_step','latency_single_ms')}, d['e2e']['value'], d['e2e_dropin']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline_fp32'], d['cpu_baseline']['value'])"
python tools/stress_stream.py 25 > gpurun_out/stress_r3_stream.log 2>&1; tail -12 gpurun_out/stress_r3_stream.log
