# final-build evidence: launch list + full captures of every kernel of a step (bench must exit 0 without ncu first)
mkdir -p gpurun_out
CMD="python bench.py --no-cpu-baseline --no-extras --steps 1 --warmup 3 --no-graph"
$CMD > gpurun_out/r2f_plain.log 2>&1 || { tail -5 gpurun_out/r2f_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/r2f_launches_ncu.csv python bench.py --no-cpu-baseline --no-extras --steps 2 --warmup 3 --no-graph > gpurun_out/r2f_launches.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gl_stream -s 20 -c 3 -o gpurun_out/r2f_gl $CMD > gpurun_out/r2f_ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"mel_to_linear|frame_kernel|deemph" -c 6 -o gpurun_out/r2f_step $CMD > gpurun_out/r2f_ncu2.log 2>&1
tail -2 gpurun_out/r2f_ncu1.log gpurun_out/r2f_ncu2.log
python bench.py --steps 20 --warmup 3 2>/dev/null | tail -1 > gpurun_out/r2f_bench.json
python -c "
import json; d=json.load(open('gpurun_out/r2f_bench.json')); print({k: d[k] for k in ('value','ms_per_step','latency_single_ms')}, d['e2e']['value'], d['e2e_dropin']['value'], d['roofline']['frac'], d['roofline']['launch_ms'], d['roofline_fp32'], d['cpu_baseline']['value'])"
