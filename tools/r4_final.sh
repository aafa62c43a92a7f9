# session 5 (final build of round 2): bench line, feature-kernel A/B of the two mel schedules, one ncu capture of the segment-schedule feature kernel
mkdir -p gpurun_out
timeout 100 python bench.py --steps 20 --warmup 3 2>gpurun_out/r4_bench_err.log | tail -1 > gpurun_out/r4f_bench.json; tail -2 gpurun_out/r4_bench_err.log
for m in seg lane; do for a in "64 both" "64 mel" "32 both"; do TTSA_FEAT_MEL=$m timeout 60 python tools/time_features.py $a 2>&1 | tail -1 | sed "s/^/$m /"; done; done | tee gpurun_out/r4_feat_time.log
timeout 90 ncu --set full --clock-control none --import-source on -k regex:feat_stream -s 3 -c 1 -o gpurun_out/r4_feat python tools/time_features.py 64 both > gpurun_out/r4_feat_ncu.log 2>&1
tail -2 gpurun_out/r4_feat_ncu.log
