# session 5: final feature kernel (segment schedule + warp-uniform scalars): live times, then one ncu capture
mkdir -p gpurun_out
for a in "32 both" "64 mel"; do timeout 30 python tools/time_features.py $a 2>&1 | tail -1; done | tee gpurun_out/r5_feat_time.log
timeout 45 ncu --set full --clock-control none --import-source on -k regex:feat_stream -s 3 -c 1 -o gpurun_out/r5_feat python tools/time_features.py 64 both > gpurun_out/r5_feat_ncu.log 2>&1
tail -1 gpurun_out/r5_feat_ncu.log
