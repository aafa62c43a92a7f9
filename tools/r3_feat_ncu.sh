mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:feat_stream -s 3 -c 1 -o gpurun_out/r3_feat python tools/time_features.py 64 both > gpurun_out/r3_feat_ncu.log 2>&1
tail -2 gpurun_out/r3_feat_ncu.log
