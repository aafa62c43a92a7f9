mkdir -p gpurun_out
for k in stream tile; do for w in both lin mel; do TTSA_FEAT_KERNEL=$k timeout 120 python tools/time_features.py 64 $w 2>&1 | tail -1 | sed "s/^/$k /"; done; done | tee gpurun_out/r3_feat_time.log
