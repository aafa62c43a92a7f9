mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 > gpurun_out/r3_tests_full.log; cat gpurun_out/r3_tests_full.log
for k in stream; do for w in both lin mel; do TTSA_FEAT_KERNEL=$k timeout 120 python tools/time_features.py 64 $w 2>&1 | tail -1 | sed "s/^/$k /"; done; done | tee gpurun_out/r3_feat_time.log
for k in stream tile; do TTSA_FEAT_KERNEL=$k timeout 120 python tools/time_features.py 32 both 2>&1 | tail -1 | sed "s/^/$k /"; done | tee -a gpurun_out/r3_feat_time.log
