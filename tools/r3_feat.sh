mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "feature or spectrogram_and_mel or collate or randomised or tensor_in or full_size" 2>&1 | tail -5 > gpurun_out/r3_feat_tests.log; cat gpurun_out/r3_feat_tests.log
for k in stream; do for w in both lin mel; do TTSA_FEAT_KERNEL=$k timeout 120 python tools/time_features.py 64 $w 2>&1 | tail -1 | sed "s/^/$k /"; done; done | tee gpurun_out/r3_feat_time.log
for k in stream tile; do TTSA_FEAT_KERNEL=$k timeout 120 python tools/time_features.py 32 both 2>&1 | tail -1 | sed "s/^/$k /"; done | tee -a gpurun_out/r3_feat_time.log
