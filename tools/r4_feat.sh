# session 4: segment schedule of the mel basis in the warp-stream feature kernel -- parity tests, then A/B against the lane schedule
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "feature or spectrogram_and_mel or collate or randomised or training_batch or mel_basis_that or tensor_in" 2>&1 | tail -8 > gpurun_out/r4_feat_tests.log; cat gpurun_out/r4_feat_tests.log
for m in seg lane; do for w in both mel; do TTSA_FEAT_MEL=$m timeout 120 python tools/time_features.py 64 $w 2>&1 | tail -1 | sed "s/^/$m /"; done; done | tee gpurun_out/r4_feat_time.log
for m in seg lane; do TTSA_FEAT_MEL=$m timeout 120 python tools/time_features.py 32 both 2>&1 | tail -1 | sed "s/^/$m /"; done | tee -a gpurun_out/r4_feat_time.log
