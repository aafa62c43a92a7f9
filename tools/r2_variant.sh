# usage: VARIANTS="_x _y" bash tools/r2_variant.sh  -- parity subset + A/B timing of experiment builds
mkdir -p gpurun_out
for v in $VARIANTS; do
  TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "large_batch or bench_step or ragged or randomised" 2>&1 | tail -2
done
bash tools/r2_ab.sh
