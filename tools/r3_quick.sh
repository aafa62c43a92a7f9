mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "large_batch or stream_kernel or bench_step or full_size or ragged or griffin_lim_vs_oracle or host_pipeline_matches_oracle or randomised or edge or silent" 2>&1 | tail -15 > gpurun_out/r3_tests.log; cat gpurun_out/r3_tests.log
probe() { timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 2>gpurun_out/r3_err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'iter_ms', round(d['roofline']['launch_ms'],4), 'step_ms', round(d['ms_per_step'],3), 'value', round(d['value']), 'e2e', round(d['e2e']['value']), 'launches', d['gpu_launches'])"; }
for i in 1 2; do probe cur; done 2>&1 | tee gpurun_out/r3_ab.log
for v in $VARIANTS; do TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so probe var$v; TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so probe var$v; done 2>&1 | tee -a gpurun_out/r3_ab.log
tail -3 gpurun_out/r3_err.log
