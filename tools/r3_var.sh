mkdir -p gpurun_out
probe() { timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 10 2>gpurun_out/r3_err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'iter_ms', round(d['roofline']['launch_ms'],4), 'step_ms', round(d['ms_per_step'],3))"; }
for rep in 1 2; do for v in $VARIANTS; do TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so probe var$v; done; done 2>&1 | tee gpurun_out/r3_var.log
