# session 5: warp-uniform warp index / per-run scalars (TTSA_WARP_UNIFORM 2 = default build, 1 = _w, 0 = _nu) -- A/B on one box
mkdir -p gpurun_out
probe() { timeout 60 python bench.py --no-cpu-baseline --no-extras --steps 10 2>gpurun_out/r5_err.log | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'gl_iter_ms', round(d['roofline']['launch_ms'],4), 'step_ms', round(d['ms_per_step'],3))"; }
for v in "" _w _nu; do
  export TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so
  probe "lib$v"
  timeout 40 python tools/time_features.py 64 both 2>&1 | tail -1 | sed "s/^/lib$v /"
done 2>&1 | tee gpurun_out/r5_uniform_ab.log
