mkdir -p gpurun_out
for rep in 1 2; do
for v in "" _p8192 _p24576 _p57344; do
  TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so timeout 300 python bench.py --no-cpu-baseline --no-extras --steps 5 2>/dev/null | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('variant[$v] iter_ms', round(d['roofline']['launch_ms'],4), 'step', round(d['ms_per_step'],3))"
done; done 2>&1 | tee gpurun_out/r3_probe.log
