# A/B two builds of the library on the same box: current vs tools/lib_old.so
probe() { python bench.py --no-cpu-baseline --steps 5 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'iter_ms', round(d['roofline']['launch_ms'],4), 'step_ms', round(d['ms_per_step'],3))"; }
cp your-voice-tts_b200/libttsa_b200.so /tmp/cur.so
for i in 1 2; do
  cp /tmp/cur.so your-voice-tts_b200/libttsa_b200.so; touch your-voice-tts_b200/build/fingerprint; probe new
  cp tools/lib_old.so your-voice-tts_b200/libttsa_b200.so; probe old
done
cp /tmp/cur.so your-voice-tts_b200/libttsa_b200.so
