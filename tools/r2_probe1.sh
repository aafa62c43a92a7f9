mkdir -p gpurun_out
for d in 0 2 6 4 1; do TTSA_DEBUG=$d python bench.py --no-cpu-baseline --steps 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('debug=$d', 'iter_ms', round(d['roofline']['launch_ms'],4))"; done > gpurun_out/r2_probe1.log 2>&1
cat gpurun_out/r2_probe1.log
