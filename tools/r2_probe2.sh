timeout 120 tools/bin/ubench_tmem > gpurun_out/r2_ubench_tmem.log 2>&1; cat gpurun_out/r2_ubench_tmem.log
for d in 0 3 2 6; do TTSA_DEBUG=$d python bench.py --no-cpu-baseline --steps 3 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('debug=$d', 'iter_ms', round(d['roofline']['launch_ms'],4))"; done 2>&1 | tee gpurun_out/r2_probe2.log
