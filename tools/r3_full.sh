mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r3_tests_full.log; cat gpurun_out/r3_tests_full.log
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r3_bench_err.log | tail -1 > gpurun_out/r3_bench.json; tail -2 gpurun_out/r3_bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/r3_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'iter_ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], 'launches', d['gpu_launches'], 'lat1', d.get('latency_single_ms'), 'dropin', d.get('e2e_dropin',{}).get('value'))"
