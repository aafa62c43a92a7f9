mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r3_tests_final.log; cat gpurun_out/r3_tests_final.log
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 3 2>gpurun_out/r3_bench_err.log | tail -1 > gpurun_out/r3f_bench.json; tail -2 gpurun_out/r3_bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/r3f_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'iter_ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'], 'fp32', d['roofline_fp32']['frac'], 'launches', d['gpu_launches'], 'lat1', d.get('latency_single_ms'), 'dropin', d.get('e2e_dropin',{}).get('value'), 'cpu', d['cpu_baseline']['value'])"
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
