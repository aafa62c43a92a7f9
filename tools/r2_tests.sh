mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
python bench.py --steps 3 --no-cpu-baseline --batch 128 --iters 30 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r2s_bench_cfg4probe.json; tail -3 gpurun_out/bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/r2s_bench_cfg4probe.json')); print(d['value'], d['config']['workload'], d.get('configs4'), d['roofline']['frac'], d.get('latency_single_ms'))"
