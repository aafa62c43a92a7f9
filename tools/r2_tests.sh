mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
python bench.py --steps 5 2>gpurun_out/bench_err.log | tail -1 > gpurun_out/r2s_bench_full.json; tail -3 gpurun_out/bench_err.log
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2s_bench_full.json'))
for k in ('value','ms_per_step','e2e','e2e_dropin','latency_single_ms','roofline','roofline_fp32','cpu_baseline','clocks'): print(k, d.get(k))
PY
