mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
TTSA_WPS_GRID=3 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -8 > gpurun_out/r2s_tests_g3.log; cat gpurun_out/r2s_tests_g3.log
python bench.py --no-cpu-baseline --steps 5 2>&1 | tail -1 > gpurun_out/r2s_bench_stream.json
TTSA_GL_KERNEL=tile python bench.py --no-cpu-baseline --steps 5 2>&1 | tail -1 > gpurun_out/r2s_bench_tile.json
python - <<'PY'
import json
for k in ("stream","tile"):
    try:
        d=json.load(open('gpurun_out/r2s_bench_%s.json'%k)); print(k,'iter_ms', d['roofline']['launch_ms'], 'step', d['ms_per_step'], 'value', d['value'])
    except Exception as e: print(k, 'ERR', e, open('gpurun_out/r2s_bench_%s.json'%k).read()[:500])
PY
