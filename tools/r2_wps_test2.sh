python -m pytest tests -m gpu -x -q 2>&1 | tail -8 > gpurun_out/r2_wps_t0.log; cat gpurun_out/r2_wps_t0.log
TTSA_WPS_GRID=1 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_wps_t1.log; cat gpurun_out/r2_wps_t1.log
TTSA_WPS_GRID=3 timeout 600 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_wps_t3.log; cat gpurun_out/r2_wps_t3.log
python bench.py --no-cpu-baseline --steps 5 2>&1 | tail -1 > gpurun_out/r2_wps_bench.json; python -c "
import json; d=json.load(open('gpurun_out/r2_wps_bench.json')); print('iter_ms', d['roofline']['launch_ms'], 'step', d['ms_per_step'], 'value', d['value'])"
TTSA_GL_KERNEL=tile python bench.py --no-cpu-baseline --steps 5 2>&1 | tail -1 | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('tile iter_ms', d['roofline']['launch_ms'], 'step', d['ms_per_step'])"
