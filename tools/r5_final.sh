# session 5, final build of round 2 (warp-uniform scalars): GPU suite, then the bench line
mkdir -p gpurun_out
timeout 100 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 > gpurun_out/r5_final_gputests.log; cat gpurun_out/r5_final_gputests.log
timeout 100 python bench.py --steps 20 --warmup 3 2>gpurun_out/r5_bench_err.log | tail -1 > gpurun_out/r5f_bench.json; tail -2 gpurun_out/r5_bench_err.log
python -c "
import json; d=json.load(open('gpurun_out/r5f_bench.json')); print('value', d['value'], 'e2e', d['e2e']['value'], 'iter_ms', d['roofline']['launch_ms'], 'frac', d['roofline']['frac'])"
