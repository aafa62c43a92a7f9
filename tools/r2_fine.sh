mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2s_tests.log; cat gpurun_out/r2s_tests.log
python tools/time_single.py 2>&1 | tail -5
TTSA_GL_FINE=0 python tools/time_single.py 2>&1 | tail -5 | head -3
