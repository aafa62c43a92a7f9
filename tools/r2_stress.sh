mkdir -p gpurun_out
python tools/stress_stream.py 25 > gpurun_out/stress_r2_stream.log 2>&1; tail -40 gpurun_out/stress_r2_stream.log
python tools/stress_random.py 12 132 > gpurun_out/stress_r2_random.log 2>&1; tail -3 gpurun_out/stress_r2_random.log
python tools/stress_post.py > gpurun_out/stress_r2_post.log 2>&1; tail -3 gpurun_out/stress_r2_post.log
python tools/stress_streams.py > gpurun_out/stress_r2_streams.log 2>&1; tail -3 gpurun_out/stress_r2_streams.log
