# usage: bash tools/r3_multi.sh N   -- configs[4] cells (4096 utterances over N GPUs, 30 and 60 iterations) + the default
# weak-scaling line with the optional final NCCL gather; lines land in gpurun_out/r3_multi_N*.json
N=$1; mkdir -p gpurun_out
PER=$((4096 / N))
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29571 bench.py --gpus $N "$@" 2>gpurun_out/r3_multi_err.log | grep "^{" | tail -1; }
run --steps 10 --warmup 3 --no-cpu-baseline --no-extras --gather > gpurun_out/r3_multi_${N}gpu_default.json
run --steps 3 --warmup 3 --no-cpu-baseline --no-extras --batch $PER --iters 60 > gpurun_out/r3_multi_${N}gpu_cfg5_60.json
run --steps 3 --warmup 3 --no-cpu-baseline --no-extras --batch $PER --iters 30 > gpurun_out/r3_multi_${N}gpu_cfg5_30.json
tail -3 gpurun_out/r3_multi_err.log
python - <<PY
import json
for k in ("default","cfg5_60","cfg5_30"):
    try:
        d=json.load(open("gpurun_out/r3_multi_${N}gpu_%s.json"%k)); print(k, "value", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ms/step", round(d["ms_per_step"],2), "gather", d.get("gather",{}).get("ms"), d.get("gather",{}).get("GBps_per_rank"), d.get("configs4"))
    except Exception as e: print(k, "ERR", e)
PY
