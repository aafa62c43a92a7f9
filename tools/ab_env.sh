# A/B one build under an environment toggle on the same box:  tools/ab_env.sh VAR=VALUE
probe() { env $2 python bench.py --no-cpu-baseline --steps 5 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$1', 'iter_ms', round(d['roofline']['launch_ms'],4), 'step_ms', round(d['ms_per_step'],3), 'init_ms', round(d['roofline']['init_synthesis_ms'],4))"; }
for i in 1 2; do probe default ""; probe "$1" "$1"; done
