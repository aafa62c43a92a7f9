# tools/profile_report.sh <rep-basename-in-gpurun_out> <mangled-kernel-substring> <cu-file-stem>
# per-region / per-line executed-instruction report of one captured kernel (uses the in-tree .so for line info)
set -e
rep=gpurun_out/$1.ncu-rep; sub=$2; stem=${3:-frame_gl}
ncu -i $rep --page source --csv --print-source sass > gpurun_out/$1_src.csv 2>/dev/null
ncu -i $rep --page raw --csv > gpurun_out/$1_raw.csv 2>/dev/null
rm -rf /tmp/cub && mkdir -p /tmp/cub && (cd /tmp/cub && cuobjdump -xelf $stem /root/repo/your-voice-tts_b200/libttsa_b200.so > /dev/null)
start=$(nvdisasm -g /tmp/cub/$stem.sm_100a.cubin 2>/dev/null | grep -n "^\.text\..*$sub" | head -1 | cut -d: -f1)
nvdisasm -g /tmp/cub/$stem.sm_100a.cubin 2>/dev/null | tail -n +$start | awk 'NR>1 && /^\.text\./{exit} {print}' > /tmp/cub/k.sass
python tools/ncu_regions.py /tmp/cub/k.sass gpurun_out/$1_src.csv gpurun_out/$1_raw.csv
