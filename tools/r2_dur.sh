mkdir -p gpurun_out
for v in _p255 ""; do
TTSA_DEBUG=0 TTSA_LIB=$PWD/your-voice-tts_b200/libttsa_b200$v.so ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max --clock-control none -k regex:gl_stream -s 10 -c 20 --csv --log-file gpurun_out/r2_dur$v.csv python bench.py --no-cpu-baseline --steps 1 --warmup 3 --no-graph > /dev/null 2>&1
python - <<PY
import csv
rows=[r for r in csv.reader(open('gpurun_out/r2_dur$v.csv')) if len(r)>5]
h=rows[0]; iK=h.index('Kernel Name'); iM=h.index('Metric Name'); iV=h.index('Metric Value')
import collections
d=collections.defaultdict(list)
for r in rows[1:]:
    d[(r[iK][:40],r[iM])].append(float(r[iV].replace(',','')))
for k,v in d.items(): print('$v',k,len(v),'mean',sum(v)/len(v))
PY
done 2>&1 | tee gpurun_out/r2_dur.log
