"""profiles/gl_iter_traffic.json from one `ncu --set full` capture of the Griffin-Lim iteration kernel (64 x 482):
DRAM bytes per launch (roofline.traffic) and FP32-pipe busy cycles per SM per launch (roofline_fp32).
    python tools/make_gl_profile.py gpurun_out/r2f_gl.ncu-rep profiles/gl_iter_traffic.json"""
import csv, io, json, subprocess, sys
rep, out = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
def col(name, r):
    i = hdr.index(name)
    v = float(r[i].replace(",", ""))
    u = units[i].lower()
    return v * {"mbyte": 1e6, "kbyte": 1e3, "gbyte": 1e9, "byte": 1.0}.get(u, 1.0)
n = len(data)
rd = sum(col("dram__bytes_read.sum", r) for r in data) / n
wr = sum(col("dram__bytes_write.sum", r) for r in data) / n
act = sum(col("sm__cycles_active.avg", r) for r in data) / n
pct = sum(col("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", r) for r in data) / n
json.dump({"kernel": data[0][hdr.index("Kernel Name")], "report": rep.split("/")[-1], "launches_averaged": n,
           "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
           "algorithmic_bytes_per_launch": 194201600,
           "sm_cycles_active_avg": act, "fma_pipe_pct_of_active": pct, "fma_pipe_cycles_per_sm": act * pct / 100.0,
           "sm_mhz_nominal": 1965.0,
           "source": "ncu --set full --clock-control none, 64 utterances x 482 frames, one B200 (tools/r2_final_profile.sh)"},
          open(out, "w"), indent=1)
print(open(out).read())
