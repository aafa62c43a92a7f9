"""Per-source-line and per-opcode profile of one captured kernel.
usage: sass_profile.py <lib.so> <cu-stem> <mangled-substring> <report.ncu-rep> [topn]
Joins `ncu --page source --print-source sass` (executed instructions, stall samples per SASS instruction) with the
line info of `nvdisasm -g` on the in-tree library, and prints executed warp-instructions per source line, per opcode
and stall reasons per source line."""
import collections, csv, io, os, re, subprocess, sys, tempfile
lib, stem, sub, rep = sys.argv[1:5]
topn = int(sys.argv[5]) if len(sys.argv) > 5 else 45
tmp = tempfile.mkdtemp()
subprocess.run("cd %s && cuobjdump -xelf %s %s > /dev/null" % (tmp, stem, os.path.abspath(lib)), shell=True, check=True)
cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
dis = subprocess.run(["nvdisasm", "-g", os.path.join(tmp, cubin)], capture_output=True, text=True).stdout.split("\n")
start = next(i for i, l in enumerate(dis) if l.startswith(".text.") and sub in l)
cur = None; seq = {}
for ln in dis[start + 1:]:
    if ln.startswith(".text."): break
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: seq[int(m.group(1), 16)] = (cur, m.group(2))
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[1]; iE = hdr.index("Instructions Executed"); iS = hdr.index("# Samples")
stall = {h[6:]: i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h}
iW = hdr.index("L1 Wavefronts Shared")
line_e = collections.Counter(); line_s = collections.Counter(); op_e = collections.Counter(); op_s = collections.Counter()
line_w = collections.Counter(); line_st = collections.defaultdict(collections.Counter)
base = None
for r in rows[2:]:
    if r[0] in ("Kernel Name", "Address"): break
    a = int(r[0], 16)
    if base is None: base = a
    l = seq.get(a - base, (None, ''))[0]
    e, s = int(r[iE]), int(r[iS])
    op = r[1].strip().split()
    op = op[1] if op[0].startswith('@') else op[0]
    op = op.split('.')[0] + ('.' + op.split('.')[1] if op.startswith(('LDS', 'STS', 'LDG', 'STG', 'LDL', 'STL')) and '.' in op else '')
    line_e[l] += e; line_s[l] += s; op_e[op] += e; op_s[op] += s; line_w[l] += int(r[iW] or 0)
    for k, i in stall.items(): line_st[l][k] += int(r[i] or 0)
tot = sum(line_e.values()); ts = sum(line_s.values())
nf = float(os.environ.get("FRAMES", "31440"))
print("executed warp-instructions %d (%.0f per frame at %d frames), samples %d" % (tot, tot / nf, nf, ts))
print("--- per source line")
for l, c in line_e.most_common(topn):
    top = ", ".join("%s %d" % kv for kv in line_st[l].most_common(3))
    print("%-28s %10d %5.1f%% /frame %6.1f  samples %5.1f%%  smem-wavefronts/frame %6.1f  [%s]" % (str(l), c, 100 * c / tot, c / nf, 100 * line_s[l] / ts, line_w[l] / nf, top))
print("--- per opcode")
for o, c in op_e.most_common(30):
    print("%-12s %10d %5.1f%% /frame %6.1f  samples %5.1f%%" % (o, c, 100 * c / tot, c / nf, 100 * op_s[o] / ts))
