"""Opcode histogram (executed warp-instructions per frame) of one captured kernel, split by source region.
usage: opcode_hist.py <nvdisasm -g dump> <ncu source csv> [frames]"""
import re, csv, collections, sys
sass, ncucsv = sys.argv[1], sys.argv[2]
nfr = float(sys.argv[3]) if len(sys.argv) > 3 else 32000.0
cur = None; seq = {}
for ln in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: seq[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(ncucsv)))
hdr = rows[1]; iE = hdr.index("Instructions Executed")
inst = []
for r in rows[2:]:
    if r[0] in ("Kernel Name", "Address"): break
    inst.append((int(r[0], 16), r[1].strip(), int(r[iE])))
base = inst[0][0]
op = collections.Counter(); opl = collections.defaultdict(collections.Counter)
FP = {'FFMA2', 'FADD2', 'FMUL2', 'FFMA', 'FADD', 'FMUL'}
nonfp = collections.Counter()
for a, src, e in inst:
    l, txt = seq.get(a - base, (None, ''))
    t = src.split()
    o = t[1] if t[0].startswith('@') else t[0]
    o = o.split('.')[0]
    op[o] += e; opl[o][l] += e
    if o not in FP: nonfp[l] += e
tot = sum(op.values())
print('total/frame %.1f' % (tot / nfr))
for o, c in op.most_common(30): print(f"{o:12s} {c/nfr:8.1f}/frame {100*c/tot:5.1f}%")
print('non-FP by line:')
for l, c in nonfp.most_common(45): print(f"  {str(l):38s} {c/nfr:7.1f}")
