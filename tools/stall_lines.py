"""Stall samples per source line and reason for one captured kernel.
usage: stall_lines.py <nvdisasm -g dump> <ncu source csv> [reason ...]   (reasons: long_sb short_sb barrier wait math mio ...)"""
import re, csv, collections, sys
sass, ncucsv = sys.argv[1], sys.argv[2]
reasons = sys.argv[3:] or ['long_sb', 'short_sb', 'barrier', 'wait', 'math', 'mio', 'not_selected', 'dispatch', 'no_inst', 'branch_resolving']
cur = None; seq = {}
for ln in open(sass):
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', ln)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2))); continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);', ln)
    if m: seq[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(ncucsv)))
hdr = rows[1]
idx = {r: hdr.index('stall_' + r) for r in reasons}
iS = hdr.index('# Samples')
tot = collections.Counter(); per = {r: collections.Counter() for r in reasons}; allS = 0
base = None
for r in rows[2:]:
    if r[0] in ("Kernel Name", "Address"): break
    a = int(r[0], 16)
    if base is None: base = a
    l = seq.get(a - base, (None, ''))[0]
    allS += int(r[iS])
    for k, i in idx.items():
        v = int(r[i] or 0); per[k][l] += v; tot[k] += v
print('all samples', allS)
for k in reasons:
    print(f"== {k}: {tot[k]} ({100*tot[k]/allS:.1f}%)")
    for l, c in per[k].most_common(8): print(f"    {str(l):40s} {c:6d}")
