"""Join an ncu SASS source page (csv) with nvdisasm -g line info: executed instructions and stall samples per source line."""
import re, csv, collections, sys
sass, ncucsv = sys.argv[1], sys.argv[2]
topn = int(sys.argv[3]) if len(sys.argv) > 3 else 40
cur=None; seq={}
for ln in open(sass):
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',ln)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2))); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',ln)
    if m: seq[int(m.group(1),16)]=(cur,m.group(2))
rows=list(csv.reader(open(ncucsv)))
hdr=rows[1]; iE=hdr.index("Instructions Executed"); iSm=hdr.index("# Samples")
inst=[]
for r in rows[2:]:
    if r[0] in("Kernel Name","Address"): break
    inst.append((int(r[0],16),r[1].strip(),int(r[iE]),int(r[iSm])))
base=inst[0][0]
agg=collections.Counter(); sm=collections.Counter()
for a,src,e,s in inst:
    l=seq.get(a-base,(None,''))[0]
    agg[l]+=e; sm[l]+=s
tot=sum(agg.values()); ts=sum(sm.values())
print('total executed',tot,'samples',ts)
for l,c in sorted(agg.items(), key=lambda x:-x[1])[:topn]:
    print(f"{str(l):40s} {c:12d} {100*c/tot:5.1f}%  samples {100*sm[l]/ts:5.1f}%")
