#!/usr/bin/env python
"""Per source line warp-stall samples of one profiled kernel: joins `ncu -i X.ncu-rep --page source --csv` (SASS rows,
in address order) with `nvdisasm -g -c` of the same cubin (SASS + //## File/line markers).

    cuobjdump -xelf all mint_b200/libfact_sm100.so; nvdisasm -g -c sdpa_tc.sm_100a.cubin > sdpa_tc.sass
    python scripts/ncu_stalls_by_line.py sdpa_tc.sass FIRST_LINE LAST_LINE source.csv sdpa_tc.cu [N]
(FIRST_LINE / LAST_LINE: the line range of the kernel's .text section inside the .sass file.)"""
import re,csv,sys
from collections import defaultdict
sass_file, lo, hi, csv_file, srcname = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5]
lines=open(sass_file).read().splitlines()[lo:hi]
cur=None; ins=[]
for l in lines:
    m=re.search(r'//## File "([^"]+)", line (\d+)(.*)',l)
    if m:
        cur=(m.group(1).split('/')[-1],int(m.group(2)),m.group(3)); continue
    m=re.match(r'\s+/\*([0-9a-f]{4,})\*/\s+(.*?);',l)
    if m: ins.append((cur,m.group(2)))
rows=list(csv.reader(open(csv_file)))
hdr=rows[1]; ix={h:i for i,h in enumerate(hdr)}
data=[r for r in rows[2:] if len(r)>10]
print(len(ins), len(data))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg=defaultdict(lambda: defaultdict(int)); tot=0
for (loc,sass),r in zip(ins,data):
    try: s=int(r[ix['# Samples']])
    except: continue
    key=(loc[0],loc[1]) if loc else None
    agg[key]['n']+=s; tot+=s
    for st in stalls:
        agg[key][st]+=int(r[ix[st]] or 0)
print('total',tot)
files={}
def text(key):
    if not key: return ''
    import os
    for d in ('/root/repo/mint_b200/csrc/',):
        p=d+key[0]
        if os.path.exists(p):
            if p not in files: files[p]=open(p).read().splitlines()
            f=files[p]
            return f[key[1]-1].strip()[:85] if key[1]-1<len(f) else ''
    return ''
for key,d in sorted(agg.items(),key=lambda kv:-kv[1]['n'])[:int(sys.argv[6]) if len(sys.argv)>6 else 28]:
    st={k.replace('stall_',''):v for k,v in d.items() if k!='n' and v>0.12*d['n']}
    print(f"{d['n']:5d} {100*d['n']/tot:5.1f}% {key} | {text(key)} | {st}")
