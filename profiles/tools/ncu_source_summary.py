import csv,collections,sys,re
rows=list(csv.reader(open(sys.argv[1])))
hdr=rows[1]
ix={h:i for i,h in enumerate(hdr)}
op=collections.Counter(); stall=collections.Counter(); samples=collections.Counter()
tot=0
stall_cols=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
lines=[]
for r in rows[2:]:
    if len(r)<len(hdr): continue
    src=r[ix['Source']]
    n=int(r[ix['Instructions Executed']] or 0)
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)',src)
    o=m.group(2) if m else src[:10]
    base='.'.join(o.split('.')[:2]) if o.startswith(('LD','ST')) else o.split('.')[0]
    op[base]+=n; tot+=n
    s=int(r[ix['# Samples']] or 0)
    samples[base]+=s
    for c in stall_cols:
        stall[c]+=int(r[ix[c]] or 0)
    lines.append((s,n,src))
print('total warp instr',tot)
for k,v in op.most_common(25): print(f'{k:14s} {v:12d} {100*v/tot:5.1f}%  samples {samples[k]}')
ts=sum(stall.values())
print('stalls:')
for k,v in stall.most_common(10): print(f'  {k:28s} {100*v/ts:5.1f}%')
if len(sys.argv)>2:
    lines.sort(reverse=True)
    for s,n,src in lines[:int(sys.argv[2])]: print(s,n,src)
