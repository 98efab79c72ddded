#!/usr/bin/env python3
"""Groups the SASS lines of an ncu source page (`ncu -i report.ncu-rep --page source --csv`) into straight-line blocks with equal execution
counts and prints, for the blocks holding most stall samples, instructions per element, the sample share, the dominant opcodes and
stall reasons.  Usage: python scripts/ncu_sass_blocks.py src_sass.csv [min sample share, default 0.008] (E = elements per launch below)."""
import csv,collections,sys
rows=list(csv.reader(open(sys.argv[1])))
h=rows[1]; ix={c:i for i,c in enumerate(h)}
E=2000376
stallcols=[c for c in h if c.startswith('stall_') and 'Not Issued' not in c]
blocks=[]
cur=None
for n,r in enumerate(rows[2:]):
    ex=int(r[ix['Instructions Executed']] or 0); sm=int(r[ix['# Samples']] or 0)
    s=r[ix['Source']].strip().split()
    op=(s[1] if s[0].startswith('@') else s[0])
    if cur is None or abs(cur['ex']-ex)>0.02*max(ex,1) or op.startswith('BAR') :
        cur={'start':n,'ex':ex,'n':0,'samples':0,'ops':collections.Counter(),'st':collections.Counter()}
        blocks.append(cur)
    cur['n']+=1; cur['samples']+=sm; cur['ops'][op.split('.')[0]]+=1
    for c in stallcols:
        if r[ix[c]]: cur['st'][c[6:]]+=int(r[ix[c]])
tot=sum(b['samples'] for b in blocks)
print('total samples',tot)
thr=float(sys.argv[2]) if len(sys.argv)>2 else 0.008
for b in blocks:
    if b['samples']<tot*thr: continue
    print(f"{b['start']:5d} n={b['n']:4d} exec/elem={b['ex']/E:6.3f} instr/elem={b['ex']*b['n']/E:6.1f} samples={100*b['samples']/tot:5.1f}%  {dict(b['ops'].most_common(5))} {dict((k,round(100*v/max(1,sum(b['st'].values())))) for k,v in b['st'].most_common(3))}")
