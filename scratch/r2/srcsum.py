import csv, sys
f=sys.argv[1]; top=int(sys.argv[2]) if len(sys.argv)>2 else 40
rows=list(csv.reader(open(f)))
hdr=rows[1]
si=hdr.index('# Samples'); src=hdr.index('Source'); ie=hdr.index('Instructions Executed')
stall_cols=[(i,h) for i,h in enumerate(hdr) if h.startswith('stall_') and 'Not Issued' not in h]
data=[]
for n,r in enumerate(rows[2:]):
    try: s=int(r[si])
    except: continue
    st=sorted(((int(r[i] or 0),h) for i,h in stall_cols), reverse=True)[:2]
    data.append((s,n,r[src].strip(),r[ie],st))
tot=sum(d[0] for d in data)
print('total samples',tot,'instructions',len(data))
for s,n,t,ie,st in sorted(data,reverse=True)[:top]:
    print('%5.1f%% #%4d exec %8s  %-70s %s'%(100.0*s/tot,n,ie,t[:70],' '.join('%s=%d'%(h[6:],v) for v,h in st)))
# cumulative by 100-instruction bucket
print('--- buckets of 50 instructions')
b={}
for s,n,t,ie,st in data: b[n//50]=b.get(n//50,0)+s
for k in sorted(b): 
    if b[k]*100.0/tot>=1.0: print('  insts %4d-%4d: %5.1f%%'%(k*50,k*50+49,100.0*b[k]/tot))
