import csv,sys,collections,re
rows=list(csv.reader(open(sys.argv[1])))
# multiple kernel instances: split at "Kernel Name" rows; use the first
blocks=[];cur=None
for r in rows:
    if r and r[0]=='Kernel Name':
        cur={'name':r[1],'hdr':None,'data':[]}; blocks.append(cur); continue
    if cur is None: continue
    if cur['hdr'] is None: cur['hdr']=r; continue
    if len(r)==len(cur['hdr']): cur['data'].append(r)
b=blocks[0]; hdr=b['hdr']; data=b['data']
print(b['name'][:100], 'instances', len(blocks))
ia=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples'); isrc=hdr.index('Source')
tot=sum(int(r[ia]) for r in data); ts=sum(int(r[isamp]) for r in data)
print('total warp inst',tot,'samples',ts,'n sass',len(data))
op=collections.Counter(); ops=collections.Counter()
for r in data:
    m=re.match(r'\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)', r[isrc])
    k=m.group(2).split('.')[0] if m else '?'
    op[k]+=int(r[ia]); ops[k]+=int(r[isamp])
for k,v in op.most_common(16): print(f'{k:10s} inst {v/tot*100:5.1f}%  samples {ops[k]/ts*100:5.1f}%')
for r in sorted(data,key=lambda r:-int(r[isamp]))[:12]: print(r[isamp], r[ia], r[isrc][:100])
