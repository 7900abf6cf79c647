import csv,collections,sys
lines=[l for l in open(sys.argv[1]) if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict()
for row in r:
    k=row['Kernel Name'][:64]
    v=float(row['Metric Value'].replace(',',''))
    agg.setdefault(k,[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in agg.items():
    print(f"{k:66s} n={len(v):3d} avg={sum(v)/len(v)/1e6:8.3f} ms share={sum(v)/tot*100:5.1f}%")
