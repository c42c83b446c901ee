import csv,collections,re,sys
rows=[r for r in csv.reader(open(sys.argv[1])) if len(r)>5]
hdr=rows[0]
ki=hdr.index('Kernel Name'); vi=hdr.index('Metric Value'); ui=hdr.index('Metric Unit')
agg=collections.OrderedDict()
for r in rows[1:]:
    if r[hdr.index('Metric Name')]!='gpu__time_duration.sum': continue
    name=re.sub(r'\(.*','',r[ki])[:60]
    v=float(r[vi].replace(',',''))
    if r[ui]=='ns': v/=1e3
    elif r[ui]=='ms': v*=1e3
    agg.setdefault(name,[]).append(v)
tot=sum(sum(v) for v in agg.values())
for k,v in sorted(agg.items(), key=lambda kv:-sum(kv[1])):
    print(f"{k:60s} n={len(v):4d} total={sum(v)/1e3:9.3f} ms  mean={sum(v)/len(v):10.1f} us  share={100*sum(v)/tot:5.1f}%")
