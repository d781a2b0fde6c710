import csv, collections, re, sys
path = sys.argv[1]
rows=[]
with open(path) as f:
    lines=[l for l in f if not l.startswith('==')]
for r in csv.DictReader(lines):
    if r.get('Metric Name')=='gpu__time_duration.sum':
        rows.append((int(r['ID']), r['Kernel Name'], float(r['Metric Value'].replace(',','')), r.get('Grid Size'), r.get('Block Size')))
agg=collections.defaultdict(lambda:[0,0.0,None])
for i,k,v,g,b in rows:
    name=re.sub(r'\(.*','',k)
    agg[name][0]+=1; agg[name][1]+=v; agg[name][2]=(g,b)
tot=sum(v[1] for v in agg.values())
print(f"{len(rows)} launches, total {tot/1e3:.1f} us")
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[2]) if len(sys.argv)>2 else 30]:
    print(f"{k[-58:]:58s} n={v[0]:5d} total={v[1]/1e3:10.1f}us avg={v[1]/v[0]/1e3:8.2f}us share={v[1]/tot*100:5.1f}%  grid/block={v[2]}")
