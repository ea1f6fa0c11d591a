import csv,collections,sys
fn=sys.argv[1]; skip=int(sys.argv[2]) if len(sys.argv)>2 else 0
with open(fn) as f:
    lines=[l for l in f if not l.startswith('==')]
r=csv.DictReader(lines)
agg=collections.OrderedDict(); tot=0; n=0
for row in r:
    if row.get('Metric Name')!='gpu__time_duration.sum': continue
    n+=1
    if n<=skip: continue
    name=row['Kernel Name'][:80]; v=float(row['Metric Value'].replace(',',''))
    unit=row['Metric Unit']
    if unit in('ns','nsecond'): v/=1000
    elif unit in ('ms','msecond'): v*=1000
    agg.setdefault(name,[0,0.0]); agg[name][0]+=1; agg[name][1]+=v; tot+=v
print(n,'launches; summed us',round(tot,1))
for k,(c,v) in sorted(agg.items(), key=lambda kv:-kv[1][1])[:int(sys.argv[3]) if len(sys.argv)>3 else 22]:
    print(f"{v:10.1f} us {100*v/tot:5.1f}% n={c:4d} avg={v/c:8.1f}  {k}")
