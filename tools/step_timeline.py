import csv,sys
fn=sys.argv[1]
with open(fn) as f: lines=[l for l in f if not l.startswith('==')]
rows=[r for r in csv.DictReader(lines) if r.get('Metric Name')=='gpu__time_duration.sum']
names=[r['Kernel Name'] for r in rows]
idx=[i for i,n in enumerate(names) if 'k_finalize_st' in n]
a,b=idx[-2]+1, idx[-1]+1
tot=0
for r in rows[a:b]:
    v=float(r['Metric Value'].replace(',','')); u=r['Metric Unit']
    if u in ('ns','nsecond'): v/=1000
    tot+=v
    print(f"{v:8.1f}  {r['Kernel Name'][:90]}  grid={r.get('Grid Size','')}")
print('step total us',tot, 'launches', b-a)
