import csv,collections,sys,subprocess
rep=sys.argv[1]; skip=sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--launch-skip',skip,'--launch-count','1'],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
print(rows[0][1][:80])
hdr=rows[1]; idx={h:i for i,h in enumerate(hdr)}
seen=set(); data=[]
for r in rows[2:]:
    if len(r)>idx['Avg. Threads Executed'] and r[0].startswith('0x') and r[0] not in seen:
        seen.add(r[0]); data.append(r)
f=lambda r,k:int(float(r[idx[k]] or 0))
tot_s=sum(f(r,'# Samples') for r in data); tot_i=sum(f(r,'Instructions Executed') for r in data)
print('instr rows',len(data),'samples',tot_s,'warp-instr executed',tot_i, 'per elem-batch', tot_i/(4096*5000/32))
stalls=[h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg=collections.Counter()
for r in data:
    for h in stalls: agg[h]+=f(r,h)
print(agg.most_common(10))
top=sorted(data,key=lambda r:-f(r,'# Samples'))[:int(sys.argv[3]) if len(sys.argv)>3 else 25]
for r in top:
    st=sorted([(f(r,h),h) for h in stalls],reverse=True)[:2]
    print(str(f(r,'# Samples')).rjust(6), str(f(r,'Instructions Executed')).rjust(9), r[idx['Avg. Threads Executed']].rjust(4), r[idx['Source']][:70].ljust(70), st)
h=collections.Counter(); hs=collections.Counter()
for r in data:
    t=r[idx['Source']].split()
    op=t[1] if t[0].startswith('@') else t[0]
    op=op.split('.')[0]
    h[op]+=f(r,'Instructions Executed'); hs[op]+=f(r,'# Samples')
print([(k,v,hs[k]) for k,v in h.most_common(24)])
