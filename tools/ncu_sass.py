import csv, sys, subprocess
rep, kidx = sys.argv[1], int(sys.argv[2])
out = subprocess.run(["ncu","-i",rep,"--page","source","--csv"],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
ks=[]; cur=None
for r in rows:
    if r and r[0]=='Kernel Name': cur={'name':r[1],'rows':[]}; ks.append(cur); continue
    if r and r[0]=='Address': cur['hdr']=r; continue
    if cur is not None and r: cur['rows'].append(r)
k=ks[kidx]; h=k['hdr']
si=h.index('# Samples'); ei=h.index('Instructions Executed')
stall_cols=[i for i,c in enumerate(h) if c.startswith('stall_')]
tot_s=sum(int(r[si]) for r in k['rows']); tot_e=sum(int(r[ei]) for r in k['rows'])
print(k['name'][:60],'samples',tot_s,'inst',tot_e,'nsass',len(k['rows']))
# stall totals (first occurrence set = all samples)
half=len(stall_cols)//2
tot={}
for r in k['rows']:
    for i in stall_cols[:half]:
        tot[h[i]]=tot.get(h[i],0)+int(r[i] or 0)
print(sorted(tot.items(), key=lambda kv:-kv[1])[:10])
prev=None; start=0; acc=0; accs=0
for i,r in enumerate(k['rows']):
    e=int(r[ei]); s_=int(r[si])
    if prev is not None and e!=prev:
        print(f"[{start:4d},{i:4d}) x{prev:>9d} n={i-start:3d} inst={acc/1e6:8.2f}M samples={accs:6d}  {k['rows'][start][1].strip()[:50]}")
        start=i; acc=0; accs=0
    prev=e; acc+=e; accs+=s_
print(f"[{start:4d},{len(k['rows']):4d}) x{prev:>9d} inst={acc/1e6:8.2f}M samples={accs}")
