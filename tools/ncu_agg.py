import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg={}
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>2 and r[0]=='Line No': hdr=r; ie=r.index('Instructions Executed'); continue
    if hdr and len(r)>ie and r[0] not in ('','Function Name'):
        try: n=int(r[ie])
        except: continue
        agg[(cur,int(r[0]))]=agg.get((cur,int(r[0])),0)+n
tot=sum(agg.values())
print('total',tot)
files={}
for (f,l),n in agg.items(): files.setdefault(f,[]).append((l,n))
for f,v in files.items():
    print(f, sum(n for l,n in v), '%.1f%%'%(100*sum(n for l,n in v)/tot))
ranges=eval(sys.argv[3])
for name,(f,a,b) in ranges.items():
    s=sum(n for (ff,l),n in agg.items() if ff.startswith(f) and a<=l<=b)
    print(name, s, '%.1f%%'%(100*s/tot))
