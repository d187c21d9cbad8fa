import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
top=int(sys.argv[3]) if len(sys.argv)>3 else 45
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','cuda,sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
cur=None; hdr=None; agg={}
for r in rows:
    if len(r)>=2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>2 and r[0]=='Line No': hdr=r; ie=r.index('Instructions Executed'); iw=r.index('L1 Wavefronts Shared'); ii=r.index('L1 Wavefronts Shared Ideal'); continue
    if hdr and len(r)>ie and r[0] not in ('','Function Name'):
        try: n=int(r[ie]); w=int(r[iw]) if r[iw].isdigit() else 0; wi=int(r[ii]) if r[ii].isdigit() else 0
        except: continue
        agg[(cur,int(r[0]))]=(n,w,wi,r[1][:100])
tot=sum(v[0] for v in agg.values()); tw=sum(v[1] for v in agg.values())
print('total instr',tot,'smem wavefronts',tw)
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][0])[:top]:
    print(k[0][:16],k[1],v[0],'%.1f%%'%(100*v[0]/tot),'wf',v[1],'ideal',v[2],v[3])
