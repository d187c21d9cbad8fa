import csv,sys,subprocess
rep,kern=sys.argv[1],sys.argv[2]
out=subprocess.run(['ncu','-i',rep,'--page','source','--csv','--print-source','sass','--kernel-name','regex:'+kern],capture_output=True,text=True).stdout
rows=list(csv.reader(out.splitlines()))
hdr=None
for i,r in enumerate(rows):
    if 'Address' in r and 'Source' in r: hdr=r; start=i+1; break
cols=[j for j,n in enumerate(hdr) if n.startswith('stall_') and 'Not Issued' not in n]
tot={hdr[j]:0 for j in cols}
for r in rows[start:]:
    if len(r)<len(hdr): continue
    for j in cols:
        try: tot[hdr[j]]+=int(r[j])
        except: pass
s=sum(tot.values())
for k,v in sorted(tot.items(), key=lambda kv:-kv[1]): 
    if v: print(k, v, '%.1f%%'%(100*v/s))
# top SASS lines by samples
isamp=hdr.index('# Samples'); isrc=hdr.index('Source')
top=sorted([r for r in rows[start:] if len(r)>=len(hdr) and r[isamp].isdigit()], key=lambda r:-int(r[isamp]))[:25]
for r in top: print(r[isamp], r[isrc][:80], {hdr[j]:r[j] for j in cols if r[j] not in ('0','')})
