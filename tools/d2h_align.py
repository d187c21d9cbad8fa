import torch, time
n=34511414
h=torch.empty(400_000_000,dtype=torch.uint8).pin_memory()
d=torch.empty(64_000_000,dtype=torch.uint8,device='cuda')
hin=torch.empty(200_000_000,dtype=torch.uint8).pin_memory()
din=torch.empty(200_000_000,dtype=torch.uint8,device='cuda')
s2=torch.cuda.Stream()
def run(doff, soff, with_h2d, size=n, reps=11):
    torch.cuda.synchronize()
    t0=time.perf_counter()
    if with_h2d:
        with torch.cuda.stream(s2):
            for o in range(0,189_000_000,16<<20): din[o:o+(16<<20)].copy_(hin[o:o+(16<<20)],non_blocking=True)
    off=doff
    for i in range(reps):
        h[off:off+size].copy_(d[soff:soff+size],non_blocking=True)
        off+=size
    torch.cuda.current_stream().synchronize()
    dt=time.perf_counter()-t0
    torch.cuda.synchronize()
    return reps*size/dt/1e9
for name,doff,soff,size in (('aligned',0,0,34511360),('dst+1 size odd',1,0,n),('dst+1 src+1 (congruent)',1,1,n),('dst+64',64,0,34511360),('dst+4',4,0,34511360),('dst+16',16,0,34511360)):
    for w in (False,True):
        r=[run(doff,soff,w,size) for _ in range(3)]
        print(name,'with_h2d' if w else 'alone', [round(x,1) for x in r])
