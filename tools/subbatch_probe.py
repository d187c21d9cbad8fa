# how much would K record-aligned sub-batches on concurrent streams (own scratch each) gain over one pass?
import sys, time
sys.path.insert(0, '.')
import numpy as np, torch
import bench, gotranseq_b200, workloads
from gotranseq_b200 import _lib
L = _lib.lib()
wl = sys.argv[1] if len(sys.argv) > 1 else 'readme189'
b = bench.Batch(wl, 0, 1, L)
arr = b.array()
_, opts = workloads.WORKLOADS[wl]
O = gotranseq_b200.Options(Frame=opts.get("frame", "1"), Table=opts.get("table", 0), Clean=opts.get("clean", False),
                           Alternative=opts.get("alternative", False), Trim=opts.get("trim", False))
n = b.nbytes
def cuts(K):
    c = [0]
    for k in range(1, K):
        p = n * k // K
        while not (arr[p] == 62 and arr[p - 1] == 10): p += 1
        c.append(p)
    c.append(n)
    return c
for K, nstreams in ((1, 1), (2, 2), (4, 2), (4, 4), (8, 2), (8, 4), (8, 8)):
    c = cuts(K)
    trs = [gotranseq_b200.Translator(O, device=0) for _ in range(K)]
    ins = [torch.from_numpy(arr[c[i]:c[i + 1]].copy()).cuda() for i in range(K)]
    outs = [torch.empty((c[i + 1] - c[i]) * 4 + 4096, dtype=torch.uint8, device='cuda') for i in range(K)]
    streams = [torch.cuda.Stream() for _ in range(nstreams)]
    def step():
        for i in range(K):
            s = streams[i % nstreams]
            trs[i].translate_device_async(ins[i].data_ptr(), c[i + 1] - c[i], outs[i].data_ptr(), outs[i].numel(), s.cuda_stream)
    for _ in range(3): step()
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    ev0.record(main)
    for s in streams: s.wait_event(ev0)
    steps = 20
    for _ in range(steps): step()
    for s in streams:
        e = torch.cuda.Event(); e.record(s); main.wait_event(e)
    ev1.record(main)
    torch.cuda.synchronize()
    tot = sum(t.device_result() for t in trs)
    print(wl, 'K', K, 'streams', nstreams, 'ms/step', round(ev0.elapsed_time(ev1) / steps, 4), 'out', tot)
    del trs, ins, outs
    torch.cuda.empty_cache()
