# Two contexts (own scratch each) on two streams, whole batches alternating between them: how much of
# pass i+1's k_parse (ALU pipe) runs next to pass i's k_lines (shared-memory pipe)?  GTS_LINES_WAVE
# (k_lines CTAs per SM) decides how many registers are left for the other pass's parse warps.
# usage: pipeline_probe.py WORKLOAD [offset]   (offset: the second stream starts half a pass later)
import os, sys
os.environ['GTS_TUNE'] = '1'
sys.path.insert(0, '.')
import torch
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
d_in = torch.from_numpy(arr).cuda()
cap = n * 4 + 4096
for wave, nctx, prio, offset in ((7, 1, 0, 0), (7, 2, 0, 0), (7, 2, 0, 1), (7, 2, 1, 1), (6, 2, 0, 0), (5, 2, 0, 0), (5, 2, 0, 1), (4, 2, 0, 0), (4, 2, 0, 1), (4, 2, 1, 1), (3, 2, 0, 1), (7, 3, 0, 0), (5, 3, 0, 1), (4, 3, 0, 1)):
    os.environ['GTS_LINES_WAVE'] = str(wave)
    trs = [gotranseq_b200.Translator(O, device=0) for _ in range(nctx)]
    outs = [torch.empty(cap, dtype=torch.uint8, device='cuda') for _ in range(nctx)]
    streams = [torch.cuda.Stream(priority=(-1 if (prio and i % 2) else 0)) for i in range(nctx)]
    d_info = torch.empty(4096, dtype=torch.uint8, device='cuda')
    want = trs[0].translate_device(d_in.data_ptr(), n, outs[0].data_ptr(), cap, streams[0].cuda_stream)
    def run(steps):
        for i in range(steps):
            k = i % nctx
            trs[k].translate_device_async(d_in.data_ptr(), n, outs[k].data_ptr(), cap, streams[k].cuda_stream)
    run(2 * nctx)
    spare = gotranseq_b200.Translator(O, device=0)
    spare.piece_scan_async(d_in.data_ptr(), n, d_info.data_ptr(), streams[0].cuda_stream)
    torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    main = torch.cuda.current_stream()
    ev0.record(main)
    for s in streams: s.wait_event(ev0)
    if offset:
        # delay stream k by k front halves (k_parse + k_tile_scan of the same batch on a spare context)
        for k in range(1, nctx):
            for _ in range(k):
                spare.piece_scan_async(d_in.data_ptr(), n, d_info.data_ptr(), streams[k].cuda_stream)
    steps = 48
    run(steps)
    for s in streams:
        e = torch.cuda.Event(); e.record(s); main.wait_event(e)
    ev1.record(main)
    torch.cuda.synchronize()
    ok = all(t.device_result() == want for t in trs)
    same = all(bool(torch.equal(outs[0][:want], o[:want])) for o in outs[1:])
    print(wl, 'wave', wave, 'contexts', nctx, 'prio', prio, 'offset', offset, 'ms/step', round(ev0.elapsed_time(ev1) / steps, 4), 'ok', ok and same, flush=True)
    del trs, outs
    torch.cuda.empty_cache()
