#!/usr/bin/env python
"""bench.py -- the translation hot path on synthetic FASTA of the README benchmark shape.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload readme189] [--impl ours|reference]

A step is one pass of the whole hot path (parse -> encode -> translate all six frames -> format)
over one batch of FASTA (BASELINE.json configs[1]: 189 MB, multi-record, -f 6 -t 0).
  value   device-resident Gbp/s (input and output in HBM, CUDA events on the launch stream)
  e2e     the same metric through gts_translate_buffer with pinned HOST buffers, PCIe copies timed
  roofline  the dominant kernel (k_lines) against the measured HBM copy rate
  cpu_baseline / --impl reference  the restated reference (oracle/, all host threads)
Under torchrun (N > 1) every rank translates its own batch (independent shards, no collective);
the time is the max over ranks and the value the sum of the nucleotides over that time.
"""
import argparse
import ctypes
import json
import os
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "6-frame Gbp/s device-resident"
# DRAM bytes of one k_lines launch on the default workload, from the committed ncu capture
# (profiles/r1h_ncu_full_summary.txt: 140.83 MB read + 334.67 MB written); null for other workloads
NCU_TRAFFIC = {"readme189": 140825600 + 334672640}
UNIT = "Gbp/s"


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler(threading.Thread):
    """Samples SM clock and throttle reasons of one GPU through NVML while a timed region runs."""

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop_evt = threading.Event()
        self.ok = False
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
            self.ok = True
        except Exception:
            self.ok = False

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop_evt.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()
        self.join(timeout=2)
        s = sorted(self.samples)
        return {
            "sm_mhz": s[len(s) // 2] if s else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(s),
        }


def env_rank():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def bind_to_gpu_numa_node(index):
    """Run this rank on the CPUs next to its GPU (NVML's affinity mask), before any pinned memory is
    allocated: first touch then puts the host buffers on the GPU's NUMA node.  Returns the CPU count
    or None when NVML / the scheduler call is not available."""
    try:
        import pynvml
        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        words = (os.cpu_count() + 63) // 64
        mask = pynvml.nvmlDeviceGetCpuAffinity(h, words)
        cpus = {64 * w + b for w, m in enumerate(mask) for b in range(64) if (m >> b) & 1}
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
            return len(cpus)
    except Exception:
        pass
    return None


def record_cut(buf, limit):
    """Largest prefix of buf of at most `limit` bytes that ends at a record boundary."""
    if len(buf) <= limit:
        return len(buf)
    i = buf.rfind(b"\n>", 0, limit)
    return i + 1 if i > 0 else len(buf)


def count_nt(data, opts):
    """Nucleotides of the batch (what Gbp/s counts): sequence bytes of the synthetic FASTA."""
    import numpy as np
    a = np.frombuffer(data, dtype=np.uint8)
    nl = int((a == 10).sum())
    hdr_starts = np.flatnonzero(a == ord(">"))
    # synthetic workloads: '>' only starts header lines; header bytes = up to the next newline
    nls = np.flatnonzero(a == 10)
    ends = nls[np.searchsorted(nls, hdr_starts)]
    hdr_bytes = int((ends - hdr_starts).sum())
    return len(a) - nl - hdr_bytes


def run_reference(args, rank, world):
    """The restated reference (oracle/, see its header) on the host cores: bounded sample per step."""
    if rank != 0:
        return
    import numpy as np
    import oracle
    import workloads
    gen, opts = workloads.WORKLOADS[args.workload]
    data = gen()
    cores = os.cpu_count() or 1
    cut = record_cut(data, args.ref_sample_bytes)
    sample = np.frombuffer(data[:cut], dtype=np.uint8)
    nt = count_nt(data[:cut], opts)
    for _ in range(args.warmup):
        oracle.translate(sample, num_worker=cores, **opts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.translate(sample, num_worker=cores, **opts)
    dt = (time.perf_counter() - t0) / args.steps
    v = nt / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, len(data)),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "first %d bytes (%d nt) of the workload per step, restated reference "
                                   "(oracle/transeq_oracle.c: 1 reader + %d workers)" % (cut, nt, cores)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def workload_config(args, nbytes):
    import workloads
    _, opts = workloads.WORKLOADS[args.workload]
    flags = "-f %s -t %d%s%s%s" % (opts.get("frame", "1"), opts.get("table", 0), " -c" if opts.get("clean") else "",
                                    " -a" if opts.get("alternative") else "", " -T" if opts.get("trim") else "")
    return {"workload": "%s: %d-byte synthetic FASTA per GPU (%s), flags %s" % (
                args.workload, nbytes, workloads.WORKLOADS[args.workload][0].__doc__.split("--")[0].strip().rstrip(":"),
                flags),
            "l2": "input + output per step exceed the 126 MB L2 several times over; no explicit flush",
            "sharding": "one independent batch per GPU, no collective"}


def run_split(args, rank, local_rank, world):
    """--workload chromosome_split (BASELINE config 4 at N GPUs): ONE record cut into N pieces, one per
    rank.  A step = every rank scans its piece, the ranks all-gather three integers (NCCL), every rank
    translates its piece into its byte ranges of the record's output.  Strong scaling: the record is
    the same at every N."""
    import numpy as np
    import torch
    import torch.distributed as dist
    import gotranseq_b200
    import workloads
    from gotranseq_b200 import split

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    gen, opts = workloads.WORKLOADS["chromosome"]
    data = gen()
    lo, hi = split.piece_ranges(data, world)[rank]
    piece_bytes = data[lo:hi]
    n = len(piece_bytes)
    O = gotranseq_b200.Options(Frame=opts["frame"], Alternative=opts.get("alternative", False))
    tr = gotranseq_b200.Translator(O, device=local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    d_in = torch.from_numpy(np.frombuffer(piece_bytes, dtype=np.uint8).copy()).cuda()
    cap = len(data) * 2 + 4096  # three frames of n/3 residues + newlines + headers
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    W = 4 + 192  # nucleotide count, header length, last two codes, first 192 codes
    mine = torch.zeros(W, dtype=torch.int64, device="cuda")
    every = torch.zeros(W * world, dtype=torch.int64, device="cuda")

    def step():
        info = tr.piece_scan_device(d_in.data_ptr(), n, rank == 0, stream)
        mine.copy_(torch.tensor([info["nucleotides"], info["header_len"], info["tail"][0], info["tail"][1]] +
                                list(info["head"])))
        if world > 1:
            dist.all_gather_into_tensor(every, mine)
            rows = every.view(world, W).tolist()
        else:
            rows = [mine.tolist()]
        infos = [dict(nucleotides=r[0], header_len=r[1], tail=(r[2], r[3]), head=bytes(r[4:])) for r in rows]
        piece = split.plan_pieces(infos)[rank]
        return tr.piece_translate_device(d_in.data_ptr(), n, piece, d_out.data_ptr(), cap, stream), infos

    warm = max(args.warmup, 3)
    for _ in range(warm):
        (total, segs), infos = step()
    launches0 = tr.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        step()
    ev1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    dev_s = ev0.elapsed_time(ev1) / 1e3
    launches = tr.stats()["kernel_launches"] - launches0
    nt_total = sum(i["nucleotides"] for i in infos)

    # every rank checks its own byte ranges against the oracle's output of the whole record
    ok = True
    if args.verify:
        import oracle
        want = oracle.translate(np.frombuffer(data, dtype=np.uint8), num_worker=1, **opts)
        ok = total == len(want)
        host = d_out[:total].cpu().numpy().tobytes() if ok else b""
        for off, length in segs:
            ok = ok and host[off:off + length] == want[off:off + length]
    seg_bytes = sum(length for _, length in segs)
    t = torch.tensor([dev_s, 0.0 if ok else 1.0], dtype=torch.float64, device="cuda")
    sb = torch.tensor([seg_bytes], dtype=torch.int64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(sb, op=dist.ReduceOp.SUM)
    dev_s = float(t[0])
    if rank == 0:
        peak, peak_src = measured_peak()
        line = {
            "metric": METRIC, "value": nt_total * args.steps / dev_s / 1e9, "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": warm, "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": "chromosome_split: one %d-byte record (%d nt, 60-column lines), flags -f R -a, cut "
                                   "into %d pieces at line starts, one per GPU" % (len(data), nt_total, world),
                       "l2": "piece + output per step exceed the 126 MB L2 at N <= 2; no explicit flush",
                       "sharding": "intra-record split; exchange = one all-gather of 196 small integers per rank per step"},
            "e2e": None, "gpu_launches": launches, "clocks": clocks,
            "segments_cover_output": int(sb[0]) == total,
            "verified": (float(t[1]) == 0.0) if args.verify else None,
            "bytes": {"in": len(data), "out": total, "nucleotides": nt_total, "records": 1},
            "peak_GBps": peak, "peak_source": peak_src,
        }
        print(json.dumps(line))
    tr.close()
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="readme189")
    ap.add_argument("--ref-sample-bytes", type=int, default=48_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end loop (default: min(steps, 10))")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end loop (profiling runs)")
    ap.add_argument("--chunk-mb", type=float, default=0, help="pipeline chunk of the host entry point (0 = library default)")
    ap.add_argument("--verify", action="store_true", help="chromosome_split: compare every rank's byte ranges with the oracle")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    rank, local_rank, world = env_rank()

    if args.workload == "chromosome_split":
        if args.impl == "reference":
            args.workload = "chromosome"
        else:
            run_split(args, rank, local_rank, world)
            return
    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import numpy as np
    import torch
    import torch.distributed as dist
    import gotranseq_b200
    import workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (gotranseq_b200 has no CPU fallback)")
    numa_cpus = bind_to_gpu_numa_node(local_rank) if world > 1 else None
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    gen, opts = workloads.WORKLOADS[args.workload]
    data = gen(seed=42 + rank) if args.workload == "readme189" else gen()
    nbytes = len(data)
    nt = count_nt(data, opts)
    warm = max(args.warmup, 3)

    O = gotranseq_b200.Options(Frame=opts.get("frame", "1"), Table=opts.get("table", 0), Clean=opts.get("clean", False),
                               Alternative=opts.get("alternative", False), Trim=opts.get("trim", False))
    tr = gotranseq_b200.Translator(O, device=local_rank, chunk_bytes=int(args.chunk_mb * (1 << 20)))
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident: input already in HBM, output stays in HBM ----
    d_in = torch.from_numpy(np.frombuffer(data, dtype=np.uint8).copy()).cuda()
    cap = nbytes * 4 + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
    out_n = tr.translate_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, stream)
    for _ in range(warm):
        tr.translate_device_async(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, stream)
    tr.device_result()
    tr.set_profiling(True)
    tr.kernel_times()
    launches0 = tr.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(args.steps):
        tr.translate_device_async(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, stream)
    ev1.record()
    barrier()
    clocks = sampler.stop()
    assert tr.device_result() == out_n
    dev_s = ev0.elapsed_time(ev1) / 1e3
    kms, kpasses = tr.kernel_times()
    tr.set_profiling(False)
    launches = tr.stats()["kernel_launches"] - launches0

    # ---- end to end: pinned host FASTA in, pinned host protein FASTA out, PCIe inside the clock ----
    L = gotranseq_b200._lib.lib()
    h_in = L.gts_host_alloc(nbytes + 64)
    ctypes.memmove(h_in, data, nbytes)
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    e2e_s, host_out = 0.0, None
    if not args.no_e2e:
        for _ in range(2):
            optr, on = tr.translate_host_ptr(h_in, nbytes)
        assert on == out_n
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            optr, on = tr.translate_host_ptr(h_in, nbytes)
        torch.cuda.synchronize()
        e2e_s = time.perf_counter() - t0
        host_out = bytes((ctypes.c_char * on).from_address(optr))

    # ---- max over ranks, sum of the work ----
    if world > 1:
        tt = torch.tensor([dev_s, e2e_s], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, e2e_s = float(tt[0]), float(tt[1])
        ws = torch.tensor([nt, nbytes, out_n], dtype=torch.float64, device="cuda")
        dist.all_reduce(ws, op=dist.ReduceOp.SUM)
        nt_all, in_all, out_all = float(ws[0]), float(ws[1]), float(ws[2])
    else:
        nt_all, in_all, out_all = float(nt), float(nbytes), float(out_n)

    if rank == 0:
        peak, peak_src = measured_peak()
        value = nt_all * args.steps / dev_s / 1e9
        e2e_v = nt_all * e2e_steps / e2e_s / 1e9 if e2e_s > 0 else None
        # algorithmic bytes of one k_lines launch: the packed symbol stream it reads (4 bits per
        # symbol: nucleotides + 2 pads per record) + the residue lines it writes (output - headers)
        st = tr.stats()
        nrec = st["records"]
        # header bytes written by k_headers: 6 x (H + 3) per record
        hdr_in = nbytes - nt - data.count(b"\n")  # header bytes incl '>'
        nframes = {"1": 1, "2": 1, "3": 1, "F": 3, "-1": 1, "-2": 1, "-3": 1, "R": 3, "6": 6}[O.Frame]
        hdr_out = nframes * (hdr_in + 3 * nrec)
        sym_bytes = (nt + 2 * nrec + 4) / 2
        trans_bytes = sym_bytes + (out_n - hdr_out)
        trans_ms = kms[2] / max(kpasses, 1)
        achieved = trans_bytes / (trans_ms * 1e-3) / 1e9 if trans_ms > 0 else 0.0
        path_bytes = nbytes + out_n
        path_gbs = path_bytes * args.steps / dev_s / 1e9 if world == 1 else (in_all + out_all) * args.steps / dev_s / 1e9 / world
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, nbytes),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": out_n + 64,
                    "ms_per_step": e2e_s / e2e_steps * 1e3 if e2e_s > 0 else None, "steps": e2e_steps,
                    "api": "gts_translate_buffer (pinned host buffers)",
                    "cpus_per_rank": numa_cpus},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_lines", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(args.workload), "peak_source": peak_src,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_lines launch, "
                                           "ncu --set full capture profiles/r1h_ncu_full_summary.txt",
                         "algorithmic_bytes_per_launch": trans_bytes, "ms_per_launch": trans_ms},
            "path": {"bytes_per_step_per_gpu": path_bytes, "achieved_GBps_per_gpu": path_gbs, "frac_of_peak": path_gbs / peak,
                     "frac_of_nominal_8TBps": path_gbs / 8000.0,
                     "note": "(FASTA in + protein FASTA out) / device time of the whole path"},
            "kernels_ms_per_step": {"reset+k_parse": kms[0] / max(kpasses, 1),
                                    "k_tile_scan+k_records+k_size": kms[1] / max(kpasses, 1),
                                    "k_lines": kms[2] / max(kpasses, 1), "k_headers": kms[3] / max(kpasses, 1)},
            "bytes": {"in": nbytes, "out": out_n, "nucleotides": nt, "records": nrec},
        }
        if world == 1 and not args.no_cpu_baseline:
            import oracle
            cores = os.cpu_count() or 1
            arr = np.frombuffer(data, dtype=np.uint8)
            t0 = time.perf_counter()
            want = oracle.translate(arr, num_worker=1, **opts)
            t1 = time.perf_counter()
            oracle.translate(arr, num_worker=cores, **opts)
            t2 = time.perf_counter()
            line["cpu_baseline"] = {
                "value": nt / (t2 - t1) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                "sample": "the whole %d-byte batch once, restated reference (oracle/transeq_oracle.c, 1 reader + %d "
                          "workers); single worker: %.4f Gbp/s" % (nbytes, cores, nt / (t1 - t0) / 1e9)}
            line["verified"] = bool(want == host_out) if host_out is not None else None
            if not line["verified"]:
                print("bench: GPU output differs from the oracle", file=sys.stderr)
        print(json.dumps(line))
    L.gts_host_free(h_in)
    tr.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
