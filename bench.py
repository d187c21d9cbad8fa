#!/usr/bin/env python
"""bench.py -- the translation hot path on synthetic FASTA of the BASELINE shapes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload readme189] [--impl ours|reference]
                  [--stream-gb G] [--one-context]

A step is one pass of the whole hot path (parse -> encode -> translate all six frames -> format)
over one batch of FASTA (BASELINE.json configs[1]: 189 MB, multi-record, -f 6 -t 0).
  value      device-resident Gbp/s (input and output in HBM, CUDA events on the launch stream)
  e2e        the same metric through gts_translate_buffer with pinned HOST buffers: the library's chunk
             engine (record-aligned chunks, H2D / kernels / D2H of different chunks overlapped), PCIe
             copies inside the clock
  e2e_stream the same through gts_acquire_input / gts_submit / gts_wait_output (the calls the Go shim
             makes): the reader copies the FASTA into the library's pinned buffers (an io.Reader)
  raw_copy   cudaMemcpyAsync of the same H2D + D2H bytes in the same chunks, nothing else: what the
             box's PCIe / host memory path allows (the roofline of e2e)
  roofline   the dominant kernel against the measured HBM copy rate
  cpu_baseline / --impl reference  the restated reference (oracle/, all host threads)
Under torchrun (N > 1) ONE stream of N x 189 MB is cut into N shards at record boundaries
(gotranseq_b200.shard / workloads.Stream.shard_records); rank r translates shard r on GPU r; the
shards' outputs are the consecutive byte ranges of the stream's output (offsets from an all-gather
of the sizes).  No collective on the data path; the time is the max over ranks.  Rank 0 then also
drives all N GPUs from ONE gts_ctx (the product API: gts_options.num_gpus) over the whole stream.
--stream-gb G streams G GB (per job, all ranks together) through the streaming interface.
"""
import argparse
import ctypes
import hashlib
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "6-frame Gbp/s device-resident"
UNIT = "Gbp/s"
# DRAM bytes of one launch of the dominant kernel on the default workload, from the committed ncu
# capture named in NCU_SOURCE (dram__bytes_read.sum + dram__bytes_write.sum); null for other workloads
NCU_TRAFFIC = {"readme189": 140930304 + 333649920}
NCU_SOURCE = "profiles/r2b_ncu_full_summary.txt"


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
            time.sleep(0.001)

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


def flags_text(opts):
    return "-f %s -t %d%s%s%s" % (opts.get("frame", "1"), opts.get("table", 0), " -c" if opts.get("clean") else "",
                                   " -a" if opts.get("alternative") else "", " -T" if opts.get("trim") else "")


# ---------------------------------------------------------------------------------------------
# the batch of this rank: a shard of ONE stream
# ---------------------------------------------------------------------------------------------
class Batch:
    """Shard `rank` of `world` of the workload's record stream, generated straight into pinned host
    memory (gts_host_alloc)."""

    def __init__(self, workload, rank, world, L, per_gpu=None):
        """L: the C ABI library (pinned memory through gts_host_alloc), or None: a plain numpy buffer
        (the reference arm loads nothing of the product)."""
        import workloads
        self.workload, self.L = workload, L
        _, self.opts = workloads.WORKLOADS[workload]
        if workload == "chromosome":
            # one record cannot be sharded by records: every rank gets a chromosome of its own
            self.stream = workloads.Stream("chromosome", 44 + rank, 1, scale_nt=per_gpu or 250_000_000)
            self.lo, self.hi = 0, 1
            self.stream_desc = "one %d-nt record per GPU" % self.stream.nucleotides()
        else:
            shape, seed, avg = workloads.STREAMS[workload]
            if workload == "readme189":
                count = workloads._count_for_bytes(shape, seed, (per_gpu or 189_000_000) * world, avg)
            elif workload == "contigs":
                count = (per_gpu or 40) * world
            else:
                count = (per_gpu or 1_000_000) * world
            self.stream = workloads.Stream(shape, seed, count)
            self.lo, self.hi = self.stream.shard_records(world)[rank]
            self.stream_desc = "shard %d of %d of one stream of %d records, %d bytes" % (rank, world, count, self.stream.nbytes)
        self.nbytes = int(self.stream.offsets[self.hi] - self.stream.offsets[self.lo])
        self.nt = self.stream.nucleotides(self.lo, self.hi)
        self.records = self.hi - self.lo
        if L is None:
            import numpy as np
            self._np = np.empty(self.nbytes + 64, dtype=np.uint8)
            self.h_in = self._np.ctypes.data
        else:
            self.h_in = L.gts_host_alloc(self.nbytes + 64)
            if not self.h_in:
                raise MemoryError("gts_host_alloc")
        self.stream.generate(self.lo, self.hi, out_addr=self.h_in)

    def array(self):
        import numpy as np
        return np.ctypeslib.as_array((ctypes.c_uint8 * self.nbytes).from_address(self.h_in))

    def free(self):
        if self.h_in and self.L is not None:
            self.L.gts_host_free(self.h_in)
        self.h_in = None


def record_cut(arr, limit):
    """Largest prefix of the FASTA bytes (numpy uint8) of at most `limit` bytes that ends at a record boundary."""
    n = len(arr)
    if n <= limit:
        return n
    b = bytes(arr[max(0, limit - (64 << 20)):limit])
    i = b.rfind(b"\n>")
    return limit - len(b) + i + 1 if i >= 0 else n


def workload_config(args, batch, world):
    return {"workload": "%s: %d-byte synthetic FASTA per GPU (%s; %s), flags %s" % (
                args.workload, batch.nbytes, batch.stream_desc,
                "workloads.py " + args.workload, flags_text(batch.opts)),
            "l2": "input + output per step exceed the 126 MB L2 several times over; no explicit flush",
            "sharding": "one stream cut at record boundaries into one shard per GPU, no collective" if world > 1 else "one GPU"}


def run_reference(args, rank, world):
    """The restated reference (oracle/, see its header) on the host cores: bounded sample per step."""
    if rank != 0:
        return
    import oracle
    batch = Batch(args.workload, 0, 1, None)
    arr = batch.array()
    cores = os.cpu_count() or 1
    cut = record_cut(arr, args.ref_sample_bytes)
    sample = arr[:cut]
    # nucleotides of the sample: records are whole, so count through the stream
    import numpy as np
    k = int(np.searchsorted(batch.stream.offsets, batch.stream.offsets[batch.lo] + cut, side="right")) - 1
    nt = batch.stream.nucleotides(batch.lo, k)
    for _ in range(args.warmup):
        oracle.translate(sample, num_worker=cores, **batch.opts)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        oracle.translate(sample, num_worker=cores, **batch.opts)
    dt = (time.perf_counter() - t0) / args.steps
    v = nt / dt / 1e9
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": workload_config(args, batch, 1),
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": "first %d bytes (%d nt) of the workload per step, restated reference "
                                   "(oracle/transeq_oracle.c: 1 reader + %d workers), not the Go binary "
                                   "(no Go toolchain in this image)" % (cut, nt, cores)},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    batch.free()


# ---------------------------------------------------------------------------------------------
# end-to-end legs
# ---------------------------------------------------------------------------------------------
def e2e_buffer(tr, h_in, nbytes, steps, sync):
    """gts_translate_buffer from pinned host memory; returns (seconds, out pointer, out size)."""
    for _ in range(2):
        optr, on = tr.translate_host_ptr(h_in, nbytes)
    sync()
    t0 = time.perf_counter()
    for _ in range(steps):
        optr, on = tr.translate_host_ptr(h_in, nbytes)
    dt = time.perf_counter() - t0
    return dt, optr, on


_POOL = None


def _copy(dst, src, n, threads):
    """memcpy by a few threads (ctypes releases the GIL): the reader side of the streaming legs -- an
    io.Reader over a file or a buffer is free to fill the destination with several threads."""
    global _POOL
    if threads <= 1 or n < (4 << 20):
        ctypes.memmove(dst, src, n)
        return
    if _POOL is None:
        from concurrent.futures import ThreadPoolExecutor
        _POOL = ThreadPoolExecutor(max_workers=8)
    per = (n + threads - 1) // threads
    futs = [_POOL.submit(ctypes.memmove, dst + o, src + o, min(per, n - o)) for o in range(0, n, per)]
    for f in futs:
        f.result()


class _StreamApi(ctypes.Structure):
    _fields_ = [(n, ctypes.c_void_p) for n in ("acquire_input", "submit", "wait_output", "release_output", "cancel")]


class _StreamResult(ctypes.Structure):
    _fields_ = [("out_bytes", ctypes.c_uint64), ("chunks", ctypes.c_uint64), ("rc", ctypes.c_int)]


_HARNESS = {}
HARNESS_SRC = os.path.join(ROOT, "tools", "stream_harness.cpp")
HARNESS_LIB = os.path.join(ROOT, "tools", "libstream_harness.so")


def stream_harness(force=False):
    """tools/stream_harness.cpp: the shim's loop in C++ (reader thread + memcpy helpers, writer thread), so
    that the streaming legs measure the library and the copies, not the interpreter's thread hand-overs.
    Compiled with g++ on first use; returns the loaded library, or None when it cannot be built."""
    if "lib" in _HARNESS and not force:
        return _HARNESS["lib"]
    lib = None
    try:
        stale = not os.path.exists(HARNESS_LIB) or os.path.getmtime(HARNESS_LIB) < os.path.getmtime(HARNESS_SRC)
        if stale or force:
            tmp = "%s.%d.tmp" % (HARNESS_LIB, os.getpid())  # (several ranks may build at once)
            subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-shared", "-fPIC", "-Wall", "-Wextra", "-o", tmp,
                            HARNESS_SRC], check=True, capture_output=True)
            os.replace(tmp, HARNESS_LIB)
        lib = ctypes.CDLL(HARNESS_LIB)
        lib.stream_harness_run.restype = ctypes.c_int
        lib.stream_harness_run.argtypes = [ctypes.POINTER(_StreamApi), ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                           ctypes.c_int, ctypes.c_int, ctypes.c_size_t, ctypes.POINTER(_StreamResult)]
    except Exception as e:  # no compiler on this box: the Python loop below does the same, with more jitter
        print("bench: stream harness not available (%s), using the Python loop" % e, file=sys.stderr)
        lib = None
    _HARNESS["lib"] = lib
    return lib


def stream_once(tr, h_in, nbytes, repeats=1, hasher=None, read_size=0, reader_threads=1, python_loop=False):
    """The shim's loop: reader = this thread (memcpy from h_in into the acquired pinned buffer, with
    reader_threads helpers), writer thread = gts_wait_output / gts_release_output.  Returns the counters.
    Driven from C++ (tools/stream_harness.cpp) unless the output is to be hashed or python_loop is set."""
    L, h = tr._L, tr._h
    res = {"out": 0, "chunks": 0, "err": None, "driver": "python"}
    H = None if (hasher is not None or python_loop) else stream_harness()
    if H is not None:
        api = _StreamApi(*[ctypes.cast(getattr(L, "gts_" + n), ctypes.c_void_p).value for n, _ in _StreamApi._fields_])
        out = _StreamResult()
        rc = H.stream_harness_run(ctypes.byref(api), h, ctypes.c_void_p(h_in), nbytes, repeats, reader_threads, read_size,
                                  ctypes.byref(out))
        if rc != 0:
            msg = L.gts_last_error(h).decode()
            L.gts_stream_reset(h)
            raise RuntimeError("stream harness: rc %d: %s" % (rc, msg))
        return {"out": out.out_bytes, "chunks": out.chunks, "err": None, "driver": "c++"}

    def writer():
        p, m = ctypes.c_void_p(), ctypes.c_size_t()
        while True:
            rc = L.gts_wait_output(h, ctypes.byref(p), ctypes.byref(m))
            if rc != 0:
                res["err"] = L.gts_last_error(h).decode()
                return
            if m.value == 0:
                return
            if hasher is not None:
                hasher.update((ctypes.c_char * m.value).from_address(p.value))
            res["out"] += m.value
            res["chunks"] += 1
            L.gts_release_output(h)

    wt = threading.Thread(target=writer)
    wt.start()
    buf, cap = ctypes.c_void_p(), ctypes.c_size_t()
    try:
        for r in range(repeats):
            pos = 0
            while pos < nbytes:
                rc = L.gts_acquire_input(h, ctypes.byref(buf), ctypes.byref(cap))
                if rc != 0:
                    raise RuntimeError(L.gts_last_error(h).decode())
                k = min(cap.value, nbytes - pos)
                if read_size:
                    k = min(k, read_size)
                _copy(buf.value, h_in + pos, k, reader_threads)
                pos += k
                rc = L.gts_submit(h, k, 0)
                if rc != 0:
                    raise RuntimeError(L.gts_last_error(h).decode())
        rc = L.gts_submit(h, 0, 1)
        if rc != 0:
            raise RuntimeError(L.gts_last_error(h).decode())
    finally:
        if res["err"] is None and wt.is_alive() and sys.exc_info()[0] is not None:
            L.gts_cancel(h)
        wt.join()
    if res["err"]:
        raise RuntimeError(res["err"])
    return res


def raw_copy(torch, nbytes_in, nbytes_out, steps, chunk=16 << 20, devices=None):
    """cudaMemcpyAsync of nbytes_in H2D and nbytes_out D2H per device in `chunk`-sized pieces, the two
    directions on separate streams: the copy roofline of the end-to-end path on this box."""
    devices = devices or [torch.cuda.current_device()]
    st = []
    for d in devices:
        with torch.cuda.device(d):
            h_a = torch.empty(nbytes_in, dtype=torch.uint8).pin_memory()
            h_b = torch.empty(nbytes_out, dtype=torch.uint8).pin_memory()
            d_a = torch.empty(nbytes_in, dtype=torch.uint8, device=torch.device("cuda", d))
            d_b = torch.empty(nbytes_out, dtype=torch.uint8, device=torch.device("cuda", d))
            st.append((d, h_a, h_b, d_a, d_b, torch.cuda.Stream(), torch.cuda.Stream()))

    def once():
        for d, h_a, h_b, d_a, d_b, s1, s2 in st:
            with torch.cuda.device(d):
                with torch.cuda.stream(s1):
                    for o in range(0, nbytes_in, chunk):
                        d_a[o:o + chunk].copy_(h_a[o:o + chunk], non_blocking=True)
                with torch.cuda.stream(s2):
                    for o in range(0, nbytes_out, 2 * chunk):
                        h_b[o:o + 2 * chunk].copy_(d_b[o:o + 2 * chunk], non_blocking=True)
        for d, *_ in st:
            torch.cuda.synchronize(d)

    once()
    t0 = time.perf_counter()
    for _ in range(steps):
        once()
    return (time.perf_counter() - t0) / steps


# ---------------------------------------------------------------------------------------------
# config 4 over N GPUs: one record cut into N pieces
# ---------------------------------------------------------------------------------------------
def run_split(args, rank, local_rank, world):
    """--workload chromosome_split (BASELINE config 4 at N GPUs): ONE record cut into N pieces, one per
    rank.  A step = every rank scans its piece, the ranks all-gather ~200 small integers (NCCL), every
    rank translates its piece into its byte ranges of the record's output.  Strong scaling: the record
    is the same at every N."""
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
    d_out = torch.empty(cap, dtype=torch.uint8, device=torch.device("cuda", local_rank))
    from gotranseq_b200 import _lib
    IB = _lib.GTS_PIECE_INFO_BYTES
    d_info = torch.zeros(IB, dtype=torch.uint8, device=torch.device("cuda", local_rank))
    d_all = torch.zeros(IB * world, dtype=torch.uint8, device=torch.device("cuda", local_rank))

    def step():
        # scan -> all-gather of the 256-byte summaries (NCCL, same stream) -> plan on the device +
        # translate on the scan's slots -> one synchronisation
        tr.piece_scan_async(d_in.data_ptr(), n, d_info.data_ptr(), stream)
        if world > 1:
            dist.all_gather_into_tensor(d_all, d_info)
        else:
            d_all.copy_(d_info)
        tr.piece_translate_async(d_in.data_ptr(), n, rank, world, d_all.data_ptr(), d_out.data_ptr(), cap, stream)
        total, segs, piece = tr.piece_result()
        return (total, segs), piece

    warm = max(args.warmup, 3)
    for _ in range(warm):
        (total, segs), piece = step()
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
    nt_total = piece["nt_total"]

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
    t = torch.tensor([dev_s, 0.0 if ok else 1.0], dtype=torch.float64, device=torch.device("cuda", local_rank))
    sb = torch.tensor([seg_bytes], dtype=torch.int64, device=torch.device("cuda", local_rank))
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
                       "sharding": "intra-record split; exchange = one NCCL all-gather of 256 bytes per rank per step, the piece is planned on "
                                   "the device (no host round trip between scan and translate)"},
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


# ---------------------------------------------------------------------------------------------
def main():
    # The streaming legs run a reader (this thread + memcpy helpers) and a writer thread in one Python
    # process: with the default 5 ms switch interval a thread that needs the GIL back after a ctypes
    # call can wait 5 ms for it, several times per step -- more than the step itself.
    sys.setswitchinterval(1e-4)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="readme189")
    ap.add_argument("--per-gpu", type=int, default=0, help="size of a GPU's batch: bytes (readme189), records (contigs, short_reads), nucleotides (chromosome)")
    ap.add_argument("--ref-sample-bytes", type=int, default=48_000_000)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-steps", type=int, default=0, help="steps of the end-to-end loops (default: min(steps, 10))")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end loops (profiling runs)")
    ap.add_argument("--chunk-mb", type=float, default=0, help="chunk of the host engine (0 = library default, 16 MB)")
    ap.add_argument("--verify", action="store_true", help="chromosome_split: compare every rank's byte ranges with the oracle")
    ap.add_argument("--stream-gb", type=float, default=0, help="also stream this many GB (all ranks together) through the streaming interface")
    ap.add_argument("--one-context", action="store_true", help="plain python, --gpus N: ONE gts_ctx drives N GPUs (no torchrun)")
    ap.add_argument("--reader-threads", type=int, default=4, help="threads of the streaming legs' reader (memcpy into the pinned buffer)")
    ap.add_argument("--no-verify", action="store_true", help="skip the comparison with the oracle")
    ap.add_argument("--warnings", action="store_true", help="install a warning callback (the CLI / shim configuration)")
    ap.add_argument("--python-stream-loop", action="store_true", help="drive the streaming legs from Python threads instead of tools/stream_harness.cpp")
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
    from gotranseq_b200 import _lib

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

    L = _lib.lib()
    one_ctx_n = args.gpus if (args.one_context and world == 1) else 0
    batch = Batch(args.workload, rank, world, L, per_gpu=(args.per_gpu or None) if not one_ctx_n else
                  (args.per_gpu or {"readme189": 189_000_000, "contigs": 40, "short_reads": 1_000_000}.get(args.workload, 0)) * one_ctx_n)
    opts = batch.opts
    nbytes, nt = batch.nbytes, batch.nt
    warm = max(args.warmup, 3)

    O = gotranseq_b200.Options(Frame=opts.get("frame", "1"), Table=opts.get("table", 0), Clean=opts.get("clean", False),
                               Alternative=opts.get("alternative", False), Trim=opts.get("trim", False))
    sink = (lambda line: None) if args.warnings else None
    chunk_bytes = int(args.chunk_mb * (1 << 20))
    tr = gotranseq_b200.Translator(O, device=local_rank, chunk_bytes=chunk_bytes, warnings=sink,
                                   devices=list(range(one_ctx_n)) if one_ctx_n else None)
    stream = torch.cuda.current_stream().cuda_stream

    # ---- device-resident: input already in HBM, output stays in HBM ----
    d_in = torch.from_numpy(batch.array()).cuda()
    cap = nbytes * 4 + 4096
    d_out = torch.empty(cap, dtype=torch.uint8, device=torch.device("cuda", local_rank))
    out_n = tr.translate_device(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, stream)
    sampler = ClockSampler(local_rank)
    sampler.start()  # (sampled across the warm-up as well: the timed region is only ~10 ms long)
    for _ in range(warm):
        tr.translate_device_async(d_in.data_ptr(), nbytes, d_out.data_ptr(), cap, stream)
    tr.device_result()
    tr.set_profiling(True)
    tr.kernel_times()
    launches0 = tr.stats()["kernel_launches"]
    barrier()
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
    dev_out_sha = None
    if args.workload == "readme189" and world == 1 and not one_ctx_n:
        dev_out_sha = hashlib.sha256(d_out[:out_n].cpu().numpy().tobytes()).hexdigest()
    del d_in, d_out
    torch.cuda.empty_cache()

    # ---- end to end: pinned host FASTA in, pinned host protein FASTA out, PCIe inside the clock ----
    e2e_steps = args.e2e_steps or min(args.steps, 10)
    e2e_s = stream_s = stream1_s = raw_s = 0.0
    host_out = None
    stream_res = None
    if not args.no_e2e:
        e2e_s, optr, on = e2e_buffer(tr, batch.h_in, nbytes, e2e_steps, barrier)
        assert on == out_n, (on, out_n)
        host_out = (ctypes.c_char * on).from_address(optr)
        barrier()
        # the streaming interface (reader memcpy into the library's pinned buffers, writer thread)
        pl = args.python_stream_loop
        for _ in range(2):  # (the pinned buffer pools settle on their sizes)
            stream_once(tr, batch.h_in, nbytes, python_loop=pl)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            stream_res = stream_once(tr, batch.h_in, nbytes, python_loop=pl)
        stream1_s = time.perf_counter() - t0
        assert stream_res["out"] == out_n, (stream_res, out_n)
        barrier()
        for _ in range(2):
            stream_once(tr, batch.h_in, nbytes, reader_threads=args.reader_threads, python_loop=pl)
        barrier()
        t0 = time.perf_counter()
        for _ in range(e2e_steps):
            stream_res = stream_once(tr, batch.h_in, nbytes, reader_threads=args.reader_threads, python_loop=pl)
        stream_s = time.perf_counter() - t0
        assert stream_res["out"] == out_n, (stream_res, out_n)
        barrier()
        raw_s = raw_copy(torch, nbytes // max(one_ctx_n, 1), out_n // max(one_ctx_n, 1), e2e_steps,
                         chunk=chunk_bytes or (16 << 20), devices=list(range(one_ctx_n)) if one_ctx_n else None) * e2e_steps
        barrier()

    # ---- verification ----
    verified = None
    cpu_base = None
    if host_out is not None and not args.no_verify:
        import oracle
        arr = batch.array()
        cores = os.cpu_count() or 1
        if world == 1 and not one_ctx_n and nbytes <= 300_000_000 and not args.no_cpu_baseline:
            t0 = time.perf_counter()
            want = oracle.translate(arr, num_worker=1, **opts)
            t1 = time.perf_counter()
            want_all = oracle.translate(arr, num_worker=cores, **opts)
            t2 = time.perf_counter()
            verified = bytes(host_out) == want and len(want_all) == len(want)
            if dev_out_sha is not None:
                verified = verified and dev_out_sha == hashlib.sha256(want).hexdigest()
            cpu_base = {"value": nt / (t2 - t1) / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
                        "sample": "the whole %d-byte batch once, restated reference (oracle/transeq_oracle.c, 1 reader + %d "
                                  "workers; not the Go binary: no Go toolchain in this image); single worker: %.4f Gbp/s"
                                  % (nbytes, cores, nt / (t1 - t0) / 1e9)}
        else:
            # a record-aligned prefix of this rank's shard, single worker (the reference's -n 1 order)
            cut = record_cut(arr, 8_000_000)
            want = oracle.translate(arr[:cut], num_worker=1, **opts)
            # (every chunk is a whole number of records, so the prefix's output is a prefix of the output;
            # a trailing record-0 rule cannot interfere: the shard goes on behind the prefix)
            verified = bytes(host_out[:len(want)]) == want if cut < nbytes else bytes(host_out) == want

    # ---- max over ranks, sum of the work ----
    if world > 1:
        tt = torch.tensor([dev_s, e2e_s, stream_s, raw_s, 0.0 if verified in (True, None) else 1.0, stream1_s], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        dev_s, e2e_s, stream_s, raw_s = (float(x) for x in tt[:4])
        stream1_s = float(tt[5])
        verified = (float(tt[4]) == 0.0) if verified is not None else None
        ws = torch.tensor([nt, nbytes, out_n], dtype=torch.float64, device=torch.device("cuda", local_rank))
        dist.all_reduce(ws, op=dist.ReduceOp.SUM)
        nt_all, in_all, out_all = float(ws[0]), float(ws[1]), float(ws[2])
        sizes = [torch.zeros(1, dtype=torch.int64, device=torch.device("cuda", local_rank)) for _ in range(world)]
        dist.all_gather(sizes, torch.tensor([out_n], dtype=torch.int64, device=torch.device("cuda", local_rank)))
        shard_out_offsets = np.concatenate(([0], np.cumsum([int(s) for s in sizes]))).tolist()
    else:
        nt_all, in_all, out_all = float(nt), float(nbytes), float(out_n)
        shard_out_offsets = [0, out_n]

    # ---- rank 0 drives every GPU from ONE context over the whole stream (the product's multi-GPU API) ----
    one_ctx = None
    if world > 1 and not args.no_e2e:
        barrier()
        if rank == 0:
            whole = Batch(args.workload, 0, 1, L, per_gpu=(args.per_gpu or {"readme189": 189_000_000, "contigs": 40,
                                                                           "short_reads": 1_000_000}.get(args.workload, 0)) * world)
            with gotranseq_b200.Translator(O, devices=list(range(world)), chunk_bytes=chunk_bytes) as trn:
                dt, optr, on = e2e_buffer(trn, whole.h_in, whole.nbytes, e2e_steps, lambda: None)
                ok = on == int(out_all)
                one_ctx = {"api": "gts_translate_buffer, gts_options.num_gpus = %d, one process" % world,
                           "value": whole.nt * e2e_steps / dt / 1e9, "unit": UNIT, "ms_per_step": dt / e2e_steps * 1e3,
                           "bytes_in": whole.nbytes, "bytes_out": on, "output_size_matches_the_sharded_run": ok}
            whole.free()
        barrier()

    # ---- optional: a long stream through the streaming interface ----
    long_stream = None
    if args.stream_gb > 0:
        per_rank = args.stream_gb * 1e9 / world
        repeats = max(1, int(round(per_rank / nbytes)))
        h1, h2 = hashlib.sha256(), hashlib.sha256()
        r1 = stream_once(tr, batch.h_in, nbytes, repeats=1, hasher=h1)            # untimed: one period, hashed
        r2 = stream_once(tr, batch.h_in, nbytes, repeats=2, hasher=h2)            # two periods back to back
        hh = hashlib.sha256()
        hh.update(host_out) if host_out is not None else None
        periods_ok = r2["out"] == 2 * r1["out"] and (host_out is None or h1.hexdigest() == hh.hexdigest())
        barrier()
        t0 = time.perf_counter()
        rr = stream_once(tr, batch.h_in, nbytes, repeats=repeats, reader_threads=args.reader_threads, python_loop=args.python_stream_loop)
        dt = time.perf_counter() - t0
        tt = torch.tensor([dt, 0.0 if (periods_ok and rr["out"] == repeats * out_n) else 1.0], dtype=torch.float64, device=torch.device("cuda", local_rank))
        if world > 1:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        long_stream = {"api": "gts_acquire_input / gts_submit / gts_wait_output, reader = memcpy into the pinned buffer (%d threads)" % args.reader_threads,
                       "gb_in": repeats * in_all / 1e9, "gb_out": repeats * out_all / 1e9, "seconds": float(tt[0]),
                       "value": repeats * nt_all / float(tt[0]) / 1e9, "unit": UNIT, "periods": repeats,
                       "period_bytes": nbytes, "chunks": rr["chunks"],
                       "verified": ("exact total size; one period's SHA-256 equals the one-shot output (whose record-aligned "
                                    "prefix is compared with the oracle: 'verified'); two periods = twice one")
                       if float(tt[1]) == 0.0 else False}

    if rank == 0:
        peak, peak_src = measured_peak()
        value = nt_all * args.steps / dev_s / 1e9
        e2e_v = nt_all * e2e_steps / e2e_s / 1e9 if e2e_s > 0 else None
        # algorithmic bytes of one k_lines launch: the packed symbol stream it reads (4 bits per
        # symbol: nucleotides + 2 pads per record) + the residue lines it writes (output - headers)
        nrec = batch.records
        arr = batch.array()
        hdr_in = nbytes - nt - int((arr == 10).sum())  # header bytes incl '>'
        nframes = {"1": 1, "2": 1, "3": 1, "F": 3, "-1": 1, "-2": 1, "-3": 1, "R": 3, "6": 6}[O.Frame]
        hdr_out = nframes * (hdr_in + 3 * nrec)
        sym_bytes = (nt + 2 * nrec + 4) / 2
        trans_bytes = sym_bytes + (out_n - hdr_out)
        trans_ms = kms[2] / max(kpasses, 1)
        achieved = trans_bytes / (trans_ms * 1e-3) / 1e9 if trans_ms > 0 else 0.0
        path_bytes = nbytes + out_n
        path_gbs = (in_all + out_all) * args.steps / dev_s / 1e9 / world
        ndev = max(one_ctx_n, 1)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world if not one_ctx_n else one_ctx_n,
            **({"note": "--one-context: value is device-resident on ONE GPU over the whole batch; e2e / e2e_stream are ONE "
                        "gts_ctx dealing the batch's chunks to %d GPUs" % one_ctx_n} if one_ctx_n else {}),
            "steps": args.steps, "warmup": warm,
            "ms_per_step": dev_s / args.steps * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": workload_config(args, batch, world),
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": out_n,
                    "ms_per_step": e2e_s / e2e_steps * 1e3 if e2e_s > 0 else None, "steps": e2e_steps,
                    "api": "gts_translate_buffer (pinned host buffers; the chunk engine of the streaming interface, chunks "
                           "read in place)" + (", ONE context over %d GPUs" % one_ctx_n if one_ctx_n else ""),
                    "cpus_per_rank": numa_cpus},
            "e2e_stream": {"value": nt_all * e2e_steps / stream_s / 1e9 if stream_s > 0 else None, "unit": UNIT,
                           "ms_per_step": stream_s / e2e_steps * 1e3 if stream_s > 0 else None,
                           "api": "gts_acquire_input / gts_submit / gts_wait_output / gts_release_output (the Go shim's calls); "
                                  "the reader memcpy's the FASTA into the acquired pinned buffer with %d threads" % args.reader_threads,
                           "driver": ("tools/stream_harness.cpp (reader thread + memcpy helpers, writer thread, no interpreter in the loop)"
                                      if stream_res and stream_res.get("driver") == "c++" else "Python threads (ctypes)"),
                           "single_thread_reader": {"value": nt_all * e2e_steps / stream1_s / 1e9 if stream1_s > 0 else None,
                                                    "ms_per_step": stream1_s / e2e_steps * 1e3 if stream1_s > 0 else None},
                           "chunks_per_step": stream_res["chunks"] if stream_res else None},
            "raw_copy": {"ms_per_step": raw_s / e2e_steps * 1e3 if raw_s > 0 else None,
                         "value": nt_all * e2e_steps / raw_s / 1e9 if raw_s > 0 else None, "unit": UNIT,
                         "GBps_per_gpu": (nbytes + out_n) / ndev * e2e_steps / raw_s / 1e9 if raw_s > 0 else None,
                         "note": "cudaMemcpyAsync of the same H2D and D2H bytes per GPU (pinned, both directions concurrently, "
                                 "all ranks at once): the ceiling of e2e on this box; e2e / raw_copy = %.2f"
                                 % ((raw_s / e2e_s) if e2e_s > 0 and raw_s > 0 else 0.0)},
            "gpu_launches": launches,
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": "k_lines", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": NCU_TRAFFIC.get(args.workload), "peak_source": peak_src,
                         "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one k_lines launch, "
                                           "ncu --set full capture " + NCU_SOURCE,
                         "algorithmic_bytes_per_launch": trans_bytes, "ms_per_launch": trans_ms},
            "path": {"bytes_per_step_per_gpu": path_bytes, "achieved_GBps_per_gpu": path_gbs, "frac_of_peak": path_gbs / peak,
                     "frac_of_nominal_8TBps": path_gbs / 8000.0,
                     "note": "(FASTA in + protein FASTA out) / device time of the whole path"},
            "kernels_ms_per_step": {"reset+k_parse": kms[0] / max(kpasses, 1),
                                    "k_tile_scan+k_records+k_size": kms[1] / max(kpasses, 1),
                                    "k_lines": kms[2] / max(kpasses, 1),
                                    "k_headers+k_short (side streams: what is left of them after k_lines)": kms[3] / max(kpasses, 1)},
            "bytes": {"in": nbytes, "out": out_n, "nucleotides": nt, "records": nrec},
            "shard_output_offsets": shard_out_offsets,
            "verified": verified,
        }
        if one_ctx is not None:
            line["one_context_all_gpus"] = one_ctx
        if long_stream is not None:
            line["stream"] = long_stream
        if cpu_base is not None:
            line["cpu_baseline"] = cpu_base
        if verified is False:
            print("bench: GPU output differs from the oracle", file=sys.stderr)
        print(json.dumps(line))
    tr.close()
    batch.free()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
