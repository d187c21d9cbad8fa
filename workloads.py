"""Synthetic FASTA of the benchmark shapes (SURVEY.md 8d / BASELINE.json configs) -- BENCH AND TEST
INFRASTRUCTURE.  Every record is a pure function of (shape, seed, record index) (workloads_gen.c:
splitmix64 over a counter), so any shard of a stream can be generated independently on any rank and
by several threads, and is identical on every box.

  readme189(total_bytes, seed)       configs[1]: n ~ U[300,3000], '>seq%07d len=%d', 60 columns, 0.1 % N
  contigs(n_records, ...)            configs[2]: ~5 Mbp contigs, 60 columns, 0.01 % N
  chromosome(n)                      configs[3]: one record, 60 columns
  short_reads(n_reads)               configs[4]: 150-bp reads on one line, 52-byte Illumina-style headers
  Stream(shape, seed, ...)           record-indexed view: sizes, byte offsets, shard ranges, generation
"""
import ctypes
import os
import subprocess
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "workloads_gen.c")
_LIB = os.path.join(_HERE, "libworkloads_gen.so")
_lib = None
KINDS = {"readme": 0, "contigs": 1, "short_reads": 2, "chromosome": 3}


def build(force=False):
    if force or not os.path.exists(_LIB) or os.path.getmtime(_LIB) < os.path.getmtime(_SRC):
        subprocess.run(["gcc", "-O2", "-fPIC", "-shared", "-o", _LIB, _SRC], check=True, capture_output=True)
    return _LIB


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB)
        u64, vp = ctypes.c_uint64, ctypes.c_void_p
        L.wg_record_size.restype = u64
        L.wg_record_size.argtypes = [ctypes.c_int, u64, u64, u64]
        L.wg_record_nt.restype = u64
        L.wg_record_nt.argtypes = [ctypes.c_int, u64, u64, u64]
        L.wg_total_nt.restype = u64
        L.wg_total_nt.argtypes = [ctypes.c_int, u64, u64, u64, u64]
        L.wg_record_sizes.restype = None
        L.wg_record_sizes.argtypes = [ctypes.c_int, u64, u64, u64, u64, vp]
        L.wg_generate.restype = u64
        L.wg_generate.argtypes = [ctypes.c_int, u64, u64, u64, u64, vp, u64]
        L.wg_generate_part.restype = u64
        L.wg_generate_part.argtypes = [ctypes.c_int, u64, u64, u64, u64, u64, vp]
        _lib = L
    return _lib


class Stream:
    """A stream of `count` records of one shape: record sizes, byte offsets, generation of any
    record range into a numpy uint8 array (or a caller-provided address)."""

    def __init__(self, shape, seed, count, scale_nt=0):
        self.kind, self.seed, self.count, self.scale_nt = KINDS[shape], int(seed), int(count), int(scale_nt)
        sizes = np.empty(self.count, dtype=np.uint64)
        if self.count:
            lib().wg_record_sizes(self.kind, self.seed, 0, self.count, self.scale_nt, sizes.ctypes.data)
        self.offsets = np.concatenate(([0], np.cumsum(sizes, dtype=np.uint64))).astype(np.int64)

    @property
    def nbytes(self):
        return int(self.offsets[-1])

    def nucleotides(self, lo=0, hi=None):
        hi = self.count if hi is None else hi
        return int(lib().wg_total_nt(self.kind, self.seed, lo, hi - lo, self.scale_nt))

    def shard_records(self, nshards):
        """Record ranges [(lo, hi), ...] of nshards contiguous shards of roughly equal bytes -- what
        gotranseq_b200.shard.shard_ranges gives on the bytes (cuts at the record boundary at or below
        k * total / nshards)."""
        total = self.nbytes
        cuts = [0]
        for k in range(1, nshards):
            pos = (total * k) // nshards
            r = int(np.searchsorted(self.offsets, pos, side="right")) - 1  # the record holding byte pos
            cuts.append(max(r, cuts[-1]))
        cuts.append(self.count)
        return [(cuts[k], cuts[k + 1]) for k in range(nshards)]

    def generate(self, lo=0, hi=None, out_addr=None, threads=None):
        """Records [lo, hi) as a numpy uint8 array (or written at out_addr; returns the byte count)."""
        hi = self.count if hi is None else hi
        nbytes = int(self.offsets[hi] - self.offsets[lo])
        arr = None
        if out_addr is None:
            arr = np.empty(nbytes, dtype=np.uint8)
            out_addr = arr.ctypes.data
        L = lib()
        threads = threads or min(16, os.cpu_count() or 1)
        jobs = []
        if hi - lo == 1 and nbytes > (8 << 20):
            # one giant record: split by nucleotide ranges (multiples of 120 keep blocks and lines whole)
            n = L.wg_record_nt(self.kind, self.seed, lo, self.scale_nt)
            w = 150 if self.kind == 2 else 60
            hdr = L.wg_record_size(self.kind, self.seed, lo, self.scale_nt) - n - (n + w - 1) // w
            step = max(120, (n // threads + 119) // 120 * 120)
            a = 0
            while a < n:
                b = min(n, a + step)
                off = (hdr if a else 0) + a + a // w
                jobs.append(("part", lo, a, b, off))
                a = b
        else:
            per = max(1, (hi - lo + threads - 1) // threads)
            for a in range(lo, hi, per):
                b = min(hi, a + per)
                jobs.append(("recs", a, b, 0, int(self.offsets[a] - self.offsets[lo])))

        def run(job):
            kind, a, b, c, off = job
            if kind == "recs":
                got = L.wg_generate(self.kind, self.seed, a, b - a, self.scale_nt, out_addr + off,
                                    int(self.offsets[b] - self.offsets[a]))
                assert got == int(self.offsets[b] - self.offsets[a])
            else:
                L.wg_generate_part(self.kind, self.seed, a, self.scale_nt, b, c, out_addr + off)

        if len(jobs) == 1:
            run(jobs[0])
        else:
            ts = [threading.Thread(target=run, args=(j,)) for j in jobs]
            for t in ts:
                t.start()
            for t in ts:
                t.join()
        return arr if arr is not None else nbytes


def _count_for_bytes(shape, seed, total_bytes, guess_avg):
    """Smallest record count whose stream reaches total_bytes."""
    count = max(1, int(total_bytes / guess_avg * 1.05) + 16)
    while True:
        s = Stream(shape, seed, count)
        if s.nbytes >= total_bytes:
            break
        count *= 2
    k = int(np.searchsorted(s.offsets, total_bytes, side="left"))
    return max(1, k)


def readme_stream(total_bytes=189_000_000, seed=42):
    return Stream("readme", seed, _count_for_bytes("readme", seed, total_bytes, 1700))


def readme189(total_bytes=189_000_000, seed=42):
    """configs[1]: the README benchmark shape -- multi-record FASTA, n ~ U[300,3000],
    header '>seq%07d len=%d', 60-column lines, uniform ACGT with 0.1 % N."""
    return readme_stream(total_bytes, seed).generate().tobytes()


def contigs(n_records=40, len_lo=4_500_000, len_hi=5_500_000, seed=43):
    """configs[2]: bacterial-genome-shaped contigs (scaled by n_records), 60-column lines, 0.01 % N."""
    if (len_lo, len_hi) == (4_500_000, 5_500_000):
        return Stream("contigs", seed, n_records).generate().tobytes()
    # other lengths (tests): records of the chromosome shape with per-record lengths from the seed
    rng = np.random.default_rng(seed)
    parts = []
    for r in range(n_records):
        n = int(rng.integers(len_lo, len_hi + 1))
        body = Stream("chromosome", seed * 1000 + r, 1, scale_nt=n).generate().tobytes()
        parts.append(b">contig_%05d\n" % r + body[body.index(b"\n") + 1:])
    return b"".join(parts)


def chromosome(n=250_000_000, seed=44):
    """configs[3]: one chromosome-shaped record, 60-column lines."""
    return Stream("chromosome", seed, 1, scale_nt=n).generate().tobytes()


def short_reads(n_reads=1_000_000, seed=45):
    """configs[4]: 150-bp reads on one line, 52-byte Illumina-style headers (204 B per record)."""
    return Stream("short_reads", seed, n_reads).generate().tobytes()


WORKLOADS = {
    "readme189": (readme189, dict(frame="6", table=0)),
    "contigs": (contigs, dict(frame="6", table=11, clean=True, trim=True)),
    "chromosome": (chromosome, dict(frame="R", alternative=True)),
    "short_reads": (short_reads, dict(frame="6", table=2)),
}
# shape of the record stream behind each workload: (Stream shape, seed, avg bytes per record)
STREAMS = {
    "readme189": ("readme", 42, 1700),
    "contigs": ("contigs", 43, 5_083_000),
    "short_reads": ("short_reads", 45, 204),
}
