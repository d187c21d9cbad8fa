"""Synthetic FASTA of the benchmark shapes (SURVEY.md 8d / BASELINE.json configs).  Seeded numpy
generation, identical on every box; used by bench.py and by the full-size GPU property tests."""
import numpy as np

_ACGT = np.frombuffer(b"ACGT", dtype=np.uint8)


def _bases(rng, n, n_rate):
    b = _ACGT[rng.integers(0, 4, size=n, dtype=np.uint8)]
    if n_rate:
        b[rng.integers(0, n_rate, size=n, dtype=np.uint32) == 0] = ord("N")
    return b


def _wrap(seq, width):
    n = len(seq)
    if n == 0:
        return b""
    return b"\n".join([seq[i:i + width] for i in range(0, n, width)]) + b"\n"


def readme189(total_bytes=189_000_000, seed=42):
    """configs[1]: the README benchmark shape -- multi-record FASTA, n ~ U[300,3000],
    header '>seq%07d len=%d', 60-column lines, uniform ACGT with 0.1 % N."""
    rng = np.random.default_rng(seed)
    parts, size, r = [], 0, 0
    while size < total_bytes:
        ns = rng.integers(300, 3001, size=4096)
        b = _bases(rng, int(ns.sum()), 1000).tobytes()
        off = 0
        for n in ns:
            n = int(n)
            h = b">seq%07d len=%d\n" % (r, n)
            body = _wrap(b[off:off + n], 60)
            off += n
            parts.append(h)
            parts.append(body)
            size += len(h) + len(body)
            r += 1
            if size >= total_bytes:
                break
    return b"".join(parts)


def contigs(n_records=40, len_lo=4_500_000, len_hi=5_500_000, seed=43):
    """configs[2]: bacterial-genome-shaped contigs (scaled by n_records), 60-column lines, 0.01 % N."""
    rng = np.random.default_rng(seed)
    parts = []
    for r in range(n_records):
        n = int(rng.integers(len_lo, len_hi + 1))
        b = _bases(rng, n, 10000)
        full = n // 60
        lines = np.full((full, 61), ord("\n"), dtype=np.uint8)
        lines[:, :60] = b[:full * 60].reshape(full, 60)
        parts.append(b">contig_%05d\n" % r)
        parts.append(lines.tobytes())
        if n % 60:
            parts.append(b[full * 60:].tobytes() + b"\n")
    return b"".join(parts)


def chromosome(n=250_000_000, seed=44):
    """configs[3]: one chromosome-shaped record, 60-column lines."""
    return contigs(1, n, n, seed).replace(b">contig_00000", b">chr1", 1)


def short_reads(n_reads=1_000_000, seed=45):
    """configs[4]: 150-bp reads on one line, 52-byte Illumina-style headers (204 B per record)."""
    rng = np.random.default_rng(seed)
    b = _bases(rng, n_reads * 150, 0).reshape(n_reads, 150)
    hdr = np.frombuffer(b">A00123:45:HXXXXXXXX:1:1101:00000:00000 1:N:0:ATCACG\n", dtype=np.uint8)
    H = len(hdr)  # 52 header bytes + newline
    rec = np.empty((n_reads, H + 151), dtype=np.uint8)
    xy = rng.integers(0, 100000, size=(n_reads, 2))
    rec[:, :H] = hdr
    for col, src in ((28, 0), (34, 1)):
        v = xy[:, src].copy()
        for d in range(4, -1, -1):
            rec[:, col + d] = ord("0") + (v % 10)
            v //= 10
    rec[:, H:H + 150] = b
    rec[:, H + 150] = ord("\n")
    return rec.tobytes()


WORKLOADS = {
    "readme189": (readme189, dict(frame="6", table=0)),
    "contigs": (contigs, dict(frame="6", table=11, clean=True, trim=True)),
    "chromosome": (chromosome, dict(frame="R", alternative=True)),
    "short_reads": (short_reads, dict(frame="6", table=2)),
}
