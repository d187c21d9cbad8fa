"""Record sharding for multi-GPU runs: the FASTA byte stream is cut into contiguous
ranges at record boundaries -- a '>' that starts a line (transeq.go:198) -- so that every shard is a
complete FASTA of its own and the shards' outputs, concatenated in shard order, are byte-identical
to the single-device output.  Shards are independent: no collective is needed on the data path.
"""


def record_cut(buf, pos):
    """Largest record boundary <= pos (0 if there is none): the index of a '>' preceded by '\\n'."""
    if pos <= 0:
        return 0
    if pos >= len(buf):
        return len(buf)
    i = buf.rfind(b"\n>", 0, pos + 1)
    return i + 1 if i >= 0 else 0


def shard_ranges(buf, nshards):
    """nshards contiguous (start, end) byte ranges of buf, of roughly equal size, cut at record
    boundaries.  Empty ranges are possible (more shards than records); shard 0 always starts at 0,
    so data before the first header line (the reference's slot-0 record) stays with shard 0."""
    n = len(buf)
    cuts = [0]
    for k in range(1, nshards):
        c = record_cut(buf, (n * k) // nshards)
        cuts.append(max(c, cuts[-1]))
    cuts.append(n)
    return [(cuts[k], cuts[k + 1]) for k in range(nshards)]


def translate_shard(buf, rank, world, translate):
    """What rank `rank` of `world` computes: its range of buf through `translate` (bytes -> bytes).
    An empty non-first shard produces nothing (the reference emits its one "always" record only for
    a stream that is empty altogether, transeq.go:213)."""
    lo, hi = shard_ranges(buf, world)[rank]
    if hi == lo and not (rank == 0 and len(buf) == 0):
        return b""
    if hi < len(buf) and _blank_only(buf[lo:hi]):
        return b""  # (only shard 0 can be blank: the others start with a header line)
    return translate(buf[lo:hi])


def _blank_only(part):
    """True if every line of `part` is empty after bufio.ScanLines dropped one trailing CR
    (transeq.go:195-197 skips such lines): alone it would be an input that is empty altogether."""
    if part.strip(b"\r\n"):
        return False
    return all(line in (b"", b"\r") for line in part.split(b"\n"))
