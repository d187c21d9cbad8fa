"""Seeded random FASTA generators for the differential tests (oracle vs model vs CUDA path).

`nasty_fasta` aims at every parser / formatter edge the reference's code has and its goldens do
not pin: CRLF, blank lines, data before the first '>', header-only records,
records of length 0..4, lines of any width, lower case, U, IUPAC letters, '>' inside a line,
spaces and tabs in headers, missing final newline, lone CRs.
"""
import random

ALPHABETS = [
    "ACGT",
    "ACGTN",
    "ACGTacgtNnUu",
    "ACGTRYSWKMBDHVN-*.",
    "AAAAAAAAAAAAAAAAAAAN",   # long runs of one base + N (trim walks)
    "N",
]


def nasty_fasta(seed, max_records=12, max_len=400):
    rng = random.Random(seed)
    eol = rng.choice(["\n", "\n", "\n", "\r\n"])
    parts = []
    if rng.random() < 0.25:  # data before the first header
        parts.append("".join(rng.choice("ACGT") for _ in range(rng.randint(0, 70))) + eol)
    if rng.random() < 0.1:
        parts.append(eol * rng.randint(1, 3))
    for r in range(rng.randint(0, max_records)):
        hdr_kind = rng.random()
        if hdr_kind < 0.3:
            hdr = ">s%d" % r
        elif hdr_kind < 0.6:
            hdr = ">seq%d some description here" % r
        elif hdr_kind < 0.7:
            hdr = ">"
        elif hdr_kind < 0.8:
            hdr = "> leading space %d" % r
        elif hdr_kind < 0.9:
            hdr = ">tab\tseparated %d  two  spaces" % r
        else:
            hdr = ">" + "h" * rng.randint(1, 200)
        parts.append(hdr + eol)
        kind = rng.random()
        if kind < 0.15:
            n = rng.randint(0, 5)
        elif kind < 0.8:
            n = rng.randint(0, max_len)
        else:
            n = rng.randint(0, 4 * max_len)
        alpha = rng.choice(ALPHABETS)
        seq = "".join(rng.choice(alpha) for _ in range(n))
        if rng.random() < 0.1 and n > 3:
            i = rng.randint(0, n - 1)
            seq = seq[:i] + ">" + seq[i + 1:]
        if rng.random() < 0.1 and n > 3:
            i = rng.randint(0, n - 1)
            seq = seq[:i] + "\r" + seq[i + 1:]   # a CR in the middle of a line is a nucleotide (-> N)
        width = rng.choice([60, 60, 60, 70, 80, 1, 2, 3, 7, 61, 10 ** 9])
        lines = [seq[i:i + width] for i in range(0, len(seq), width)] if width < 10 ** 9 else ([seq] if seq else [])
        for ln in lines:
            parts.append(ln + eol)
            if rng.random() < 0.03:
                parts.append(eol)  # blank line inside a record
    data = "".join(parts)
    if data.endswith(eol) and rng.random() < 0.4:
        data = data[: -len(eol)]  # no final newline
        if rng.random() < 0.3:
            data += "\r"           # ... but a lone final CR
    return data.encode("latin-1")


def regular_fasta(seed, n_records, len_lo, len_hi, width=60, header="seq", n_frac=0.001):
    """Well-formed FASTA of the benchmark shapes, python-speed: keep it small."""
    rng = random.Random(seed)
    out = []
    for r in range(n_records):
        n = rng.randint(len_lo, len_hi)
        out.append(">%s%07d len=%d\n" % (header, r, n))
        seq = "".join(rng.choice("ACGT") if rng.random() >= n_frac else "N" for _ in range(n))
        for i in range(0, n, width):
            out.append(seq[i:i + width] + "\n")
    return "".join(out).encode("ascii")


def mid_fasta(seed, target=None):
    """0.02-1.5 MB of records with heavy-tailed lengths (0 .. 120 k), odd line widths, header lines from
    empty to longer than a parse tile, CRLF, blank lines, data before the first header: many tiles, many
    record pieces per tile, blocks and frames ending everywhere."""
    rng = random.Random(seed)
    parts = []
    pick = rng.choice((20_000, 200_000, 1_500_000))
    target = target or pick
    size = 0
    if rng.random() < 0.3:
        parts.append(bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(0, 500))) + b"\n")
    while size < target:
        hl = rng.choice((0, 1, 5, 20, 60, 61, 64, 65, 120, 300, 9000)) if rng.random() < 0.3 else rng.randint(1, 40)
        hdr = b">" + bytes(rng.choice(b"abcXYZ01 _|") for _ in range(hl))
        n = int(rng.choice((0, 1, 2, 3, 59, 60, 61, 179, 180, 181, 189, 190, rng.paretovariate(0.7) * 50))) % 120_000
        alpha = rng.choice((b"ACGT", b"ACGTN", b"ACGTUacgtunRYKM", b"AC"))
        seq = bytes(rng.choice(alpha) for _ in range(n))
        w = rng.choice((60, 70, 80, 1, 7, 10 ** 9, rng.randint(1, 200)))
        eol = b"\r\n" if rng.random() < 0.2 else b"\n"
        lines = [seq[i:i + w] for i in range(0, n, w)]
        if rng.random() < 0.1 and lines:
            lines.insert(rng.randrange(len(lines) + 1), b"")
        rec = hdr + eol + eol.join(lines) + (eol if rng.random() < 0.9 else b"")
        parts.append(rec)
        size += len(rec)
    return b"".join(parts)


ALL_FRAMES = ["1", "2", "3", "F", "-1", "-2", "-3", "R", "6"]
ALL_TABLES = [0, 2, 3, 4, 5, 6, 9, 10, 11, 12, 13, 14, 16, 21, 22, 23, 24, 25, 26, 29, 30]


def random_options(seed):
    rng = random.Random(seed * 7919 + 13)
    return dict(
        frame=rng.choice(ALL_FRAMES + ["6", "6"]),
        table=rng.choice(ALL_TABLES),
        clean=rng.random() < 0.4,
        alternative=rng.random() < 0.4,
        trim=rng.random() < 0.4,
    )
