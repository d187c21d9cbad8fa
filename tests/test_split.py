"""One record split into pieces (gotranseq_b200/split.py).

CPU part: the host logic (cuts, carry planning, assembly) against a fake translator that answers
from the oracle, with the segment geometry restated here in Python.  GPU part: the real pieces
through the C ABI, assembled, against the oracle."""
import random

import pytest

import oracle
from gotranseq_b200 import split

CODES = {c: 0 for c in range(256)}
for ch, v in (("A", 1), ("C", 2), ("T", 3), ("U", 3), ("G", 4)):
    CODES[ord(ch)] = v
    CODES[ord(ch.lower())] = v

FRAME_MASKS = {"1": [0], "2": [1], "3": [2], "F": [0, 1, 2], "-1": [3], "-2": [4], "-3": [5], "R": [3, 4, 5],
               "6": [0, 1, 2, 3, 4, 5]}


def nucleotides(piece, first):
    """Bytes of a piece that the reference keeps as nucleotides (transeq.go:183-216): header line
    dropped, line ends ('\\n', '\\r\\n') dropped."""
    lines = piece.split(b"\n")
    if first:
        lines = lines[1:]
    out = bytearray()
    for k, ln in enumerate(lines):
        last = k == len(lines) - 1
        if ln.endswith(b"\r") and not last:
            ln = ln[:-1]
        elif ln.endswith(b"\r") and last:
            ln = ln[:-1]  # bufio.ScanLines drops a CR at EOF too
        out += ln
    return bytes(out)


def reverse_start(n, i, alt):
    return i if alt else (n % 3 - i) % 3


def frame_lengths(n, alt):
    L = [(n - k + 2) // 3 if n >= k else 0 for k in range(3)]
    for i in range(3):
        p = reverse_start(n, i, alt)
        L.append((n - p + 2) // 3 if n >= p else 0)
    return L


def segments(frames, alt, H, n, before, m, first, last, trim_len=None):
    """Byte ranges of the record's output owned by the piece holding nucleotides [before, before+m).
    Output line b of a strand = one block of 180 nucleotides, owned by the piece that holds the
    block's anchor (the last symbol of its lowest codon): forward 2 + 180 b; reverse n - 189 - 180 b,
    or 0 for the blocks at the record's start (gts_lines.cu)."""
    L = list(trim_len) if trim_len is not None else frame_lengths(n, alt)
    a, e = before, before + m + (2 if last else 0)
    f_lo = -((2 - a) // 180) if a > 2 else 0
    f_hi = -((2 - e) // 180) if e > 2 else 0
    Lc = (n - 189) // 180 + 1 if n >= 189 else 0
    r_lo = (n - 189 - e) // 180 + 1 if n - 189 >= e else 0
    r_hi = min((n - 189 - a) // 180 + 1 if n - 189 >= a else 0, Lc)
    if a == 0 and e > 0:
        if r_hi <= r_lo:
            r_lo = Lc
        r_hi = 1 << 62
    segs, base = [], 0
    for f in frames:
        body = base + H + 3
        if first:
            segs.append((base, H + 3))
        NL = (L[f] + 59) // 60
        base += H + 3 + L[f] + NL
        lo, hi = (f_lo, f_hi) if f < 3 else (r_lo, r_hi)
        hi = min(hi, NL)
        if hi <= lo:
            continue
        segs.append((body + 61 * lo, 61 * (hi - lo) - (60 * NL - L[f] if hi == NL else 0)))
    return base, [s for s in segs if s[1]]


class FakeTranslator:
    """Answers piece_scan / piece_trim / piece_translate from the oracle: the output of the WHOLE
    record for the bytes, the reference's codon array for the trim walk (restated here on the
    piece's own nucleotides plus the carried / following codes, like the device does it)."""

    def __init__(self, whole, frame, alt, trim=False, table=0, clean=False):
        self.whole, self.frame, self.alt, self.trim = whole, frame, alt, trim
        self.kw = dict(frame=frame, alternative=alt, trim=trim, table=table, clean=clean)
        self.codes = oracle.code_array(table, clean)

    def piece_scan(self, data, first):
        self.nts = nucleotides(data, first)
        hl = 0
        if first:
            h = data.split(b"\n")[0]
            hl = len(h[:-1] if h.endswith(b"\r") else h)
        codes = [CODES[c] for c in self.nts[-2:]]
        tail = (codes[-1] if codes else 0, codes[-2] if len(codes) > 1 else 0)
        head = bytes(CODES[c] for c in self.nts[:192])
        return dict(nucleotides=len(self.nts), headers=1 if first else 0, header_len=hl, tail=tail,
                    head=head + bytes(192 - len(head)))

    def frame_lengths(self, n):
        L = frame_lengths(n, self.alt)
        return [L[f] if f in FRAME_MASKS[self.frame] else 0 for f in range(6)]

    def piece_trim(self, p, lens):
        """writer.go:141-163 on the residues whose codon starts in this piece."""
        comp = lambda c: 0 if c == 0 else ((c - 1) ^ 2) + 1
        own = [CODES[c] for c in self.nts]
        m, n, before = len(own), p["nt_total"], p["nt_before"]
        nh = bytes(p.get("next_head", b""))

        def sym(i):
            if i < 0:
                return 0 if p["first"] else (p["carry"][0] if i == -1 else p["carry"][1])
            if i >= m:
                return 0 if p["last"] else (nh[i - m] if i - m < len(nh) else 0)
            return own[i]

        out, res = list(lens), [False] * 6
        for f in range(6):
            if f not in FRAME_MASKS[self.frame] or out[f] == 0:
                res[f] = True
                continue
            lo = -2 if p["first"] else 0
            while True:
                j = out[f] - 1
                q = f + 3 * j if f < 3 else n - 3 - reverse_start(n, f - 3, self.alt) - 3 * j
                i = q - before
                if i < lo or i >= m:
                    break
                c0, c1, c2 = sym(i), sym(i + 1), sym(i + 2)
                aa = self.codes[c0 | c1 << 8 | c2 << 16] if f < 3 else self.codes[comp(c2) | comp(c1) << 8 | comp(c0) << 16]
                if aa not in b"X*":
                    res[f] = True
                    break
                out[f] -= 1
                if out[f] == 0:
                    res[f] = True
                    break
        return out, res

    def piece_translate(self, p):
        want = oracle.translate(self.whole, **self.kw)
        total, segs = segments(FRAME_MASKS[self.frame], self.alt, p["header_len"], p["nt_total"], p["nt_before"],
                               len(self.nts), p["first"], p["last"], p.get("trim_len"))
        assert total == len(want)
        return total, [(o, want[o:o + n]) for o, n in segs]

    def close(self):
        pass


def one_record(seed, n, width=60, crlf=False, header=b">chr1 test record"):
    rng = random.Random(seed)
    seq = bytes(rng.choice(b"ACGTNacgtRYK") for _ in range(n))
    eol = b"\r\n" if crlf else b"\n"
    body = eol.join(seq[i:i + width] for i in range(0, n, width))
    return header + eol + body + (eol if rng.random() < 0.5 else b"")


def test_piece_ranges_cover_and_cut_at_line_starts():
    for seed in range(20):
        buf = one_record(seed, 50 + 37 * seed, width=7 + seed)
        for k in (1, 2, 3, 8, 40):
            r = split.piece_ranges(buf, k)
            assert r[0][0] == 0 and r[-1][1] == len(buf)
            assert all(r[i][1] == r[i + 1][0] for i in range(k - 1))
            hdr_end = buf.find(b"\n") + 1
            for lo, _ in r[1:]:
                assert lo >= hdr_end and (lo == len(buf) or buf[lo - 1:lo] == b"\n")


def test_plan_pieces_carry():
    infos = [dict(nucleotides=5, headers=1, header_len=4, tail=(3, 2), head=bytes([1, 2, 3, 2, 3]) + bytes(187)),
             dict(nucleotides=0, headers=0, header_len=0, tail=(0, 0), head=bytes(192)),
             dict(nucleotides=1, headers=0, header_len=0, tail=(4, 0), head=bytes([4]) + bytes(191)),
             dict(nucleotides=7, headers=0, header_len=0, tail=(1, 1), head=bytes([2, 2, 2, 2, 2, 1, 1]) + bytes(185))]
    p = split.plan_pieces(infos)
    assert [x["nt_before"] for x in p] == [0, 5, 5, 6]
    assert all(x["nt_total"] == 13 and x["header_len"] == 4 for x in p)
    assert [x["carry"] for x in p] == [(0, 0), (3, 2), (3, 2), (4, 3)]
    assert [x["first"] for x in p] == [True, False, False, False] and p[-1]["last"] and not p[0]["last"]
    assert [x["next_head"] for x in p] == [bytes([4, 2, 2, 2, 2, 2, 1, 1]), bytes([4, 2, 2, 2, 2, 2, 1, 1]),
                                           bytes([2, 2, 2, 2, 2, 1, 1]), b""]


@pytest.mark.parametrize("frame", ["1", "2", "3", "F", "-1", "-2", "-3", "R", "6"])
def test_split_host_logic_against_oracle(frame):
    for seed in range(12):
        n = [0, 1, 2, 3, 5, 59, 60, 61, 179, 180, 181, 1000][seed]
        buf = one_record(seed, n, width=(60, 7, 13)[seed % 3], crlf=seed % 4 == 1)
        for alt in (False, True):
            want = oracle.translate(buf, frame=frame, alternative=alt)
            for k in (1, 2, 3, 5):
                got = split.translate_split(buf, k, lambda: FakeTranslator(buf, frame, alt))
                assert got == want, (seed, frame, alt, k)


def trim_record(seed, n, width=60):
    """A record whose ends translate to long runs of 'X' / '*': N runs, stop codons, lower case."""
    rng = random.Random(seed)
    def run(k):
        kind = rng.random()
        if kind < 0.4:
            return b"N" * k
        if kind < 0.7:
            return (b"TAA" * (k // 3 + 1))[:k]
        return bytes(rng.choice(b"NNNTAGX") for _ in range(k))
    mid = bytes(rng.choice(b"ACGTacgtN") for _ in range(n))
    seq = run(rng.choice((0, 1, 5, 200, 2000))) + mid + run(rng.choice((0, 2, 7, 300, 2500)))
    return b">t%d some words\n" % seed + b"\n".join(seq[i:i + width] for i in range(0, len(seq), width)) + b"\n"


@pytest.mark.parametrize("frame", ["1", "F", "-2", "R", "6"])
def test_split_trim_host_logic_against_oracle(frame):
    """--trim on a split record: the walk over the pieces (split.settle_trim) with the fake pieces."""
    for seed in range(10):
        n = [0, 1, 4, 60, 181, 700, 3000, 10, 999, 2][seed]
        buf = trim_record(seed, n, width=(60, 7, 13)[seed % 3])
        for alt in (False, True):
            for table, clean in ((0, False), (11, True)):
                want = oracle.translate(buf, frame=frame, alternative=alt, trim=True, table=table, clean=clean)
                for k in (1, 2, 3, 5, 9):
                    got = split.translate_split(buf, k, lambda: FakeTranslator(buf, frame, alt, True, table, clean), trim=True)
                    assert got == want, (seed, frame, alt, table, k)
    # a record that is nothing but N: every piece hands the walk on to the next one
    buf = b">n\n" + b"\n".join([b"N" * 60] * 40) + b"\n"
    want = oracle.translate(buf, frame="6", trim=True)
    assert split.translate_split(buf, 7, lambda: FakeTranslator(buf, "6", False, True), trim=True) == want


def test_assemble_rejects_gaps():
    with pytest.raises(ValueError):
        split.assemble(10, [[(0, b"abc")], [(4, b"defghi")]])
    with pytest.raises(ValueError):
        split.assemble(10, [[(0, b"abc")], [(3, b"def")]])


def _split_worker(rank, world, port, buf, frame, alt, q):
    import os
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def all_gather(obj):
        objs = [None] * world
        dist.all_gather_object(objs, obj)
        return objs

    total, segs = split.translate_split_rank(buf, rank, world, FakeTranslator(buf, frame, alt, trim=True), all_gather,
                                             trim=True)
    q.put((rank, total, segs))
    dist.barrier()
    dist.destroy_process_group()


def test_split_two_ranks_gloo():
    """The exchange between the ranks of a split record (three integers per piece), world_size 2."""
    import socket
    import torch.multiprocessing as mp
    buf = trim_record(3, 5000, width=60)
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_split_worker, args=(r, 2, port, buf, "6", False, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == res[1][1]
    assert split.assemble(res[0][1], [r[2] for r in res]) == oracle.translate(buf, frame="6", trim=True)


# ---- GPU: the real pieces ----

def make_gpu(frame, alt, table=0, clean=False, trim=False):
    import gotranseq_b200
    return lambda: gotranseq_b200.Translator(gotranseq_b200.Options(Frame=frame, Alternative=alt, Table=table,
                                                                    Clean=clean, Trim=trim))


@pytest.mark.gpu
@pytest.mark.parametrize("frame", ["1", "2", "3", "F", "-1", "-2", "-3", "R", "6"])
def test_gpu_split_small(frame):
    for seed in range(12):
        n = [0, 1, 2, 3, 5, 59, 60, 61, 179, 180, 181, 1000][seed]
        buf = one_record(seed, n, width=(60, 7, 13)[seed % 3], crlf=seed % 4 == 1)
        for alt in (False, True):
            want = oracle.translate(buf, frame=frame, alternative=alt)
            for k in (1, 2, 3, 5):
                assert split.translate_split(buf, k, make_gpu(frame, alt)) == want, (seed, frame, alt, k)


@pytest.mark.gpu
@pytest.mark.parametrize("npieces", [1, 2, 3, 8])
def test_gpu_split_chromosome_shaped(npieces):
    """BASELINE config 4's shape (-f R -a, one long record, 60-column lines), and all six frames."""
    buf = one_record(77, 3_000_003 + npieces, width=60, header=b">chr21 Homo sapiens chromosome 21")
    for frame, alt, table, clean in (("R", True, 0, False), ("6", False, 11, True), ("F", False, 0, False)):
        want = oracle.translate(buf, frame=frame, alternative=alt, table=table, clean=clean)
        assert split.translate_split(buf, npieces, make_gpu(frame, alt, table=table, clean=clean)) == want


@pytest.mark.gpu
def test_gpu_split_ragged_lines_and_tiny_pieces():
    rng = random.Random(5)
    for seed in range(30):
        n = rng.randrange(0, 40000)
        lines, seq = [], bytes(rng.choice(b"ACGTN") for _ in range(n))
        i = 0
        while i < n:
            w = rng.choice((1, 2, 3, 10, 61, 500, 9000))
            lines.append(seq[i:i + w])
            if rng.random() < 0.05:
                lines.append(b"")  # blank line
            i += w
        eol = b"\r\n" if seed % 3 == 0 else b"\n"
        buf = b">r" + str(seed).encode() + b" x y" + eol + eol.join(lines) + (eol if seed % 2 else b"")
        frame = rng.choice(["6", "R", "F", "-2", "3"])
        alt = rng.random() < 0.5
        want = oracle.translate(buf, frame=frame, alternative=alt)
        for k in (2, 4, 7):
            assert split.translate_split(buf, k, make_gpu(frame, alt)) == want, (seed, frame, alt, k)


@pytest.mark.gpu
@pytest.mark.parametrize("frame", ["1", "F", "-3", "R", "6"])
def test_gpu_split_trim(frame):
    """--trim on a split record through the C ABI (gts_piece_trim): ends that are long runs of X / *."""
    for seed in range(10):
        n = [0, 1, 4, 60, 181, 700, 30000, 10, 999, 2][seed]
        buf = trim_record(seed, n, width=(60, 7, 13)[seed % 3])
        for alt in (False, True):
            for table, clean in ((0, False), (11, True)):
                want = oracle.translate(buf, frame=frame, alternative=alt, trim=True, table=table, clean=clean)
                for k in (1, 2, 3, 5, 9):
                    got = split.translate_split(buf, k, make_gpu(frame, alt, table=table, clean=clean, trim=True), trim=True)
                    assert got == want, (seed, frame, alt, table, k)
    buf = b">n\n" + b"\n".join([b"N" * 60] * 4000) + b"\n"  # nothing but N: the walk crosses every piece
    want = oracle.translate(buf, frame="6", trim=True)
    assert split.translate_split(buf, 5, make_gpu("6", False, trim=True), trim=True) == want
    buf = one_record(78, 2_000_000, width=60)
    want = oracle.translate(buf, frame="6", table=11, clean=True, trim=True)
    assert split.translate_split(buf, 4, make_gpu("6", False, table=11, clean=True, trim=True), trim=True) == want


@pytest.mark.gpu
def test_gpu_split_planned_on_the_device():
    """gts_piece_scan_async -> (all-gather) -> gts_piece_translate_async -> gts_piece_result: the piece
    description is computed by k_piece_plan from the gathered 256-byte summaries.  One process plays all
    ranks: one Translator per piece, the gather is a copy into one device tensor."""
    import torch
    from gotranseq_b200 import _lib
    IB = _lib.GTS_PIECE_INFO_BYTES
    rng = random.Random(9)
    cases = [one_record(90 + i, n, width=w, crlf=(i % 3 == 1)) for i, (n, w) in enumerate(
        [(0, 60), (1, 60), (2, 7), (5, 60), (181, 60), (1000, 13), (40_000, 60), (700_001, 60)])]
    for buf in cases:
        for frame, alt in (("6", False), ("R", True), ("F", False), ("-2", False)):
            want = oracle.translate(buf, frame=frame, alternative=alt)
            for world in (1, 2, 3, 8):
                ranges = split.piece_ranges(buf, world)
                trs = [make_gpu(frame, alt)() for _ in ranges]
                d_ins = [torch.frombuffer(bytearray(buf[lo:hi] or b"\0"), dtype=torch.uint8).cuda() for lo, hi in ranges]
                d_all = torch.zeros(IB * world, dtype=torch.uint8, device="cuda")
                d_out = torch.full((len(want) + 64,), 0xEE, dtype=torch.uint8, device="cuda")
                for k, (tr, (lo, hi)) in enumerate(zip(trs, ranges)):
                    tr.piece_scan_async(d_ins[k].data_ptr(), hi - lo, d_all[IB * k:].data_ptr(), 0)
                torch.cuda.synchronize()
                seg_lists = []
                for k, (tr, (lo, hi)) in enumerate(zip(trs, ranges)):
                    tr.piece_translate_async(d_ins[k].data_ptr(), hi - lo, k, world, d_all.data_ptr(), d_out.data_ptr(),
                                             d_out.numel(), 0)
                    total, segs, piece = tr.piece_result()
                    assert total == len(want)
                    host = bytes(d_out[:total].cpu().numpy())
                    seg_lists.append([(o, host[o:o + n]) for o, n in segs])
                for tr in trs:
                    tr.close()
                assert split.assemble(len(want), seg_lists) == want, (len(buf), frame, alt, world)
    _ = rng


@pytest.mark.gpu
def test_gpu_split_argument_errors():
    import gotranseq_b200
    tr = make_gpu("6", False)()
    with pytest.raises(gotranseq_b200.TranseqError, match="exactly one header"):
        tr.piece_scan(b">a\nACG\n>b\nAC\n", True)
    with pytest.raises(gotranseq_b200.TranseqError, match="start with the header"):
        tr.piece_scan(b"ACG\n>b\nAC\n", True)
    with pytest.raises(gotranseq_b200.TranseqError, match="only the first piece"):
        tr.piece_scan(b"ACG\n>b\nAC\n", False)
    tr.close()
    tr = make_gpu("6", False, trim=True)()
    info = tr.piece_scan(b">a\nACG\n", True)
    piece = split.plan_pieces([info])[0]
    with pytest.raises(gotranseq_b200.TranseqError, match="needs the trimmed frame lengths"):
        tr.piece_translate(piece)
    tr.close()
