"""Parity of the CUDA path (through the C ABI) with the CPU oracle.  Needs a B200: -m gpu."""
import io

import pytest

import oracle
from fasta_gen import mid_fasta, nasty_fasta, random_options, regular_fasta, ALL_FRAMES, ALL_TABLES
from optparse_ref import parse_option_string

pytestmark = pytest.mark.gpu


def gpu(data, frame="1", table=0, clean=False, alternative=False, trim=False, **kw):
    import gotranseq_b200
    return gotranseq_b200.translate_bytes(data, frame=frame, table=table, clean=clean,
                                          alternative=alternative, trim=trim, **kw)


@pytest.mark.parametrize("idx", range(29))
def test_reference_goldens(goldens, idx):
    """transeq/transeq_test.go:14-62 -- the reference's own table-driven test, on the GPU path."""
    case = goldens["cases"][idx]
    kw = parse_option_string(case["options"])
    assert gpu(goldens["input"].encode(), **kw).decode() == case["expected"]


def test_quirks():
    for data in (b"", b"\n", b"\r", b"\r\n", b">", b">\n", b">\r", b"A", b"\n\nAC", b">a\n>b\n>c",
                 b"AC\r", b">a b\r\nAC\r\r\nG", b"x>y\n>z", b"ACGT\n>x\nACG", b">a b c\nATG"):
        for frame in ("6", "F", "R", "2"):
            for trim in (False, True):
                assert gpu(data, frame=frame, trim=trim) == oracle.translate(data, frame=frame, trim=trim), (data, frame, trim)


@pytest.mark.parametrize("seed", range(120))
def test_random_nasty(seed):
    data = nasty_fasta(seed)
    kw = random_options(seed)
    assert gpu(data, **kw) == oracle.translate(data, **kw), (seed, kw)


@pytest.mark.parametrize("seed", range(40))
def test_random_mid_size(seed):
    data = mid_fasta(5000 + seed)
    kw = random_options(5000 + seed)
    assert gpu(data, **kw) == oracle.translate(data, **kw), (seed, kw)


@pytest.mark.parametrize("frame", ALL_FRAMES)
def test_all_frames_all_tables(frame):
    data = nasty_fasta(1000, max_records=30, max_len=900)
    for table in ALL_TABLES:
        for clean in (False, True):
            want = oracle.translate(data, frame=frame, table=table, clean=clean, trim=True)
            assert gpu(data, frame=frame, table=table, clean=clean, trim=True) == want, (frame, table, clean)


def test_errors_match_reference_texts():
    import gotranseq_b200
    with pytest.raises(gotranseq_b200.TranseqError, match=r"wrong value for -f \| --frame parameter: 7"):
        gpu(b">a\nACG\n", frame="7")
    with pytest.raises(gotranseq_b200.TranseqError, match=r"invalid table code: 1$"):
        gpu(b">a\nACG\n", frame="1", table=1)


@pytest.mark.parametrize("shape", ["readme", "long", "short_reads", "one_line_records"])
def test_regular_shapes(shape):
    """Benchmark-shaped inputs at sizes the oracle finishes in a second."""
    if shape == "readme":
        data = regular_fasta(42, 600, 300, 3000)
    elif shape == "long":
        data = regular_fasta(43, 3, 200_000, 400_000, n_frac=0.0001)
    elif shape == "short_reads":
        data = regular_fasta(45, 20_000, 150, 150, width=10 ** 9, header="A00123:45:HXXXXXXXX:1:1101:")
    else:
        data = regular_fasta(46, 50, 20_000, 60_000, width=10 ** 9)
    for kw in (dict(frame="6"), dict(frame="6", table=11, clean=True, trim=True), dict(frame="R", alternative=True),
               dict(frame="6", table=2)):
        assert gpu(data, **kw) == oracle.translate(data, **kw), (shape, kw)


def test_translate_stream_matches_one_shot():
    import gotranseq_b200
    data = regular_fasta(7, 3000, 100, 2000)
    want = oracle.translate(data, frame="6", trim=True)
    for chunk in (1 << 16, 1 << 20):
        out = io.BytesIO()
        with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6", Trim=True), chunk_bytes=chunk) as t:
            t.translate_stream(io.BytesIO(data), out)
            # the context is reusable for a second stream
            out2 = io.BytesIO()
            t.translate_stream(io.BytesIO(data), out2)
        assert out.getvalue() == want
        assert out2.getvalue() == want


def test_translate_stream_overlapped_pipeline():
    """The streaming interface keeps a worker translating while the caller reads and writes: small
    reads (the input buffer fills over many calls), records larger than a chunk (the buffer grows),
    a writer that lags, and more data behind the last record boundary of every chunk."""
    import random
    import time
    import gotranseq_b200

    class Dribble:
        def __init__(self, data, rng):
            self.data, self.pos, self.rng = data, 0, rng

        def read(self, n):
            k = min(n, self.rng.choice((1, 100, 4096, 70000, 1 << 20)))
            out = self.data[self.pos:self.pos + k]
            self.pos += len(out)
            return out

    class SlowWriter:
        def __init__(self):
            self.parts = []

        def write(self, b):
            time.sleep(0.002)
            self.parts.append(bytes(b))

    rng = random.Random(21)
    data = regular_fasta(8, 400, 100, 3000) + regular_fasta(9, 2, 300_000, 400_000) + regular_fasta(10, 300, 10, 900)
    for kw, okw in ((dict(Frame="6"), dict(frame="6")), (dict(Frame="R", Alternative=True, Trim=True),
                                                          dict(frame="R", alternative=True, trim=True))):
        want = oracle.translate(data, **okw)
        with gotranseq_b200.Translator(gotranseq_b200.Options(**kw), chunk_bytes=1 << 17) as t:
            for _ in range(2):
                w = SlowWriter()
                t.translate_stream(Dribble(data, rng), w)
                assert b"".join(w.parts) == want


def test_translate_api():
    import gotranseq_b200
    data = nasty_fasta(5)
    out = io.BytesIO()
    gotranseq_b200.Translate(io.BytesIO(data), out, gotranseq_b200.Options(Frame="6", Table=11, Clean=True))
    assert out.getvalue() == oracle.translate(data, frame="6", table=11, clean=True)


def test_device_resident_entry_point():
    import torch
    import gotranseq_b200
    data = regular_fasta(11, 500, 500, 5000)
    want = oracle.translate(data, frame="6")
    d_in = torch.frombuffer(bytearray(data), dtype=torch.uint8).cuda()
    d_out = torch.empty(len(want) + 100, dtype=torch.uint8, device="cuda")
    with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6")) as t:
        n = t.translate_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), d_out.numel(),
                               torch.cuda.current_stream().cuda_stream)
        assert n == len(want)
        assert bytes(d_out[:n].cpu().numpy()) == want
        # too small: nothing written, exact size reported
        with pytest.raises(gotranseq_b200.TranseqError, match=str(len(want))):
            t.translate_device(d_in.data_ptr(), d_in.numel(), d_out.data_ptr(), 10, 0)
        st = t.stats()
        assert st["nucleotides"] > 0 and st["kernel_launches"] >= 4


def test_cli_reference_goldens(goldens, tmp_path):
    """The `gotranseq` CLI (gotranseq_b200/host/main.cpp) with the reference's flags (main.go:21-37)."""
    import os
    import subprocess
    exe = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gotranseq_b200", "bin", "gotranseq")
    src = tmp_path / "test.fna"
    src.write_bytes(goldens["input"].encode())
    want_warn = oracle.translate_with_warnings(goldens["input"].encode(), frame="1")[1]
    for case in goldens["cases"][::3]:
        flags = ["-" + t if t.startswith("-") else t for t in case["options"].split()]  # -frame=6 -> --frame=6
        dst = tmp_path / "out.faa"
        r = subprocess.run([exe, "-s", str(src), "-o", str(dst)] + flags, capture_output=True, timeout=120)
        # stdout carries the WARNING lines of encodedSequence.go:39 (sequence13 holds S and K), nothing else
        assert r.returncode == 0 and r.stdout == want_warn, r.stdout
        assert dst.read_bytes().decode() == case["expected"], case["options"]
    # errors are printed, exit code stays 0 (main.go:87-90)
    r = subprocess.run([exe, "-s", str(src), "-o", str(tmp_path / "o"), "-f", "9"], capture_output=True, text=True)
    assert r.returncode == 0 and "wrong value for -f | --frame parameter: 9" in r.stdout


@pytest.mark.parametrize("name", ["readme189", "contigs", "chromosome", "short_reads"])
def test_benchmark_shapes_full_size(name):
    """BASELINE.json configs at (or near) their full sizes, byte-compared with the oracle, plus
    size-independent properties: sharding invariance and frame decomposition."""
    import numpy as np
    import gotranseq_b200
    import workloads
    from gotranseq_b200 import shard
    gen, opts = workloads.WORKLOADS[name]
    if name == "contigs":
        data = gen(n_records=12)            # 60 MB of the 25 GB config
    elif name == "short_reads":
        data = gen(n_reads=1_000_000)       # 204 MB of the 20 GB config
    else:
        data = gen()
    want = oracle.translate(np.frombuffer(data, dtype=np.uint8), num_worker=1, **opts)
    O = gotranseq_b200.Options(Frame=opts["frame"], Table=opts.get("table", 0), Clean=opts.get("clean", False),
                               Alternative=opts.get("alternative", False), Trim=opts.get("trim", False))
    with gotranseq_b200.Translator(O) as t:
        got = t.translate_bytes(data)
        assert len(got) == len(want)
        assert got == want
        # record sharding (what each GPU of a multi-GPU run computes) leaves the bytes unchanged
        if name != "chromosome":
            parts = [t.translate_bytes(data[a:b]) for a, b in shard.shard_ranges(data, 3) if b > a]
            assert b"".join(parts) == want


def test_over_long_lines():
    """transeq.go:27,186: the reference's scanner gives up at a line of >= 100 MiB and the rest of
    the input is silently ignored.  Checked with the limit lowered on both sides."""
    import random
    import gotranseq_b200
    cap = 20000
    rng = random.Random(5)
    seq = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    rec = lambda i, n, w=60: ">r%d\n" % i + "\n".join(seq(n)[j:j + w] for j in range(0, n, w)) + "\n"
    cases = {
        "long sequence line in the middle": rec(0, 500) + ">long\n" + seq(30000) + "\n" + rec(1, 300),
        "long header line": rec(0, 500) + ">" + "h" * 25000 + "\nACGT\n" + rec(1, 300),
        "first line too long": seq(cap + 5) + "\n" + rec(0, 100),
        "unterminated last line too long": rec(0, 700) + rec(1, 50) + ">x\n" + seq(cap),
        "exactly at the limit": rec(0, 90) + ">y\n" + seq(cap) + "\n" + rec(1, 90),
        "one below the limit": rec(0, 90) + ">y\n" + seq(cap - 1) + "\n" + rec(1, 90),
        "CR counts": rec(0, 90) + ">y\n" + seq(cap - 1) + "\r\n" + rec(1, 90),
        "many lines, none too long": rec(0, 90) + ">z\n" + "\n".join(seq(cap - 1) for _ in range(4)) + "\n" + rec(2, 10),
        "second of two long lines": rec(0, 10) + ">a\n" + seq(cap - 1) + "\n" + seq(cap + 1) + "\n" + seq(3 * cap) + "\n" + rec(1, 10),
        "far into a large input": "".join(rec(i, 3000) for i in range(60)) + ">q\n" + seq(2 * cap) + "\n" + rec(99, 30),
    }
    oracle.set_max_line_length(cap)
    try:
        for name, text in cases.items():
            data = text.encode()
            for kw in (dict(frame="6"), dict(frame="F", trim=True)):
                want = oracle.translate(data, **kw)
                O = gotranseq_b200.Options(Frame=kw["frame"], Trim=kw.get("trim", False))
                with gotranseq_b200.Translator(O, chunk_bytes=1 << 16) as t:
                    t.set_max_line_length(cap)
                    assert t.translate_bytes(data) == want, name
                    out = io.BytesIO()
                    t.translate_stream(io.BytesIO(data), out)
                    assert out.getvalue() == want, name + " (stream)"
    finally:
        oracle.set_max_line_length(100 << 20)


# ---- shapes that stress k_lines' tile / block / batch geometry ----

def _seq(rng, n):
    return bytes(rng.choice(b"ACGTN") for _ in range(n))


def test_tiles_without_symbols_and_short_tiles():
    """Header lines longer than a parse tile (8192 bytes): tiles with no symbols at all, and the
    192 symbols behind a tile have to be collected from several later slots."""
    import random
    rng = random.Random(11)
    parts = []
    for r in range(12):
        hdr = b">" + bytes(rng.choice(b"abcdefgh ") for _ in range(rng.choice((5, 9000, 17000, 8191, 8192, 8193))))
        parts.append(hdr + b"\n" + _seq(rng, rng.choice((0, 1, 2, 50, 179, 180, 181, 400, 9000))) + b"\n")
    data = b"".join(parts)
    for kw in (dict(frame="6"), dict(frame="6", trim=True, alternative=True), dict(frame="R"), dict(frame="2")):
        assert gpu(data, **kw) == oracle.translate(data, **kw), kw


def test_header_geometry_sweep():
    """Header lines of every length around the 16-byte lanes / 64-byte steps of k_records and the
    4-byte lanes / 64-byte steps of k_headers, at every alignment of the '>' in the input, with the
    first space (writer.go:121) at the start, in the middle, at the end or absent, LF and CRLF line
    ends, and a last header line that ends with the input (with and without a CR)."""
    import random
    rng = random.Random(21)
    parts = []
    for H in list(range(1, 40)) + list(range(58, 72)) + [127, 128, 129, 190, 200, 257]:
        for variant in range(4):
            body = bytearray(rng.choice(b"abcdefghijklmnopqrstuvwxyz0123456789_|.") for _ in range(H - 1))
            if variant == 1 and H > 2:
                body[rng.randrange(len(body))] = 0x20
            elif variant == 2 and H > 1:
                body[-1] = 0x20
            elif variant == 3 and H > 1:
                body[0] = 0x20
            eol = b"\r\n" if (H + variant) % 3 == 0 else b"\n"
            # (the sequence length moves the next '>' to another alignment)
            parts.append(b">" + bytes(body) + eol + _seq(rng, rng.randrange(0, 70)) + eol)
    data = b"".join(parts)
    for tail in (b">last header without a newline", b">last header with a CR\r", b">x", b">"):
        d = data + tail
        for kw in (dict(frame="6"), dict(frame="1", trim=True), dict(frame="R", alternative=True)):
            assert gpu(d, **kw) == oracle.translate(d, **kw), (tail, kw)


def test_thousands_of_tiny_records_per_tile():
    """More record pieces in a tile than one batch of the run table holds (16), up to 4096 empty ones."""
    import random
    rng = random.Random(12)
    parts = []
    for r in range(9000):
        parts.append(b">" + (b"r%d" % r if r % 3 else b"") + b"\n" + _seq(rng, rng.choice((0, 0, 0, 1, 2, 3, 4, 7))) +
                     (b"\n" if r % 5 else b""))
    data = b"".join(parts)
    for kw in (dict(frame="6"), dict(frame="6", trim=True), dict(frame="F", table=4), dict(frame="-2", alternative=True)):
        assert gpu(data, **kw) == oracle.translate(data, **kw), kw


def test_image_capacity_batches():
    """Sixteen record pieces of a tile whose lines do not fit the output image together."""
    import random
    rng = random.Random(13)
    parts = []
    for r in range(400):
        n = rng.choice((420, 480, 500, 511, 540, 600))
        parts.append(b">q%d\n" % r + _seq(rng, n) + b"\n")  # one-line records: 16+ of them per 8 KB tile
    data = b"".join(parts)
    for kw in (dict(frame="6"), dict(frame="6", trim=True, clean=True)):
        assert gpu(data, **kw) == oracle.translate(data, **kw), kw


@pytest.mark.parametrize("frame", ["6", "F", "R", "1", "-3"])
def test_block_and_tile_boundary_sweep(frame):
    """One record whose length moves the 180-nucleotide blocks and the ends of the frames across the
    8192-byte tile boundaries, with and without a second record behind it."""
    import random
    rng = random.Random(14)
    base = _seq(rng, 3 * 8192 + 400)
    for n in list(range(8180, 8200)) + list(range(16370, 16392)) + [8192 + 177, 8192 + 180, 8192 + 183, 24576, 24577]:
        for width in (10 ** 9, 60):
            seq = base[:n]
            body = b"\n".join(seq[i:i + width] for i in range(0, n, width))
            data = b">x\n" + body + b"\n>y\nACGTAC\n"
            for alt in (False, True):
                want = oracle.translate(data, frame=frame, alternative=alt, trim=(n % 2 == 0))
                assert gpu(data, frame=frame, alternative=alt, trim=(n % 2 == 0)) == want, (n, width, alt)


def test_offsets_beyond_32_bits():
    """2.2 GB in, 4.4 GB out in one device pass: input offsets beyond 2^31, output offsets beyond 2^32."""
    import hashlib
    import numpy as np
    import torch
    import gotranseq_b200
    import workloads
    data = workloads.contigs(n_records=430, len_lo=4_500_000, len_hi=5_500_000, seed=7)
    cut = 1_000_000_007
    data = data[:cut] + b"\n>tiny x\r\nACGTNACG\r\n>e\n\n" + data[cut:]  # small records between the big offsets
    n = len(data)
    arr = np.frombuffer(data, dtype=np.uint8)
    with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6", Table=11, Clean=True)) as tr:
        d_in = torch.from_numpy(arr.copy()).cuda()
        cap = 2 * n + n // 8 + (1 << 20)
        d_out = torch.empty(cap, dtype=torch.uint8, device="cuda")
        out_n = tr.translate_device(d_in.data_ptr(), n, d_out.data_ptr(), cap, torch.cuda.current_stream().cuda_stream)
        got = d_out[:out_n].cpu().numpy()
        del d_in, d_out
    want = oracle.translate(arr, frame="6", table=11, clean=True)
    assert n > 2 ** 31 and out_n > 2 ** 32 and out_n == len(want)
    assert hashlib.sha256(got.tobytes()).digest() == hashlib.sha256(want).digest()
