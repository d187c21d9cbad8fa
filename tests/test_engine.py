"""The host chunk engine behind the C ABI on the device (-m gpu): several device contexts in one
gts_ctx with in-order hand-back, the WARNING side channel (encodedSequence.go:39), write errors and
cancellation (writer.go:171-179, transeq.go:145-149), and the plain-C driver of the Go shim's call
sequence (tests/c/stream_drive.c).  A single-GPU box exercises the multi-GPU dealing by using
device 0 several times."""
import io
import os
import random
import subprocess
import threading

import pytest

import oracle
from fasta_gen import mid_fasta, nasty_fasta, regular_fasta

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def bulk(nbytes, seed):
    """README-shaped records (300..3000 nt), numpy speed."""
    import workloads
    return workloads.readme189(total_bytes=nbytes, seed=seed)


def _ndev():
    import torch
    return torch.cuda.device_count()


def _devsets():
    n = _ndev()
    sets = [[0], [0, 0], [0, 0, 0]]
    if n >= 2:
        sets.append(list(range(n)))
    return sets


def test_multi_device_contexts_match_the_oracle():
    import gotranseq_b200
    data = bulk(6_000_000, 31) + regular_fasta(32, 2, 200_000, 300_000) + regular_fasta(33, 500, 0, 50) + bulk(1_000_000, 34)
    for kw, okw in ((dict(Frame="6"), dict(frame="6")),
                    (dict(Frame="6", Table=11, Clean=True, Trim=True), dict(frame="6", table=11, clean=True, trim=True)),
                    (dict(Frame="R", Alternative=True), dict(frame="R", alternative=True))):
        want = oracle.translate(data, **okw)
        for devs in _devsets():
            for chunk in (1 << 16, 1 << 20):
                with gotranseq_b200.Translator(gotranseq_b200.Options(**kw), devices=devs, chunk_bytes=chunk) as t:
                    assert t.num_devices() == len(devs)
                    assert t.translate_bytes(data) == want, (kw, devs, chunk, "one-shot")
                    out = io.BytesIO()
                    t.translate_stream(io.BytesIO(data), out)
                    assert out.getvalue() == want, (kw, devs, chunk, "stream")
                    st = t.stats()
                    assert st["out_bytes"] == len(want) and st["in_bytes"] == len(data)
                    # and once more on the same context, one-shot after a stream
                    assert t.translate_bytes(data) == want


@pytest.mark.parametrize("seed", range(30))
def test_random_nasty_streams_on_three_contexts(seed):
    import gotranseq_b200
    rng = random.Random(seed)
    data = b"".join(nasty_fasta(seed * 7 + k, max_records=30, max_len=900) for k in range(6))
    kw = dict(frame=rng.choice(["1", "F", "-2", "R", "6", "6"]), table=rng.choice([0, 2, 11, 23]), clean=rng.random() < 0.5,
              alternative=rng.random() < 0.5, trim=rng.random() < 0.5)
    want = oracle.translate(data, **kw)
    O = gotranseq_b200.Options(Frame=kw["frame"], Table=kw["table"], Clean=kw["clean"], Alternative=kw["alternative"],
                               Trim=kw["trim"])
    with gotranseq_b200.Translator(O, devices=[0, 0, 0], chunk_bytes=rng.choice([4096, 9000, 1 << 15])) as t:
        out = io.BytesIO()
        t.translate_stream(io.BytesIO(data), out)
        assert out.getvalue() == want, (seed, kw, "stream")
        assert t.translate_bytes(data) == want, (seed, kw, "one-shot")


def test_blank_lines_in_front_of_the_first_chunk():
    """a first chunk that holds only blank lines must not emit the "always" record of
    transeq.go:213 -- that one belongs to an input that is empty altogether."""
    import gotranseq_b200
    body = regular_fasta(5, 200, 100, 400)
    for lead in (b"\n" * 9000, b"\r\n" * 5000, b"\n\r\n\n" * 3000):
        data = lead + body
        want = oracle.translate(data, frame="6")
        with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=[0, 0], chunk_bytes=4096) as t:
            out = io.BytesIO()
            t.translate_stream(io.BytesIO(data), out)
            assert out.getvalue() == want
            assert t.translate_bytes(data) == want
    # and the input that IS empty altogether (or blank only) still gets its one record
    for data in (b"", b"\n\n\n", b"\r\n"):
        want = oracle.translate(data, frame="6")
        with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=[0, 0], chunk_bytes=4096) as t:
            out = io.BytesIO()
            t.translate_stream(io.BytesIO(data), out)
            assert out.getvalue() == want == t.translate_bytes(data)


def test_many_chunks_through_the_stream():
    """Thousands of chunks, records straddling every cut, a record much larger than a chunk in the
    middle (the input buffer grows geometrically and every byte is scanned for a boundary once)."""
    import gotranseq_b200
    import workloads
    data = mid_fasta(3) * 3 + bulk(20_000_000, 40) + workloads.contigs(1, 6_000_000, 6_000_000, 41) + \
        regular_fasta(42, 3000, 10, 700)
    want = oracle.translate(data, frame="6", trim=True)
    with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6", Trim=True), devices=[0, 0], chunk_bytes=1 << 14) as t:
        out = io.BytesIO()
        t.translate_stream(io.BytesIO(data), out)
        assert out.getvalue() == want
        assert t.stats()["passes"] > 150


def _warning_cases(goldens):
    rng = random.Random(77)
    cases = [goldens["input"].encode()]  # sequence13 holds S and K (test.fna:45-47)
    alpha = b"ACGTacgtNnUu" * 6 + b"RYSWKMBDHV-*. \t1>\xc3\xff\x80\r"
    recs = []
    for r in range(400):
        n = rng.randint(0, 700)
        seq = bytes(rng.choice(alpha) for _ in range(n))
        w = rng.choice([60, 60, 7, 80, 200])
        hdr = rng.choice([b">r%d" % r, b">r%d with words" % r, b">", b">x\ttab %d" % r])
        recs.append(hdr + b"\n" + b"\n".join(seq[i:i + w] for i in range(0, n, w)) + b"\n")
    cases.append(b"ACGTXXACGT\nQQ\n" + b"".join(recs))          # bytes in front of the first header too
    cases.append(b"".join(recs).replace(b"\n", b"\r\n"))           # CRLF: the dropped CR is no nucleotide
    cases.append(b">only\n" + bytes(rng.choice(b"EFILPQ") for _ in range(50000)) + b"\n")  # every byte invalid
    return cases


def test_warning_side_channel_matches_the_reference_text(goldens):
    import gotranseq_b200
    for data in _warning_cases(goldens):
        want_out, want_warn = oracle.translate_with_warnings(data, frame="6")
        for devs, chunk in (([0], 1 << 20), ([0, 0, 0], 1 << 13)):
            got = []
            with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=devs, chunk_bytes=chunk,
                                           warnings=got.append) as t:
                assert t.translate_bytes(data) == want_out
                assert b"".join(got) == want_warn, ("one-shot", devs)
                del got[:]
                out = io.BytesIO()
                t.translate_stream(io.BytesIO(data), out)
                assert out.getvalue() == want_out
                assert b"".join(got) == want_warn, ("stream", devs)
    # without a sink nothing is reported and the output is the same
    assert gotranseq_b200.translate_bytes(goldens["input"].encode(), frame="6") == \
        oracle.translate(goldens["input"].encode(), frame="6")


class _FailingWriter:
    def __init__(self, fail_on):
        self.n, self.fail_on, self.parts = 0, fail_on, []

    def write(self, b):
        self.n += 1
        if self.n == self.fail_on:
            raise OSError("no space left on device")
        self.parts.append(bytes(b))


def test_write_error_cancels_the_stream_and_keeps_the_context_usable():
    """writer.go:171-179 / transeq.go:163-172: the first failed write is Translate's error."""
    import gotranseq_b200
    from gotranseq_b200 import _lib
    data = bulk(5_000_000, 50)
    want = oracle.translate(data, frame="6")
    for devs in ([0], [0, 0, 0]):
        with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=devs, chunk_bytes=1 << 16) as t:
            for fail_on in (1, 2, 7):
                w = _FailingWriter(fail_on)
                with pytest.raises(gotranseq_b200.TranseqError) as e:
                    t.translate_stream(io.BytesIO(data), w)
                assert e.value.code == _lib.GTS_ERR_WRITE
                assert str(e.value) == "fail to write to output file: no space left on device"
                assert b"".join(w.parts) == want[:sum(len(p) for p in w.parts)]  # what was written is a prefix
                out = io.BytesIO()
                t.translate_stream(io.BytesIO(data), out)  # the context is reusable
                assert out.getvalue() == want


def test_cancel_mid_stream():
    import gotranseq_b200
    from gotranseq_b200 import _lib
    data = bulk(8_000_000, 51)
    want = oracle.translate(data, frame="6")

    class Reader:
        def __init__(self, tr, after):
            self.tr, self.after, self.pos, self.calls = tr, after, 0, 0

        def read(self, n):
            self.calls += 1
            if self.calls == self.after:
                threading.Thread(target=self.tr.cancel).start()  # from another thread, like a Go ctx
            out = data[self.pos:self.pos + min(n, 50000)]
            self.pos += len(out)
            return out

    with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=[0, 0], chunk_bytes=1 << 16) as t:
        for after in (1, 5, 40):
            with pytest.raises(gotranseq_b200.TranseqError) as e:
                t.translate_stream(Reader(t, after), io.BytesIO())
            assert e.value.code == _lib.GTS_ERR_CANCELLED
            out = io.BytesIO()
            t.translate_stream(io.BytesIO(data), out)
            assert out.getvalue() == want
        # destroy with chunks queued and running must not hang
        buf_n = 0
        t2 = gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=[0, 0], chunk_bytes=1 << 16)
        import ctypes
        L, h = t2._L, t2._h
        b, c = ctypes.c_void_p(), ctypes.c_size_t()
        pos = 0
        while pos < len(data):
            assert L.gts_acquire_input(h, ctypes.byref(b), ctypes.byref(c)) == 0
            k = min(c.value, len(data) - pos)
            ctypes.memmove(b.value, data[pos:pos + k], k)
            pos += k
            assert L.gts_submit(h, k, 0) == 0
            buf_n += 1
            if buf_n == 6:
                break
        t2.close()


def _build_c(tmp_path, name, lib_dir=None, lib="gotranseq_b200"):
    exe = tmp_path / name
    lib_dir = lib_dir or os.path.join(ROOT, "gotranseq_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", name + ".c"), "-o", str(exe), "-L", lib_dir, "-l" + lib,
                    "-Wl,-rpath," + lib_dir], check=True)
    return str(exe)


def test_c_driver_of_the_shim_call_sequence(goldens, tmp_path):
    """tests/c/stream_drive.c: gts_acquire_input -> gts_submit -> gts_next_output -> gts_release_output
    from plain C on the device, the calls go/transeq/translate_b200.go makes."""
    _c_driver_cases(_build_c(tmp_path, "stream_drive"), goldens, tmp_path)


def _c_driver_cases(exe, goldens, tmp_path):
    cases = [(goldens["input"].encode(), dict(frame="6")),
             (bulk(4_000_000, 60), dict(frame="6", table=11, clean=True, trim=True)),
             (_warning_cases(goldens)[1], dict(frame="R", alternative=True))]
    for k, (data, kw) in enumerate(cases):
        src, dst = tmp_path / ("in%d.fna" % k), tmp_path / ("out%d.faa" % k)
        src.write_bytes(data)
        want_out, want_warn = oracle.translate_with_warnings(data, **kw)
        for ngpu, chunk, read_size, fail in ((0, 1 << 16, 0, None), (3, 1 << 14, 5000, None), (2, 1 << 15, 0, 0)):
            args = [exe, str(src), str(dst), kw["frame"], str(kw.get("table", 0)), str(int(kw.get("clean", False))),
                    str(int(kw.get("alternative", False))), str(int(kw.get("trim", False))), str(chunk), str(ngpu),
                    str(read_size)] + ([str(fail)] if fail is not None else [])
            r = subprocess.run(args, capture_output=True)
            assert r.returncode == 0, (k, ngpu, r.returncode, r.stderr)
            assert dst.read_bytes() == want_out, (k, ngpu)
            if fail is None:
                assert r.stdout == want_warn, (k, ngpu)


def test_cli_write_failure_and_warnings(goldens, tmp_path):
    """main.go:87-90: the error of the run is printed as "fail to translate file:\\n%v"; the WARNING
    lines of encodedSequence.go:39 go to stdout."""
    _cli_cases(os.path.join(ROOT, "gotranseq_b200", "bin", "gotranseq"), goldens, tmp_path)


def _cli_cases(exe, goldens, tmp_path):
    data = goldens["input"].encode()
    src = tmp_path / "t.fna"
    src.write_bytes(data)
    want_out, want_warn = oracle.translate_with_warnings(data, frame="6")
    r = subprocess.run([exe, "-s", str(src), "-o", str(tmp_path / "o.faa"), "-f", "6"], capture_output=True)
    assert r.returncode == 0 and (tmp_path / "o.faa").read_bytes() == want_out
    assert r.stdout == want_warn
    if os.path.exists("/dev/full"):
        r = subprocess.run([exe, "-s", str(src), "-o", "/dev/full", "-f", "6"], capture_output=True)
        assert r.returncode == 0
        assert r.stdout.endswith(b"fail to translate file:\nfail to write to output file: write /dev/full: no space left on device")


# ---- any_order: the reference's -n > 1 behaviour (writer.go:171-173: whoever is full writes) ----

def _entries(out):
    """The output as a sorted list of its per-frame entries (an entry starts with '>' at a line start)."""
    assert out[:1] in (b"", b">") and out[-1:] in (b"", b"\n")
    return sorted(p.rstrip(b"\n") for p in (b"\n" + out).split(b"\n>")[1:])


class _Blocks:
    def __init__(self):
        self.blocks = []

    def write(self, b):
        self.blocks.append(bytes(b))


def test_any_order_hands_back_every_record_exactly_once():
    import gotranseq_b200
    data = bulk(6_000_000, 51) + regular_fasta(52, 400, 10, 300) + regular_fasta(53, 3, 150_000, 200_000) + bulk(500_000, 54)
    for kw, okw in ((dict(Frame="6"), dict(frame="6")), (dict(Frame="R", Trim=True), dict(frame="R", trim=True))):
        want = oracle.translate(data, **okw)
        want_entries = _entries(want)
        for devs in _devsets():
            for chunk in (1 << 14, 1 << 18):
                with gotranseq_b200.Translator(gotranseq_b200.Options(**kw), devices=devs, chunk_bytes=chunk, any_order=True) as t:
                    w = _Blocks()
                    t.translate_stream(io.BytesIO(data), w)
                    got = b"".join(w.blocks)
                    assert len(got) == len(want), (kw, devs, chunk)
                    # every result is a whole number of records
                    assert all(b[:1] == b">" and b[-1:] == b"\n" for b in w.blocks), (kw, devs, chunk)
                    assert _entries(got) == want_entries, (kw, devs, chunk)
                    st = t.stats()
                    assert st["out_bytes"] == len(want) and st["in_bytes"] == len(data)
                    # the one-shot call is not affected: one buffer, input order
                    assert t.translate_bytes(data) == want
                    # and the context streams again
                    w2 = _Blocks()
                    t.translate_stream(io.BytesIO(data), w2)
                    assert _entries(b"".join(w2.blocks)) == want_entries


def test_any_order_never_writes_behind_an_over_long_line():
    """transeq.go:27,186: nothing behind a line of >= 100 MiB is ever written, in whatever order the
    chunks complete (a chunk may overtake only chunks too small to hold such a line)."""
    import gotranseq_b200
    cap = 20000
    rng = random.Random(9)
    seq = lambda n: "".join(rng.choice("ACGT") for _ in range(n))
    rec = lambda i, n, w=60: ">r%d\n" % i + "\n".join(seq(n)[j:j + w] for j in range(0, n, w)) + "\n"
    text = "".join(rec(i, 2000) for i in range(80)) + ">q\n" + seq(3 * cap) + "\n" + "".join(rec(100 + i, 1500) for i in range(80))
    data = text.encode()
    oracle.set_max_line_length(cap)
    try:
        want = oracle.translate(data, frame="6")
    finally:
        oracle.set_max_line_length(100 << 20)
    assert b">r99" not in want and b">r100" not in want and b">r79" in want
    for devs in ([0], [0, 0, 0]):
        with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=devs, chunk_bytes=4096, any_order=True) as t:
            t.set_max_line_length(cap)
            w = _Blocks()
            t.translate_stream(io.BytesIO(data), w)
            assert _entries(b"".join(w.blocks)) == _entries(want), devs
