"""Runs host-logic scenarios of the library on the stand-in device (tests/c/standin/: the CUDA runtime calls and
the device pass restated on the host, gts_api.cu compiled as plain C++).  Started by tests/test_engine_standin.py
in a process of its own with GOTRANSEQ_B200_LIB pointing at the stand-in build; prints "ok <name>" per scenario.
TEST INFRASTRUCTURE: the product has no CPU path and never loads this library."""
import io
import json
import os
import random
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import gotranseq_b200  # noqa: E402
import oracle  # noqa: E402
from gotranseq_b200 import _lib  # noqa: E402

assert "standin" in _lib.LIB_PATH, "the stand-in driver must not run on the product library"

import test_engine as E  # noqa: E402  (the GPU engine tests: the same functions, on the stand-in)
import test_gpu_parity as P  # noqa: E402


def goldens():
    with open(os.path.join(ROOT, "tests", "golden", "emboss_goldens.json")) as f:
        return json.load(f)


class SmallReader:
    """At most k bytes per read: every read of a few KB becomes a chunk of its own."""

    def __init__(self, data, k):
        self.d, self.p, self.k = data, 0, k

    def readinto(self, view):
        n = min(self.k, len(view), len(self.d) - self.p)
        view[:n] = self.d[self.p:self.p + n]
        self.p += n
        return n


def any_order_behind_a_slow_large_chunk():
    """A chunk large enough to hold an over-long line, slow on its device, with dozens of small chunks completing
    behind it on the other device contexts: nothing behind it may be handed back, so nothing behind it may hold an
    output buffer either (the large chunk would wait for one for ever).  GTS_STANDIN_SLOW_* make the timing."""
    assert os.environ.get("GTS_STANDIN_SLOW_BYTES") == "20000"
    cap = 20000
    rng = random.Random(9)
    seq = lambda n: "".join(rng.choice("ACGT") for _ in range(n))


    def rec(i, n, w=60):
        s = seq(n)
        return ">r%d\n" % i + "\n".join(s[j:j + w] for j in range(0, n, w)) + "\n"

    # (1) the large chunk holds an over-long line: everything behind it is dropped
    # (2) it does not (one record of many ordinary lines): everything behind it is written
    texts = ["".join(rec(i, 2000) for i in range(40)) + ">q\n" + seq(3 * cap) + "\n" + "".join(rec(100 + i, 1500) for i in range(80)),
             "".join(rec(i, 2000) for i in range(40)) + rec(999, 3 * cap) + "".join(rec(100 + i, 1500) for i in range(80))]
    for text in texts:
        data = text.encode()
        oracle.set_max_line_length(cap)
        try:
            want = oracle.translate(data, frame="6")
        finally:
            oracle.set_max_line_length(100 << 20)
        for devs in ([0], [0, 0, 0], [0, 1, 0, 1]):
            for any_order in (True, False):
                with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=devs, chunk_bytes=4096,
                                               any_order=any_order) as t:
                    t.set_max_line_length(cap)
                    w = E._Blocks()
                    t.translate_stream(SmallReader(data, 5000), w)
                    got = b"".join(w.blocks)
                    assert t.stats()["passes"] >= 20 and len(w.blocks) >= 15
                    if any_order:
                        assert E._entries(got) == E._entries(want), devs
                    else:
                        assert got == want, devs


def overflow_retries():
    """Chunks whose record table / output estimate / warning list are too small are redone by the retirer: reads
    of a few nucleotides with long header lines (more than 1 record per 128 bytes, more than 4 x the input in
    output), every nucleotide invalid."""
    rng = random.Random(4)
    recs = []
    for i in range(60000):
        recs.append(b">read%06d some description that is repeated in every frame\n" % i + bytes(rng.choice(b"ACGTRYKM") for _ in range(rng.randint(1, 9))) + b"\n")
    tiny = b"".join(b">%d\nA\n" % (i % 10) for i in range(200000))
    for data in (b"".join(recs), tiny):
        want, want_warn = oracle.translate_with_warnings(data, frame="6")
        for devs, chunk in (([0], 1 << 20), ([0, 1, 0], 1 << 16)):
            got = []
            with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=devs, chunk_bytes=chunk, warnings=got.append) as t:
                out = io.BytesIO()
                t.translate_stream(io.BytesIO(data), out)
                assert out.getvalue() == want
                assert b"".join(got) == want_warn
                del got[:]
                assert t.translate_bytes(data) == want
                assert b"".join(got) == want_warn


def device_resident_calls():
    """gts_translate_device (+ _async / gts_device_result) with "device" pointers (host memory on the stand-in)."""
    import ctypes
    import numpy as np
    data = E.bulk(3_000_000, 77)
    want = oracle.translate(data, frame="6", trim=True)
    arr = np.frombuffer(data, dtype=np.uint8).copy()
    out = np.zeros(len(data) * 4 + 4096, dtype=np.uint8)
    with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6", Trim=True)) as t:
        n = t.translate_device(arr.ctypes.data, len(data), out.ctypes.data, out.size, 0)
        assert out[:n].tobytes() == want
        out[:] = 0
        t.translate_device_async(arr.ctypes.data, len(data), out.ctypes.data, out.size, 0)
        assert t.device_result() == len(want) and out[:len(want)].tobytes() == want
        small = np.zeros(1000, dtype=np.uint8)  # an output buffer that is too small is an error, not a crash
        try:
            t.translate_device(arr.ctypes.data, len(data), small.ctypes.data, small.size, 0)
            raise AssertionError("expected an error")
        except gotranseq_b200.TranseqError:
            pass


class RaggedReader:
    """Reads of random sizes (1 .. kmax bytes)."""

    def __init__(self, data, rng, kmax):
        self.d, self.p, self.rng, self.kmax = data, 0, rng, kmax

    def readinto(self, view):
        n = min(self.rng.randint(1, self.kmax), len(view), len(self.d) - self.p)
        view[:n] = self.d[self.p:self.p + n]
        self.p += n
        return n


def fuzz_streams(seed0=1000, iters=80):
    """Random FASTA (adversarial records, blank runs, CRLF, no final newline, a record larger than the chunks) x random
    options x device contexts x chunk sizes x read sizes x in-order / any-order x with / without the warning callback,
    streamed and one-shot, against the oracle."""
    from fasta_gen import nasty_fasta, regular_fasta
    for seed in range(seed0, seed0 + iters):
        rng = random.Random(seed)
        parts = []
        for k in range(rng.randint(1, 6)):
            c = rng.random()
            if c < 0.6:
                parts.append(nasty_fasta(seed * 13 + k, max_records=rng.randint(1, 40), max_len=rng.choice([5, 60, 900, 5000])))
            elif c < 0.8:
                parts.append(regular_fasta(seed * 13 + k, rng.randint(1, 200), 0, rng.choice([10, 300, 3000])))
            elif c < 0.9:
                parts.append(b"\n" * rng.randint(0, 5000))
            else:
                parts.append(b">big\n" + bytes(rng.choice(b"ACGTN") for _ in range(rng.randint(1000, 30000))) + b"\n")
        data = b"".join(parts)
        if rng.random() < 0.15:
            data = data.replace(b"\n", b"\r\n")
        if rng.random() < 0.1:
            data = data.rstrip(b"\n")
        kw = dict(frame=rng.choice(["1", "2", "3", "F", "-1", "-2", "-3", "R", "6", "6"]), table=rng.choice([0, 2, 11, 23, 30]),
                  clean=rng.random() < 0.5, alternative=rng.random() < 0.5, trim=rng.random() < 0.5)
        want, want_warn = oracle.translate_with_warnings(data, **kw)
        O = gotranseq_b200.Options(Frame=kw["frame"], Table=kw["table"], Clean=kw["clean"], Alternative=kw["alternative"], Trim=kw["trim"])
        devs = rng.choice([[0], [0, 0], [0, 1, 0], [1, 0, 1, 0, 1]])
        chunk = rng.choice([4096, 5000, 9000, 1 << 14, 1 << 16, 1 << 20])
        any_order = rng.random() < 0.3
        warn = rng.random() < 0.5
        got_w = []
        what = (seed, kw, devs, chunk, any_order, warn)
        with gotranseq_b200.Translator(O, devices=devs, chunk_bytes=chunk, any_order=any_order, warnings=(got_w.append if warn else None)) as t:
            w = E._Blocks()
            rd = RaggedReader(data, rng, rng.choice([1, 7, 100, 5000, 100000])) if rng.random() < 0.7 else io.BytesIO(data)
            t.translate_stream(rd, w)
            got = b"".join(w.blocks)
            if any_order:  # whole records in any order: the same lines, the same size (headers may be empty or repeated)
                assert len(got) == len(want) and sorted(got.split(b"\n")) == sorted(want.split(b"\n")), what
                assert not warn or sorted(b"".join(got_w).split(b"\n")) == sorted(want_warn.split(b"\n")), what
            else:
                assert got == want, what
                assert not warn or b"".join(got_w) == want_warn, what
            del got_w[:]
            assert t.translate_bytes(data) == want, what
            assert not warn or b"".join(got_w) == want_warn, what


def bench_stream_legs():
    """bench.py's streaming legs (tools/stream_harness.cpp and the Python loop) through the real engine."""
    import hashlib
    import numpy as np
    import bench
    import workloads
    data = workloads.readme189(total_bytes=12_000_000, seed=3)
    arr = np.frombuffer(data, dtype=np.uint8).copy()
    want = oracle.translate(data, frame="6")
    with gotranseq_b200.Translator(gotranseq_b200.Options(Frame="6"), devices=[0, 1], chunk_bytes=1 << 20) as tr:
        for threads in (1, 4):
            r = bench.stream_once(tr, arr.ctypes.data, len(data), reader_threads=threads)
            assert r["out"] == len(want) and r["chunks"] >= 8, r
        r = bench.stream_once(tr, arr.ctypes.data, len(data), repeats=3, reader_threads=4)
        assert r["out"] == 3 * len(want), r
        h = hashlib.sha256()
        r = bench.stream_once(tr, arr.ctypes.data, len(data), hasher=h, read_size=300_000)
        assert r["driver"] == "python" and h.hexdigest() == hashlib.sha256(want).hexdigest()


SCENARIOS = {
    "multi_device_contexts": E.test_multi_device_contexts_match_the_oracle,
    "random_nasty_streams": lambda: [E.test_random_nasty_streams_on_three_contexts(s) for s in range(30)],
    "blank_lines_first": E.test_blank_lines_in_front_of_the_first_chunk,
    "many_chunks": E.test_many_chunks_through_the_stream,
    "warnings": lambda: E.test_warning_side_channel_matches_the_reference_text(goldens()),
    "write_error": E.test_write_error_cancels_the_stream_and_keeps_the_context_usable,
    "cancel": E.test_cancel_mid_stream,
    "any_order": E.test_any_order_hands_back_every_record_exactly_once,
    "any_order_long_line": E.test_any_order_never_writes_behind_an_over_long_line,
    "any_order_slow_large_chunk": any_order_behind_a_slow_large_chunk,
    "over_long_lines": P.test_over_long_lines,
    "overflow_retries": overflow_retries,
    "device_resident": device_resident_calls,
    "fuzz_streams": fuzz_streams,
    "bench_stream_legs": bench_stream_legs,
}

if __name__ == "__main__":
    for name in sys.argv[1:]:
        t0 = time.time()
        SCENARIOS[name]()
        print("ok %s %.1fs" % (name, time.time() - t0), flush=True)
