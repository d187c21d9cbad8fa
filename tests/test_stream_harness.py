"""tools/stream_harness.cpp (bench infrastructure: the shim's call sequence in C++ for bench.py's streaming
legs) against a stand-in for the library made of ctypes callbacks: every byte is submitted once and in order
whatever the number of memcpy helpers, the writer thread sees every result, an error from the library ends
the run with that code and a gts_cancel.  No GPU involved."""
import ctypes
import os
import queue
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


class FakeEngine:
    """gts_acquire_input / gts_submit / gts_wait_output / gts_release_output / gts_cancel of an engine whose
    result for a chunk is the chunk itself."""

    def __init__(self, cap, fail_submit_at=-1):
        self.cap = cap
        self.bufs = [np.zeros(cap, dtype=np.uint8) for _ in range(3)]
        self.turn = 0
        self.q = queue.Queue()
        self.submitted = []
        self.seen = []
        self.held = None
        self.cancelled = 0
        self.nsubmit = 0
        self.fail_submit_at = fail_submit_at
        A = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t))
        S = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int)
        R = ctypes.CFUNCTYPE(ctypes.c_int, ctypes.c_void_p)
        self.cbs = [A(self.acquire), S(self.submit), A(self.wait_output), R(self.release), R(self.cancel)]

    def acquire(self, _ctx, buf, cap):
        b = self.bufs[self.turn % 3]
        buf[0] = b.ctypes.data
        cap[0] = self.cap
        return 0

    def submit(self, _ctx, n, eof):
        if self.nsubmit == self.fail_submit_at:
            return -7
        self.nsubmit += 1
        if n:
            data = self.bufs[self.turn % 3][:n].tobytes()
            self.turn += 1
            self.submitted.append(data)
            self.q.put(data)
        if eof:
            self.q.put(None)
        return 0

    def wait_output(self, _ctx, p, n):
        item = self.q.get()
        if item is None:
            n[0] = 0
            return 0
        self.held = ctypes.create_string_buffer(item, len(item))
        p[0] = ctypes.addressof(self.held)
        n[0] = len(item)
        self.seen.append(item)
        return 0

    def release(self, _ctx):
        self.held = None
        return 0

    def cancel(self, _ctx):
        self.cancelled += 1
        self.q.put(None)
        return 0


def _run(H, bench, eng, data, repeats, threads, read_size):
    api = bench._StreamApi(*[ctypes.cast(cb, ctypes.c_void_p).value for cb in eng.cbs])
    out = bench._StreamResult()
    arr = np.frombuffer(data, dtype=np.uint8)
    rc = H.stream_harness_run(ctypes.byref(api), None, ctypes.c_void_p(arr.ctypes.data), len(data), repeats, threads,
                              read_size, ctypes.byref(out))
    return rc, out


def test_stream_harness_submits_every_byte_once_and_in_order():
    import bench
    H = bench.stream_harness()
    if H is None:
        pytest.skip("no C++ compiler")
    rng = np.random.default_rng(3)
    data = rng.integers(0, 256, size=23_456_789, dtype=np.uint8).tobytes()
    for cap, threads, read_size, repeats in ((9 << 20, 4, 0, 2), (5 << 20, 1, 0, 1), (6 << 20, 3, 1 << 20, 1), (64 << 20, 8, 0, 3)):
        eng = FakeEngine(cap)
        rc, out = _run(H, bench, eng, data, repeats, threads, read_size)
        assert rc == 0 and out.rc == 0
        assert b"".join(eng.submitted) == data * repeats, (cap, threads, read_size)
        assert b"".join(eng.seen) == data * repeats
        assert out.out_bytes == len(data) * repeats and out.chunks == len(eng.submitted)
        limit = min(cap, read_size or cap)
        assert all(len(s) <= limit for s in eng.submitted)
        assert eng.cancelled == 0


def test_stream_harness_stops_at_a_library_error():
    import bench
    H = bench.stream_harness()
    if H is None:
        pytest.skip("no C++ compiler")
    data = bytes(range(256)) * 40000
    eng = FakeEngine(1 << 20, fail_submit_at=3)
    rc, out = _run(H, bench, eng, data, 1, 2, 0)
    assert rc == -7 and out.rc == -7
    assert eng.cancelled == 1 and len(eng.submitted) == 3
    assert out.out_bytes == sum(len(s) for s in eng.seen)
