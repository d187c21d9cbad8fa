"""The HOST logic of the library without a GPU: gotranseq_b200/csrc/gts_api.cu (chunk engine: record-boundary
cuts and carry-over, buffer pools, several device contexts, in-order and any-order hand-back, the WARNING side
channel, write errors, cancellation, the over-long line rule, the retries of chunks whose tables were too small,
the device-resident calls) compiled as plain C++ against a stand-in for the device -- tests/c/standin/: the
CUDA runtime calls carried out on the host and the device pass restated behind the kernels' launchers.  The
scenarios are the GPU engine tests themselves (tests/test_engine.py, imported by tests/standin_driver.py) plus a
few that need control over timing; each runs in a process of its own with GOTRANSEQ_B200_LIB pointing at the
stand-in build, against the oracle.

TEST INFRASTRUCTURE: the stand-in is built and loaded only here; the product has no CPU path
(tests/test_abi.py::test_no_gpu_fails_loudly), and what the kernels compute is checked on the GPU (-m gpu)."""
import os
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
STANDIN = os.path.join(ROOT, "tests", "c", "standin")
LIB = os.path.join(STANDIN, "libgotranseq_standin.so")
CSRC = os.path.join(ROOT, "gotranseq_b200", "csrc")
CUDA_INC = os.environ.get("CUDA_HOME", "/usr/local/cuda") + "/include"


@pytest.fixture(scope="module")
def standin_lib():
    srcs = [os.path.join(CSRC, "gts_api.cu"), os.path.join(STANDIN, "standin_cuda.cpp"), os.path.join(STANDIN, "standin_pass.cpp")]
    deps = srcs + [os.path.join(CSRC, f) for f in ("gts_device.h", "gts_kernels.h", "gts_codes.h", "gts_boundary.h")] + \
        [os.path.join(ROOT, "include", "gotranseq_b200.h")]
    if not os.path.exists(os.path.join(CUDA_INC, "cuda_runtime_api.h")):
        pytest.skip("no CUDA headers")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(d) for d in deps):
        # -Bsymbolic: the library's CUDA calls bind to its own stand-ins even when a real libcudart is loaded
        # into the process (torch)
        subprocess.run(["g++", "-std=c++17", "-O1", "-fPIC", "-shared", "-pthread", "-Wl,-Bsymbolic", "-I", CUDA_INC,
                        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", LIB, "-x", "c++"] + srcs, check=True)
    return LIB


def _run(lib, name, env=None, timeout=600):
    e = dict(os.environ, GOTRANSEQ_B200_LIB=lib)
    e.update(env or {})
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "standin_driver.py"), name], env=e, capture_output=True,
                       text=True, timeout=timeout)
    assert r.returncode == 0 and ("ok " + name) in r.stdout, (name, r.returncode, r.stdout[-2000:], r.stderr[-4000:])


@pytest.mark.parametrize("name", ["multi_device_contexts", "random_nasty_streams", "blank_lines_first", "many_chunks",
                                  "warnings", "write_error", "cancel", "any_order", "any_order_long_line",
                                  "over_long_lines", "overflow_retries", "device_resident", "fuzz_streams", "bench_stream_legs"])
def test_engine_on_the_stand_in_device(standin_lib, name):
    _run(standin_lib, name)


def test_any_order_behind_a_slow_large_chunk(standin_lib):
    """The deadlock an earlier build had: with any_order, chunks behind an unfinished chunk that can hold an over-long
    line took all the output buffers, could not be handed back, and the large chunk waited for a buffer for ever."""
    _run(standin_lib, "any_order_slow_large_chunk", env={"GTS_STANDIN_SLOW_BYTES": "20000", "GTS_STANDIN_SLOW_MS": "200"},
         timeout=120)


def test_engine_under_thread_sanitizer(tmp_path):
    """tests/c/engine_tsan.cpp: reader, writer and canceller threads of the caller against the engine's issuer /
    retirer threads, with -fsanitize=thread: no data race reported, results equal to the oracle's."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    import workloads
    from fasta_gen import nasty_fasta
    exe = tmp_path / "engine_tsan"
    r = subprocess.run(["g++", "-std=c++17", "-O1", "-g", "-fsanitize=thread", "-pthread", "-I", CUDA_INC,
                        "-I", os.path.join(ROOT, "include"), "-I", CSRC, "-o", str(exe),
                        os.path.join(ROOT, "tests", "c", "engine_tsan.cpp"), "-x", "c++", os.path.join(CSRC, "gts_api.cu"),
                        os.path.join(STANDIN, "standin_cuda.cpp"), os.path.join(STANDIN, "standin_pass.cpp")],
                       capture_output=True, text=True)
    if r.returncode != 0:
        pytest.skip("no ThreadSanitizer build here: " + r.stderr[-300:])
    data = workloads.readme189(total_bytes=1_500_000, seed=5) + b"".join(nasty_fasta(k, max_records=30, max_len=900) for k in range(12))
    (tmp_path / "in.fna").write_bytes(data)
    (tmp_path / "want.faa").write_bytes(oracle.translate(data, frame="6", trim=True))
    r = subprocess.run([str(exe), str(tmp_path / "in.fna"), str(tmp_path / "want.faa")], capture_output=True, text=True,
                       timeout=600, env=dict(os.environ, TSAN_OPTIONS="halt_on_error=0"))
    if "FATAL: ThreadSanitizer" in r.stderr:  # (the sanitizer's runtime cannot map its shadow memory on this kernel)
        pytest.skip(r.stderr.strip().splitlines()[0])
    assert r.returncode == 0 and "tsan driver ok" in r.stdout, (r.returncode, r.stdout[-500:], r.stderr[-3000:])
    assert "ThreadSanitizer" not in r.stderr, r.stderr[:6000]


@pytest.fixture(scope="module")
def goldens_json():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "emboss_goldens.json")) as f:
        return json.load(f)


def test_c_driver_of_the_shim_call_sequence_on_the_stand_in(standin_lib, goldens_json, tmp_path):
    """tests/c/stream_drive.c (the Go shim's call sequence from plain C) linked against the stand-in build: goldens,
    4 MB of records, the WARNING text, several device contexts, a failing write followed by gts_stream_reset."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import test_engine as E
    exe = E._build_c(tmp_path, "stream_drive", lib_dir=STANDIN, lib="gotranseq_standin")
    E._c_driver_cases(exe, goldens_json, tmp_path)


def test_cli_on_the_stand_in(standin_lib, goldens_json, tmp_path):
    """The `gotranseq` CLI (gotranseq_b200/host/main.cpp, transeq.cpp: flags of main.go:21-37, file -> file through
    TranslateFd with its reader / writer threads, messages of main.go:74,89) linked against the stand-in build: the
    reference's goldens through the flags, WARNING lines on stdout, a write failure on /dev/full."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import oracle
    import test_engine as E
    host = os.path.join(ROOT, "gotranseq_b200", "host")
    exe = str(tmp_path / "gotranseq")
    subprocess.run(["g++", "-O1", "-std=c++17", "-pthread", "-I", os.path.join(ROOT, "include"), "-o", exe,
                    os.path.join(host, "main.cpp"), os.path.join(host, "transeq.cpp"), "-L", STANDIN, "-lgotranseq_standin",
                    "-Wl,-rpath," + STANDIN], check=True)
    E._cli_cases(exe, goldens_json, tmp_path)
    src = tmp_path / "test.fna"
    src.write_bytes(goldens_json["input"].encode())
    want_warn = oracle.translate_with_warnings(goldens_json["input"].encode(), frame="1")[1]
    for case in goldens_json["cases"]:
        flags = ["-" + t if t.startswith("-") else t for t in case["options"].split()]  # -frame=6 -> --frame=6
        dst = tmp_path / "out.faa"
        for env in ({}, {"GOTRANSEQ_B200_GPUS": "2", "GOTRANSEQ_B200_CHUNK_MB": "0.004"}):
            r = subprocess.run([exe, "-s", str(src), "-o", str(dst)] + flags, capture_output=True, timeout=120,
                               env=dict(os.environ, **env))
            assert r.returncode == 0 and r.stdout == want_warn, (case["options"], r.stdout, r.stderr)
            assert dst.read_bytes().decode() == case["expected"], (case["options"], env)
    # a larger file through the reader / writer threads of TranslateFd, several device contexts, any order
    import workloads
    data = workloads.readme189(total_bytes=12_000_000, seed=8)
    big = tmp_path / "big.fna"
    big.write_bytes(data)
    want = oracle.translate(data, frame="6", table=11, clean=True, trim=True)
    for env in ({"GOTRANSEQ_B200_GPUS": "2", "GOTRANSEQ_B200_CHUNK_MB": "1"},
                {"GOTRANSEQ_B200_GPUS": "2", "GOTRANSEQ_B200_CHUNK_MB": "0.25", "GOTRANSEQ_B200_ANY_ORDER": "1"}):
        dst = tmp_path / "big.faa"
        r = subprocess.run([exe, "-s", str(big), "-o", str(dst), "-f", "6", "-t", "11", "-c", "-T"], capture_output=True,
                           timeout=300, env=dict(os.environ, **env))
        assert r.returncode == 0 and r.stdout == b"", (r.stdout[-300:], r.stderr[-300:])
        got = dst.read_bytes()
        if "GOTRANSEQ_B200_ANY_ORDER" in env:
            assert E._entries(got) == E._entries(want)
        else:
            assert got == want
