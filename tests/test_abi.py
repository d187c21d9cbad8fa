"""The C-ABI shared library loads without a GPU and exports every symbol include/gotranseq_b200.h
declares; option decoding (no device needed) matches the reference's tables."""
import ctypes
import os
import re

import pytest

import oracle

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    with open(os.path.join(ROOT, "include", "gotranseq_b200.h")) as f:
        src = f.read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gts_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_the_whole_header():
    from gotranseq_b200 import _lib
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(L, n), f"{n} declared in gotranseq_b200.h but not exported"
    assert set(names) == set(_lib.EXPORTS), "ctypes binding and header disagree"


def test_frame_masks_match_compute_frames():
    # transeq.go:88-110
    from gotranseq_b200 import _lib
    L = _lib.lib()
    want = {"1": 1, "2": 2, "3": 4, "F": 7, "-1": 8, "-2": 16, "-3": 32, "R": 56, "6": 63}
    for name, m in want.items():
        mask = ctypes.c_uint32()
        rev = ctypes.c_int()
        assert L.gts_frame_mask(name.encode(), ctypes.byref(mask), ctypes.byref(rev)) == 0
        assert mask.value == m and rev.value == int(m >= 8)
    for bad in ("", "7", "f", "-4", "66"):
        assert L.gts_frame_mask(bad.encode(), None, None) == _lib.GTS_ERR_FRAME


def test_codon_tables_match_the_oracle():
    # ncbicode/code.go:367-388 (oracle: independent transcription as standard + diffs)
    from gotranseq_b200 import _lib
    L = _lib.lib()
    from fasta_gen import ALL_TABLES
    for t in ALL_TABLES:
        buf = ctypes.create_string_buffer(64)
        assert L.gts_codon_table(t, buf) == 0
        assert buf.raw.decode() == oracle.table_code(t), t
    assert L.gts_codon_table(1, ctypes.create_string_buffer(64)) == _lib.GTS_ERR_TABLE
    assert L.gts_codon_table(31, ctypes.create_string_buffer(64)) == _lib.GTS_ERR_TABLE


def test_fused_lut_matches_the_reference_code_array():
    # transeq.go:30-86: codes[n1 | n2<<8 | n3<<16]; reverse strand = complement, reversed (encodedSequence.go:71-97)
    from gotranseq_b200 import _lib
    L = _lib.lib()
    comp = lambda c: 0 if c == 0 else ((c - 1) ^ 2) + 1
    for t in (0, 2, 11, 23, 30):
        for clean in (0, 1):
            ca = oracle.code_array(t, bool(clean))
            lut = (ctypes.c_uint16 * 125)()
            assert L.gts_fused_lut(t, clean, lut) == 0
            for c0 in range(5):
                for c1 in range(5):
                    for c2 in range(5):
                        v = lut[c0 + 5 * c1 + 25 * c2]
                        assert (v & 0xFF) == ca[c0 | c1 << 8 | c2 << 16]
                        assert (v >> 8) == ca[comp(c2) | comp(c1) << 8 | comp(c0) << 16]


def test_all_codon_tables_match_the_reference_source():
    """All 21 effective codon maps (standard + every diff, U read as T) as derived from the reference's
    ncbicode/code.go by tests/golden/make_ncbi_tables.py: the library's tables and the oracle's
    independent transcription must both equal them.  When the checkout is mounted the fixture itself
    is re-derived and compared, so it cannot go stale."""
    import json
    from gotranseq_b200 import _lib
    from fasta_gen import ALL_TABLES
    L = _lib.lib()
    with open(os.path.join(ROOT, "tests", "golden", "ncbi_tables.json")) as f:
        fixture = json.load(f)
    assert sorted(int(k) for k in fixture) == sorted(ALL_TABLES) and len(fixture) == 21
    for t in ALL_TABLES:
        buf = ctypes.create_string_buffer(64)
        assert L.gts_codon_table(t, buf) == 0
        assert buf.raw.decode() == fixture[str(t)], t
        assert oracle.table_code(t) == fixture[str(t)], t
    path = "/root/reference/ncbicode/code.go"
    if os.path.exists(path):
        import importlib.util
        spec = importlib.util.spec_from_file_location("make_ncbi_tables", os.path.join(ROOT, "tests", "golden", "make_ncbi_tables.py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        assert mod.parse(open(path).read()) == fixture


def test_no_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    import gotranseq_b200
    with pytest.raises(gotranseq_b200.TranseqError, match="no CPU fallback"):
        gotranseq_b200.translate_bytes(b">a\nACG\n")


def test_header_compiles_as_c_and_links(tmp_path):
    """include/gotranseq_b200.h from plain C (what a cgo shim compiles), linked against the library."""
    import subprocess
    exe = tmp_path / "abi_smoke"
    lib_dir = os.path.join(ROOT, "gotranseq_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_smoke.c"), "-o", str(exe), "-L", lib_dir, "-lgotranseq_b200",
                    "-Wl,-rpath," + lib_dir], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout, r.stderr)
    assert r.stdout.startswith("abi ok")
    # the ctypes mirror of gts_options has the C layout (any_order: the last field)
    from gotranseq_b200 import _lib
    words = r.stdout.split()
    assert int(words[words.index("options") + 1]) == ctypes.sizeof(_lib.gts_options)
    assert int(words[words.index("any_order") + 1]) == _lib.gts_options.any_order.offset


def test_record_boundary_search_of_the_chunk_engine(tmp_path):
    """gts_boundary.h (where the engine may cut the stream, transeq.go:198) against a byte loop, on a
    96 MB buffer: the single-threaded and the multi-threaded paths of both directions."""
    import subprocess
    exe = tmp_path / "boundary_test"
    subprocess.run(["g++", "-O2", "-std=c++17", "-pthread", "-Wall", "-Wextra", os.path.join(ROOT, "tests", "c", "boundary_test.cpp"),
                    "-o", str(exe)], check=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout.startswith("ok"), (r.returncode, r.stdout, r.stderr)


def test_cli_reports_option_errors_like_the_reference(tmp_path):
    """main.go prints the error of transeq.Translate and returns (exit status 0); the option errors
    come before any device is touched, so this runs without a GPU."""
    import subprocess
    exe = os.path.join(ROOT, "gotranseq_b200", "bin", "gotranseq")
    src = tmp_path / "t.fna"
    src.write_bytes(b">a\nACG\n")
    for flags, text in ((["-f", "7"], "wrong value for -f | --frame parameter: 7"),
                        (["-f", "6", "-t", "1"], "invalid table code: 1")):
        r = subprocess.run([exe, "-s", str(src), "-o", str(tmp_path / "o.faa")] + flags, capture_output=True, text=True)
        assert text in r.stdout + r.stderr, (flags, r.stdout, r.stderr)
    # main.go:74: flag errors -> "wrong arguments: ..., try gotranseq --help for more informations", exit 1
    r = subprocess.run([exe, "-s", str(src), "-o", str(tmp_path / "o.faa"), "-t", "abc"], capture_output=True, text=True)
    assert r.returncode == 1 and r.stdout.startswith("wrong arguments: ") and \
        r.stdout.endswith(", try gotranseq --help for more informations\n")
    r = subprocess.run([exe, "--nosuchflag"], capture_output=True, text=True)
    assert r.returncode == 1 and "unknown flag `nosuchflag'" in r.stdout
    # main.go:89: run errors -> "fail to translate file:\n%v" (no trailing newline), exit 0
    r = subprocess.run([exe, "-s", str(src), "-o", str(tmp_path / "o.faa"), "-f", "7"], capture_output=True, text=True)
    assert r.returncode == 0 and r.stdout == "fail to translate file:\nwrong value for -f | --frame parameter: 7"
    r = subprocess.run([exe, "-o", str(tmp_path / "o.faa")], capture_output=True, text=True)
    assert r.stdout == "fail to translate file:\nmissing required parameter -s | -sequence, try gotranseq --help for details"
    r = subprocess.run([exe, "-s", str(tmp_path / "absent.fna"), "-o", str(tmp_path / "o.faa")], capture_output=True, text=True)
    assert r.stdout == "fail to translate file:\nopen %s: no such file or directory" % (tmp_path / "absent.fna")
    r = subprocess.run([exe, "--help"], capture_output=True, text=True)
    for flag in ("--sequence", "--outseq", "--frame", "--table", "--clean", "--alternative", "--trim", "--numcpu"):
        assert flag in r.stdout + r.stderr, flag
