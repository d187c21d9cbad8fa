"""ctypes binding of the CPU oracle (oracle/transeq_oracle.c) -- TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this package; the product (gotranseq_b200/) never does.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libtranseq_oracle.so")
_CLI_PATH = os.path.join(_HERE, "transeq_oracle_cli")
_lib = None


def build(force=False):
    """Compile the oracle (shared library + CLI) with gcc via oracle/Makefile."""
    srcs = [os.path.join(_HERE, f) for f in ("transeq_oracle.c", "transeq_oracle.h", "transeq_oracle_cli.c")]
    stale = force or not (os.path.exists(_LIB_PATH) and os.path.exists(_CLI_PATH))
    if not stale:
        newest = max(os.path.getmtime(s) for s in srcs)
        stale = min(os.path.getmtime(_LIB_PATH), os.path.getmtime(_CLI_PATH)) < newest
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B", "all"], check=True, capture_output=True)
    return _LIB_PATH


class _Options(ctypes.Structure):
    _fields_ = [
        ("frame", ctypes.c_char_p),
        ("table", ctypes.c_int),
        ("clean", ctypes.c_int),
        ("alternative", ctypes.c_int),
        ("trim", ctypes.c_int),
        ("num_worker", ctypes.c_int),
    ]


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        L.to_translate.restype = ctypes.c_int
        L.to_translate.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(_Options),
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
            ctypes.POINTER(ctypes.c_uint64), ctypes.c_char_p, ctypes.c_size_t,
        ]
        L.to_translate_warn.restype = ctypes.c_int
        L.to_translate_warn.argtypes = [
            ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(_Options),
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t),
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t), ctypes.c_char_p, ctypes.c_size_t,
        ]
        L.to_free.argtypes = [ctypes.c_void_p]
        L.to_free.restype = None
        L.to_set_max_line_length.argtypes = [ctypes.c_size_t]
        L.to_set_max_line_length.restype = None
        L.to_create_code_array.restype = ctypes.c_int
        L.to_create_code_array.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.c_char_p, ctypes.c_size_t]
        L.to_load_table_code.restype = ctypes.c_int
        L.to_load_table_code.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        _lib = L
    return _lib


class OracleError(Exception):
    pass


def translate(data, frame="1", table=0, clean=False, alternative=False, trim=False, num_worker=1,
              return_warnings=False):
    """Oracle counterpart of transeq.Translate (transeq/transeq.go:114): bytes in, bytes out."""
    L = lib()
    if isinstance(data, (bytes, bytearray)):
        n = len(data)
        buf = (ctypes.c_char * max(n, 1)).from_buffer_copy(bytes(data) if n else b"\0")
        ptr = ctypes.cast(buf, ctypes.c_void_p)
    else:  # numpy uint8 array (no copy)
        n = int(data.size)
        buf = data
        ptr = ctypes.c_void_p(data.ctypes.data)
    opt = _Options(frame.encode(), int(table), int(clean), int(alternative), int(trim), int(num_worker))
    out = ctypes.c_void_p()
    out_n = ctypes.c_size_t()
    warns = ctypes.c_uint64()
    err = ctypes.create_string_buffer(256)
    rc = L.to_translate(ptr, n, ctypes.byref(opt), ctypes.byref(out), ctypes.byref(out_n),
                        ctypes.byref(warns), err, 256)
    del buf
    if rc != 0:
        raise OracleError(err.value.decode())
    try:
        # (ctypes.string_at takes the size as a C int: wrong beyond 2 GiB)
        res = bytes((ctypes.c_char * out_n.value).from_address(out.value)) if out_n.value else b""
    finally:
        L.to_free(out)
    return (res, warns.value) if return_warnings else res


def translate_with_warnings(data, frame="1", table=0, clean=False, alternative=False, trim=False):
    """(protein FASTA, stdout text of the WARNING lines) -- encodedSequence.go:39, num_worker 1."""
    L = lib()
    data = bytes(data)
    n = len(data)
    buf = (ctypes.c_char * max(n, 1)).from_buffer_copy(data if n else b"\0")
    opt = _Options(frame.encode(), int(table), int(clean), int(alternative), int(trim), 1)
    out, wout = ctypes.c_void_p(), ctypes.c_void_p()
    out_n, w_n = ctypes.c_size_t(), ctypes.c_size_t()
    err = ctypes.create_string_buffer(256)
    rc = L.to_translate_warn(ctypes.cast(buf, ctypes.c_void_p), n, ctypes.byref(opt), ctypes.byref(out),
                             ctypes.byref(out_n), ctypes.byref(wout), ctypes.byref(w_n), err, 256)
    if rc != 0:
        if wout.value:
            L.to_free(wout)
        raise OracleError(err.value.decode())
    try:
        res = bytes((ctypes.c_char * out_n.value).from_address(out.value)) if out_n.value else b""
        wtxt = bytes((ctypes.c_char * w_n.value).from_address(wout.value)) if w_n.value else b""
    finally:
        L.to_free(out)
        L.to_free(wout)
    return res, wtxt


def set_max_line_length(n):
    lib().to_set_max_line_length(n)


def code_array(table, clean=False):
    """The reference's 263173-byte codon array (transeq/transeq.go:30-86) as bytes."""
    L = lib()
    size = (4 | 4 << 8 | 4 << 16) + 1
    buf = ctypes.create_string_buffer(size)
    err = ctypes.create_string_buffer(256)
    if L.to_create_code_array(table, int(clean), buf, err, 256) != 0:
        raise OracleError(err.value.decode())
    return buf.raw


def table_code(table):
    """64 amino acids in T,C,A,G codon order (ncbicode/code.go:367-388)."""
    L = lib()
    buf = ctypes.create_string_buffer(64)
    err = ctypes.create_string_buffer(256)
    if L.to_load_table_code(table, buf, err, 256) != 0:
        raise OracleError(err.value.decode())
    return buf.raw.decode()


def cli_path():
    build()
    return _CLI_PATH
