"""ctypes binding of the C ABI (include/gotranseq_b200.h).  The shared library is built in-tree
by `make -C gotranseq_b200/csrc` (or __graft_entry__.build()); importing this module fails loudly
when it is missing -- there is no CPU fallback."""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
# GOTRANSEQ_B200_LIB: another build of the same library (kernel tuning sweeps)
LIB_PATH = os.environ.get("GOTRANSEQ_B200_LIB") or os.path.join(_HERE, "libgotranseq_b200.so")

GTS_OK = 0
GTS_ERR_FRAME = -1
GTS_ERR_TABLE = -2
GTS_ERR_CUDA = -3
GTS_ERR_NOMEM = -4
GTS_ERR_CAPACITY = -5
GTS_ERR_ARG = -6
GTS_ERR_WRITE = -7
GTS_ERR_CANCELLED = -8


class gts_options(ctypes.Structure):
    _fields_ = [
        ("frame", ctypes.c_char_p),
        ("table", ctypes.c_int32),
        ("clean", ctypes.c_uint8),
        ("alternative", ctypes.c_uint8),
        ("trim", ctypes.c_uint8),
        ("reserved0", ctypes.c_uint8),
        ("num_worker", ctypes.c_int32),
        ("device", ctypes.c_int32),
        ("chunk_bytes", ctypes.c_uint64),
        ("num_gpus", ctypes.c_int32),
        ("device_ids", ctypes.c_int32 * 8),
        ("any_order", ctypes.c_int32),
    ]


# typedef void (*gts_warning_fn)(void *user, const uint8_t *header, size_t header_len, uint8_t byte, uint64_t pos)
gts_warning_fn = ctypes.CFUNCTYPE(None, ctypes.c_void_p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t,
                                  ctypes.c_uint8, ctypes.c_uint64)


class gts_stats(ctypes.Structure):
    _fields_ = [(n, ctypes.c_uint64) for n in (
        "records", "nucleotides", "in_bytes", "out_bytes", "kernel_launches", "h2d_bytes", "d2h_bytes", "passes")]


class gts_piece_info(ctypes.Structure):
    _fields_ = [("nucleotides", ctypes.c_uint64), ("headers", ctypes.c_uint64), ("header_len", ctypes.c_uint32),
                ("tail", ctypes.c_uint8 * 2), ("reserved", ctypes.c_uint8 * 2), ("head", ctypes.c_uint8 * 192)]


class gts_piece(ctypes.Structure):
    _fields_ = [("nt_before", ctypes.c_uint64), ("nt_total", ctypes.c_uint64), ("header_len", ctypes.c_uint32),
                ("carry", ctypes.c_uint8 * 2), ("first", ctypes.c_uint8), ("last", ctypes.c_uint8),
                ("next_head", ctypes.c_uint8 * 192), ("has_trim_len", ctypes.c_uint8), ("reserved", ctypes.c_uint8 * 7),
                ("trim_len", ctypes.c_uint64 * 6)]


class gts_segment(ctypes.Structure):
    _fields_ = [("offset", ctypes.c_uint64), ("length", ctypes.c_uint64)]


GTS_MAX_SEGMENTS = 12
GTS_PIECE_INFO_BYTES = 256

EXPORTS = {
    # name: (restype, argtypes)
    "gts_create": (ctypes.c_int, [ctypes.POINTER(gts_options), ctypes.POINTER(ctypes.c_void_p)]),
    "gts_destroy": (None, [ctypes.c_void_p]),
    "gts_last_error": (ctypes.c_char_p, [ctypes.c_void_p]),
    "gts_frame_mask": (ctypes.c_int, [ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint32), ctypes.POINTER(ctypes.c_int)]),
    "gts_codon_table": (ctypes.c_int, [ctypes.c_int32, ctypes.c_char_p]),
    "gts_fused_lut": (ctypes.c_int, [ctypes.c_int32, ctypes.c_int, ctypes.POINTER(ctypes.c_uint16)]),
    "gts_translate_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                            ctypes.c_size_t, ctypes.POINTER(ctypes.c_size_t), ctypes.c_void_p]),
    "gts_translate_device_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p,
                                                  ctypes.c_size_t, ctypes.c_void_p]),
    "gts_device_result": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t)]),
    "gts_output_bound_hint": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_size_t]),
    "gts_translate_buffer": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "gts_acquire_input": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "gts_submit": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]),
    "gts_next_output": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "gts_release_output": (ctypes.c_int, [ctypes.c_void_p]),
    "gts_wait_output": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_size_t)]),
    "gts_report_write_error": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_char_p]),
    "gts_cancel": (ctypes.c_int, [ctypes.c_void_p]),
    "gts_stream_reset": (ctypes.c_int, [ctypes.c_void_p]),
    "gts_num_devices": (ctypes.c_int, [ctypes.c_void_p]),
    "gts_set_warning_callback": (ctypes.c_int, [ctypes.c_void_p, gts_warning_fn, ctypes.c_void_p]),
    "gts_format_warning": (ctypes.c_size_t, [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_uint8, ctypes.c_uint64,
                                             ctypes.c_void_p, ctypes.c_size_t]),
    "gts_get_stats": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(gts_stats)]),
    "gts_set_profiling": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_int]),
    "gts_kernel_times": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_uint64)]),
    "gts_set_max_line_length": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_size_t]),
    "gts_piece_scan_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                             ctypes.POINTER(gts_piece_info), ctypes.c_void_p]),
    "gts_piece_translate_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(gts_piece), ctypes.c_void_p, ctypes.c_size_t,
                                                  ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(gts_segment),
                                                  ctypes.POINTER(ctypes.c_int), ctypes.c_void_p]),
    "gts_frame_lengths": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_uint64, ctypes.POINTER(ctypes.c_uint64)]),
    "gts_piece_trim_device": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.POINTER(gts_piece),
                                             ctypes.POINTER(ctypes.c_uint64), ctypes.POINTER(ctypes.c_uint8), ctypes.c_void_p]),
    "gts_piece_trim": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(gts_piece), ctypes.POINTER(ctypes.c_uint64),
                                      ctypes.POINTER(ctypes.c_uint8)]),
    "gts_piece_scan_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p]),
    "gts_piece_translate_async": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int, ctypes.c_int,
                                                 ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p]),
    "gts_piece_result": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(gts_segment),
                                        ctypes.POINTER(ctypes.c_int), ctypes.POINTER(gts_piece)]),
    "gts_piece_scan": (ctypes.c_int, [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int,
                                      ctypes.POINTER(gts_piece_info)]),
    "gts_piece_translate": (ctypes.c_int, [ctypes.c_void_p, ctypes.POINTER(gts_piece), ctypes.POINTER(ctypes.c_size_t),
                                           ctypes.POINTER(gts_segment), ctypes.POINTER(ctypes.c_void_p),
                                           ctypes.POINTER(ctypes.c_int)]),
    "gts_host_alloc": (ctypes.c_void_p, [ctypes.c_size_t]),
    "gts_host_free": (None, [ctypes.c_void_p]),
}

_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: build it with `make -C gotranseq_b200/csrc` "
                "(gotranseq_b200 has no CPU fallback)")
        L = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in EXPORTS.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib
