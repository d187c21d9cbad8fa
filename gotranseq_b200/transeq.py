"""Python mirror of the reference's `transeq` package over the C ABI.

  Options    field for field `type Options struct` (transeq/options.go:4-11)
  Translate  `func Translate(inputSequence io.Reader, out io.Writer, options Options) error`
             (transeq/transeq.go:114): errors are raised as TranseqError with the reference's texts
"""
import ctypes
import sys
from dataclasses import dataclass

from . import _lib


def _bytes_at(ptr, n):
    """n bytes at address ptr as a bytes object (ctypes.string_at takes the size as a C int: wrong beyond 2 GiB)."""
    return bytes((ctypes.c_char * n).from_address(ptr)) if n else b""


class TranseqError(Exception):
    """Mirrors the `error` returned by transeq.Translate; .code is the C ABI return code."""

    def __init__(self, code, message):
        super().__init__(message)
        self.code = code


@dataclass
class Options:
    """transeq/options.go:4-11 (same names, same defaults as the go-flags tags)."""
    Frame: str = "1"
    Table: int = 0
    Clean: bool = False
    Alternative: bool = False
    Trim: bool = False
    NumWorker: int = 0  # accepted and ignored: output is always in input order (the -n 1 order)


class Translator:
    """One gts_ctx: a decoded option set plus its device scratch.  Not thread safe."""

    def __init__(self, options=None, device=-1, chunk_bytes=0, devices=None, warnings=None, any_order=False, **kw):
        """devices: None = one GPU (`device`); "all" = every visible GPU; a list of ordinals = those GPUs
        (record-aligned chunks are dealt to them in turn, results come back in input order).
        warnings: a callable(bytes) receiving each WARNING line of encodedSequence.go:39 exactly as the
        reference prints it to stdout (None = invalid nucleotides are not reported).
        any_order: the streaming calls hand results back as they complete (the reference's `-n > 1` order:
        whole records, each exactly once, in no particular order)."""
        if options is None:
            options = Options(**kw)
        self.options = options
        L = _lib.lib()
        o = _lib.gts_options()
        if devices == "all":
            o.num_gpus = -1
        elif devices:
            o.num_gpus = len(devices)
            for i, d in enumerate(devices):
                o.device_ids[i] = int(d)
        self._frame = options.Frame.encode()
        o.frame = self._frame
        o.table = int(options.Table)
        o.clean = 1 if options.Clean else 0
        o.alternative = 1 if options.Alternative else 0
        o.trim = 1 if options.Trim else 0
        o.num_worker = int(options.NumWorker)
        o.device = int(device)
        o.chunk_bytes = int(chunk_bytes)
        o.any_order = 1 if any_order else 0
        h = ctypes.c_void_p()
        rc = L.gts_create(ctypes.byref(o), ctypes.byref(h))
        if rc != _lib.GTS_OK:
            raise TranseqError(rc, L.gts_last_error(None).decode())
        self._h = h
        self._L = L
        self._warn_cb = None
        if warnings is not None:
            self.set_warning_sink(warnings)

    def set_warning_sink(self, sink):
        """sink(bytes) is called with each WARNING line (formatted by gts_format_warning), in input order."""
        L = self._L

        def cb(_user, header, header_len, byte, pos):
            need = L.gts_format_warning(header, header_len, byte, pos, None, 0)
            buf = ctypes.create_string_buffer(need)
            L.gts_format_warning(header, header_len, byte, pos, buf, need)
            sink(buf.raw[:need])

        self._warn_cb = _lib.gts_warning_fn(cb) if sink is not None else ctypes.cast(None, _lib.gts_warning_fn)
        self._check(L.gts_set_warning_callback(self._h, self._warn_cb, None))

    def num_devices(self):
        return self._L.gts_num_devices(self._h)

    def cancel(self):
        """transeq.go:130: stop the stream from any thread; blocked calls raise TranseqError(GTS_ERR_CANCELLED)."""
        self._check(self._L.gts_cancel(self._h))

    def reset(self):
        """Drop whatever is queued and clear the stream's error: ready for a new stream."""
        self._check(self._L.gts_stream_reset(self._h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.gts_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != _lib.GTS_OK:
            raise TranseqError(rc, self._L.gts_last_error(self._h).decode())

    # ---- one-shot host call ----
    def translate_bytes(self, data):
        n = len(data)
        buf = (ctypes.c_char * max(n, 1)).from_buffer_copy(bytes(data) if n else b"\0")
        out = ctypes.c_void_p()
        out_n = ctypes.c_size_t()
        self._check(self._L.gts_translate_buffer(self._h, ctypes.cast(buf, ctypes.c_void_p), n,
                                                 ctypes.byref(out), ctypes.byref(out_n)))
        return _bytes_at(out.value, out_n.value)

    def translate_host_ptr(self, ptr, n):
        """Host pointer in, (pointer, size) of the pinned result out (valid until the next call)."""
        out = ctypes.c_void_p()
        out_n = ctypes.c_size_t()
        self._check(self._L.gts_translate_buffer(self._h, ctypes.c_void_p(ptr), n, ctypes.byref(out), ctypes.byref(out_n)))
        return out.value, out_n.value

    # ---- device-resident ----
    def translate_device(self, d_in, n, d_out, cap, stream=0):
        out_n = ctypes.c_size_t()
        rc = self._L.gts_translate_device(self._h, ctypes.c_void_p(d_in), n, ctypes.c_void_p(d_out), cap,
                                          ctypes.byref(out_n), ctypes.c_void_p(stream))
        if rc == _lib.GTS_ERR_CAPACITY:
            raise TranseqError(rc, f"output buffer too small: need {out_n.value} bytes")
        self._check(rc)
        return out_n.value

    def translate_device_async(self, d_in, n, d_out, cap, stream=0):
        self._check(self._L.gts_translate_device_async(self._h, ctypes.c_void_p(d_in), n, ctypes.c_void_p(d_out), cap,
                                                       ctypes.c_void_p(stream)))

    def device_result(self):
        out_n = ctypes.c_size_t()
        self._check(self._L.gts_device_result(self._h, ctypes.byref(out_n)))
        return out_n.value

    def stats(self):
        st = _lib.gts_stats()
        self._check(self._L.gts_get_stats(self._h, ctypes.byref(st)))
        return {n: getattr(st, n) for n, _ in st._fields_}

    def set_max_line_length(self, n):
        """Test hook: the reference's bufio.Scanner token limit (transeq.go:27), default 100 MiB."""
        self._check(self._L.gts_set_max_line_length(self._h, n))

    def set_profiling(self, enable):
        self._check(self._L.gts_set_profiling(self._h, 1 if enable else 0))

    def kernel_times(self):
        """(ms[4] summed over the passes recorded since the last call, number of passes)."""
        ms = (ctypes.c_double * 4)()
        n = ctypes.c_uint64()
        self._check(self._L.gts_kernel_times(self._h, ms, ctypes.byref(n)))
        return list(ms), n.value

    # ---- one record split over several GPUs (gotranseq_b200/split.py drives these) ----
    def piece_scan(self, data, first):
        """Parse a piece of a split record (host bytes; they stay on the device for piece_translate).
        Returns dict(nucleotides, headers, header_len, tail)."""
        n = len(data)
        buf = (ctypes.c_char * max(n, 1)).from_buffer_copy(bytes(data) if n else b"\0")
        info = _lib.gts_piece_info()
        self._check(self._L.gts_piece_scan(self._h, ctypes.cast(buf, ctypes.c_void_p), n, 1 if first else 0,
                                           ctypes.byref(info)))
        return dict(nucleotides=info.nucleotides, headers=info.headers, header_len=info.header_len,
                    tail=(info.tail[0], info.tail[1]), head=bytes(info.head))

    @staticmethod
    def _piece_struct(piece):
        p = _lib.gts_piece()
        p.nt_before, p.nt_total, p.header_len = piece["nt_before"], piece["nt_total"], piece["header_len"]
        p.carry[0], p.carry[1] = piece["carry"]
        p.first, p.last = (1 if piece["first"] else 0), (1 if piece["last"] else 0)
        nh = bytes(piece.get("next_head", b""))[:192]
        ctypes.memmove(p.next_head, nh + bytes(192 - len(nh)), 192)
        if piece.get("trim_len") is not None:
            p.has_trim_len = 1
            for f in range(6):
                p.trim_len[f] = int(piece["trim_len"][f])
        return p

    def frame_lengths(self, n):
        """Untrimmed residues of the six frames of a record with n nucleotides (0 for frames not requested)."""
        out = (ctypes.c_uint64 * 6)()
        self._check(self._L.gts_frame_lengths(self._h, n, out))
        return list(out)

    def piece_trim(self, piece, lens):
        """This piece's part of the --trim walk (the piece last scanned by piece_scan): (lens, resolved)."""
        p = self._piece_struct(piece)
        ln = (ctypes.c_uint64 * 6)(*[int(x) for x in lens])
        res = (ctypes.c_uint8 * 6)()
        self._check(self._L.gts_piece_trim(self._h, ctypes.byref(p), ln, res))
        return list(ln), [bool(x) for x in res]

    def piece_trim_device(self, d_in, n, piece, lens, stream=0):
        p = self._piece_struct(piece)
        ln = (ctypes.c_uint64 * 6)(*[int(x) for x in lens])
        res = (ctypes.c_uint8 * 6)()
        self._check(self._L.gts_piece_trim_device(self._h, ctypes.c_void_p(d_in), n, ctypes.byref(p), ln, res,
                                                  ctypes.c_void_p(stream)))
        return list(ln), [bool(x) for x in res]

    def piece_translate(self, piece):
        """Translate the piece last scanned.  Returns (size of the whole record's output,
        [(offset, bytes), ...]): the byte ranges of that output this piece produces."""
        p = self._piece_struct(piece)
        segs = (_lib.gts_segment * _lib.GTS_MAX_SEGMENTS)()
        data = (ctypes.c_void_p * _lib.GTS_MAX_SEGMENTS)()
        nseg = ctypes.c_int()
        total = ctypes.c_size_t()
        self._check(self._L.gts_piece_translate(self._h, ctypes.byref(p), ctypes.byref(total), segs, data,
                                                ctypes.byref(nseg)))
        return total.value, [(segs[i].offset, _bytes_at(data[i], segs[i].length)) for i in range(nseg.value)]

    def piece_scan_device(self, d_in, n, first, stream=0):
        info = _lib.gts_piece_info()
        self._check(self._L.gts_piece_scan_device(self._h, ctypes.c_void_p(d_in), n, 1 if first else 0,
                                                  ctypes.byref(info), ctypes.c_void_p(stream)))
        return dict(nucleotides=info.nucleotides, headers=info.headers, header_len=info.header_len,
                    tail=(info.tail[0], info.tail[1]), head=bytes(info.head))

    def piece_translate_device(self, d_in, n, piece, d_out, cap, stream=0):
        """Device-resident piece: writes its segments into d_out (the whole record's output buffer).
        Returns (size of the whole output, [(offset, length), ...])."""
        p = self._piece_struct(piece)
        segs = (_lib.gts_segment * _lib.GTS_MAX_SEGMENTS)()
        nseg = ctypes.c_int()
        total = ctypes.c_size_t()
        rc = self._L.gts_piece_translate_device(self._h, ctypes.c_void_p(d_in), n, ctypes.byref(p),
                                                ctypes.c_void_p(d_out), cap, ctypes.byref(total), segs,
                                                ctypes.byref(nseg), ctypes.c_void_p(stream))
        if rc == _lib.GTS_ERR_CAPACITY:
            raise TranseqError(rc, f"output buffer too small: need {total.value} bytes")
        self._check(rc)
        return total.value, [(segs[i].offset, segs[i].length) for i in range(nseg.value)]

    # ---- the split step planned on the device: scan -> all-gather (caller, NCCL) -> translate ----
    def piece_scan_async(self, d_in, n, d_info, stream=0):
        """Enqueue the scan of a piece; its 256-byte summary lands at device address d_info."""
        self._check(self._L.gts_piece_scan_async(self._h, ctypes.c_void_p(d_in), n, ctypes.c_void_p(d_info),
                                                 ctypes.c_void_p(stream)))

    def piece_translate_async(self, d_in, n, rank, world, d_infos, d_out, cap, stream=0):
        """Enqueue the translate pass of the piece just scanned, planned on the device from the
        all-gathered summaries at d_infos (world x 256 bytes, rank order)."""
        self._check(self._L.gts_piece_translate_async(self._h, ctypes.c_void_p(d_in), n, rank, world, ctypes.c_void_p(d_infos),
                                                      ctypes.c_void_p(d_out), cap, ctypes.c_void_p(stream)))

    def piece_result(self):
        """Wait for the enqueued piece: (size of the whole record's output, [(offset, length), ...], piece dict)."""
        segs = (_lib.gts_segment * _lib.GTS_MAX_SEGMENTS)()
        nseg = ctypes.c_int()
        total = ctypes.c_size_t()
        p = _lib.gts_piece()
        rc = self._L.gts_piece_result(self._h, ctypes.byref(total), segs, ctypes.byref(nseg), ctypes.byref(p))
        if rc == _lib.GTS_ERR_CAPACITY:
            raise TranseqError(rc, f"output buffer too small: need {total.value} bytes")
        self._check(rc)
        return (total.value, [(segs[i].offset, segs[i].length) for i in range(nseg.value)],
                dict(nt_before=p.nt_before, nt_total=p.nt_total, header_len=p.header_len, first=bool(p.first), last=bool(p.last)))

    # ---- streaming (what the Go shim's Translate loop does) ----
    def translate_stream(self, reader, writer):
        L, h = self._L, self._h
        buf = ctypes.c_void_p()
        cap = ctypes.c_size_t()
        obuf = ctypes.c_void_p()
        on = ctypes.c_size_t()
        eof = False
        try:
            while not eof:
                self._check(L.gts_acquire_input(h, ctypes.byref(buf), ctypes.byref(cap)))
                view = (ctypes.c_char * cap.value).from_address(buf.value)
                got = reader.readinto(view) if hasattr(reader, "readinto") else None
                if got is None:
                    chunk = reader.read(cap.value)
                    got = len(chunk)
                    ctypes.memmove(buf.value, chunk, got)
                eof = got == 0
                self._check(L.gts_submit(h, got, 1 if eof else 0))
                while True:
                    self._check(L.gts_next_output(h, ctypes.byref(obuf), ctypes.byref(on)))
                    if on.value == 0:
                        break
                    try:
                        writer.write(_bytes_at(obuf.value, on.value))
                    except Exception as e:
                        # writer.go:171-179: the first failed write becomes the error of Translate and
                        # cancels everything that is still queued
                        L.gts_report_write_error(h, str(e).encode())
                        self._check(L.gts_next_output(h, ctypes.byref(obuf), ctypes.byref(on)))  # raises GTS_ERR_WRITE
                    self._check(L.gts_release_output(h))
        except TranseqError:
            L.gts_stream_reset(h)  # the context stays usable (a Go caller gets the error and a clean shim)
            raise


def _stdout_sink(line):
    sys.stdout.buffer.write(line)


def Translate(inputSequence, out, options, devices="all", warnings=_stdout_sink, **kw):
    """transeq.Translate (transeq/transeq.go:114): read FASTA from `inputSequence`, write the
    protein FASTA to `out`; WARNING lines for invalid nucleotides go to stdout like the reference's
    (encodedSequence.go:39).  Raises TranseqError instead of returning an error.  Uses every visible
    GPU (the reference uses Options.NumWorker goroutines; here NumWorker is ignored)."""
    with Translator(options, devices=devices, warnings=warnings, **kw) as t:
        t.translate_stream(inputSequence, out)


def translate_bytes(data, frame="1", table=0, clean=False, alternative=False, trim=False, **kw):
    with Translator(Options(Frame=frame, Table=table, Clean=clean, Alternative=alternative, Trim=trim), **kw) as t:
        return t.translate_bytes(data)
