/*
 * gotranseq_b200.h -- C ABI of the B200-native translation hot path.
 *
 * This is the drop-in boundary for gotranseq's `transeq.Translate`
 * (reference: transeq/transeq.go:114, `func Translate(inputSequence io.Reader, out io.Writer,
 * options Options) error`).  The reference has no FFI of its own (pure Go, CGO_ENABLED=0,
 * .goreleaser.yml:7); these are the entry points a cgo shim for that function binds
 * (see INTEGRATION.md for the Go side).  Plain pointers and sizes only.
 *
 * Everything between "bytes of a nucleotide FASTA" and "bytes of the protein FASTA" runs on
 * the GPU (sm_100a kernels in gotranseq_b200/csrc/).  There is no CPU fallback: if no CUDA
 * device is usable every entry point that needs one fails with GTS_ERR_CUDA.
 */
#ifndef GOTRANSEQ_B200_H
#define GOTRANSEQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GTS_VERSION "0.2.0"

/* Return codes. 0 = success; the text for the last failure is gts_last_error(). */
enum {
    GTS_OK = 0,
    GTS_ERR_FRAME = -1,    /* "wrong value for -f | --frame parameter: %s"  (transeq.go:107) */
    GTS_ERR_TABLE = -2,    /* "invalid table code: %v"                       (ncbicode/code.go:378) */
    GTS_ERR_CUDA = -3,     /* CUDA runtime / no device (no CPU fallback)                        */
    GTS_ERR_NOMEM = -4,    /* host or device allocation failed                                  */
    GTS_ERR_CAPACITY = -5, /* caller-provided output buffer too small; needed size reported     */
    GTS_ERR_ARG = -6,      /* NULL / misaligned / out-of-range argument, wrong call order       */
    GTS_ERR_WRITE = -7,    /* "fail to write to output file: %v"             (writer.go:175)    */
    GTS_ERR_CANCELLED = -8 /* gts_cancel was called (the reference's ctx cancel, transeq.go:130) */
};

/*
 * Mirrors `type Options struct` (transeq/options.go:4-11) field for field.
 * num_worker is accepted for source compatibility and ignored (the GPU path has no worker
 * goroutines; output is in input order, i.e. the reference's `-n 1` order, unless any_order is set).
 */
typedef struct gts_options {
    const char *frame;    /* Options.Frame: "1","2","3","F","-1","-2","-3","R","6"           */
    int32_t table;        /* Options.Table: NCBI genetic code id (ncbicode/code.go:341-363)   */
    uint8_t clean;        /* Options.Clean: stop codon '*' -> 'X'                             */
    uint8_t alternative;  /* Options.Alternative: frame -1 starts at the last codon           */
    uint8_t trim;         /* Options.Trim: strip trailing 'X' / '*' of every frame            */
    uint8_t reserved0;
    int32_t num_worker;   /* Options.NumWorker: ignored                                       */
    int32_t device;       /* CUDA device ordinal; -1 = the calling thread's current device    */
    uint64_t chunk_bytes; /* streaming chunk size (bytes of FASTA per GPU pass); 0 = default  */
    /* Record sharding over the GPUs of one box (replaces the worker pool of transeq.go:132-160:
     * record-aligned chunks of the byte stream are dealt to the GPUs in turn, every GPU runs its own
     * H2D / kernels / D2H pipeline, results are handed back strictly in input order; no collective).
     *   num_gpus == 0   one GPU: `device`
     *   num_gpus  > 0   device_ids[0 .. num_gpus)
     *   num_gpus  < 0   every visible GPU
     * The device-resident and gts_piece_* entry points use the first device only. */
    int32_t num_gpus;
    int32_t device_ids[8];
    /* any_order != 0: the streaming calls hand a chunk's result back as soon as it is complete, whatever
     * chunks before it are still on their way -- the reference's `-n > 1` behaviour, where every worker
     * writes its buffer when it is full (writer.go:171-173) and the order of the records in the output
     * is up to the scheduler.  Each result is still a whole number of records, every record appears
     * exactly once, and with one GPU the order stays the input order in practice.  A chunk never
     * overtakes one that could hold a line of >= 100 MiB (transeq.go:27,186: nothing behind such a line
     * may be written).  gts_translate_buffer is not affected (its result is one buffer in input order). */
    int32_t any_order;
} gts_options;

typedef struct gts_ctx gts_ctx;

/*
 * Replaces the option decoding at the top of Translate: computeFrames (transeq.go:88-110) and
 * createCodeArray (transeq.go:30-86, ncbicode.LoadTableCode code.go:367-388).  Validates
 * frame / table with the reference's error texts, builds the fused forward / reverse-
 * complement codon table, creates the CUDA streams and scratch buffers.
 * On failure *out is NULL and gts_last_error(NULL) holds the message (thread-local).
 */
int gts_create(const gts_options *opt, gts_ctx **out);
void gts_destroy(gts_ctx *ctx);

/* Message of the last error on ctx (or of the last failed gts_create when ctx == NULL). */
const char *gts_last_error(const gts_ctx *ctx);

/*
 * Option decoding only, no device needed (used by host-side checks and the CPU-only tests).
 * gts_frame_mask: bit k set = frame index k (0..5 = 1,2,3,-1,-2,-3) requested
 * (computeFrames, transeq.go:88-110); returns GTS_ERR_FRAME for an unknown name.
 * gts_codon_table: fills aa64 with the table's 64 amino acids in T,C,A,G codon order
 * (LoadTableCode, code.go:367-388); returns GTS_ERR_TABLE for an unknown id.
 * gts_fused_lut: the 125-entry device table for (table, clean): entry c0 + 5*c1 + 25*c2
 * (codes N=0,A=1,C=2,T=3,G=4) = forward amino acid | reverse-strand amino acid << 8.
 */
int gts_frame_mask(const char *frame, uint32_t *mask, int *reverse);
int gts_codon_table(int32_t table, char aa64[64]);
int gts_fused_lut(int32_t table, int clean, uint16_t lut[125]);

/*
 * Device-resident hot path: everything Translate does between in.Read and out.Write
 * (readSequenceFromFasta transeq.go:183-216, newEncodedSequence encodedSequence.go:17-43,
 * reverseComplement encodedSequence.go:71-97, writer.translate / translate3Frames / writeHeader
 * / writeAA / trimAndReturn writer.go:57-169), on a FASTA buffer already in device memory.
 *   d_in   device pointer to n bytes of FASTA (treated as one complete input: start of file
 *          to EOF); must be 16-byte aligned (cudaMalloc / torch allocations are).  The kernels
 *          read whole aligned 16-byte units: the unit that holds byte n-1 is read to its end, so
 *          the allocation must cover n rounded up to 16 (every CUDA allocation does)
 *   d_out  device pointer, capacity cap bytes; receives the protein FASTA
 *   out_n  receives the exact output size; if it exceeds cap nothing is written past cap and
 *          GTS_ERR_CAPACITY is returned (out_n is still set, so the caller can retry)
 *   stream a cudaStream_t (as void*; NULL = the legacy default stream).  All kernels are
 *          enqueued on it; the call returns after the stream has finished them.
 */
int gts_translate_device(gts_ctx *ctx, const void *d_in, size_t n, void *d_out, size_t cap,
                         size_t *out_n, void *stream);

/*
 * Same work, enqueue only (no host synchronisation): used to time the kernels with CUDA
 * events on `stream`.  gts_device_result() synchronises and reports the outcome of the last
 * enqueue (GTS_OK / GTS_ERR_CAPACITY / ...) and its output size.
 */
int gts_translate_device_async(gts_ctx *ctx, const void *d_in, size_t n, void *d_out, size_t cap,
                               void *stream);
int gts_device_result(gts_ctx *ctx, size_t *out_n);

/* Upper bound of the output size used by the library when it sizes its own buffers. */
size_t gts_output_bound_hint(const gts_ctx *ctx, size_t n);

/*
 * One-shot host call: host FASTA bytes in, host protein-FASTA bytes out (H2D, kernels, D2H).
 * *out points into a pinned buffer owned by ctx, valid until the next call on ctx.
 * The input is one complete FASTA (start of file to EOF).  It runs on the same chunk pipeline as
 * the streaming interface (record-aligned chunks dealt to the context's GPUs, H2D / kernels / D2H
 * of different chunks overlapped), with the chunks read straight from `in` (fastest when `in` is
 * pinned, gts_host_alloc) and copied back straight to their final place in *out.
 */
int gts_translate_buffer(gts_ctx *ctx, const uint8_t *in, size_t n, const uint8_t **out,
                         size_t *out_n);

/*
 * Streaming host interface (what the Go shim's Translate loop drives; replaces the reader
 * goroutine + channel + worker flushes of transeq.go:126-163 and writer.go:171-181):
 *
 *   for {
 *       gts_acquire_input(ctx, &buf, &cap);        // pinned, C-owned
 *       n = in.Read(buf[:cap]);
 *       gts_submit(ctx, n, eof);                   // async H2D + kernels + D2H
 *       while (gts_next_output(ctx, &p, &m) == GTS_OK && m > 0) {
 *           out.Write(p[:m]); gts_release_output(ctx);
 *       }
 *       if eof { break }
 *   }
 *
 * The library cuts the byte stream at record boundaries ("\n>"), carries the unfinished last
 * record over to the next pass and hands results back strictly in input order (unless
 * gts_options.any_order is set: then the oldest COMPLETE chunk comes back).  gts_submit only
 * queues the pass: a worker thread owned by ctx runs it (H2D, kernels, D2H) while the caller reads
 * the next chunk into the other pinned input buffer and writes the previous result, so reading,
 * translating and writing overlap.  Before eof gts_next_output does not block: n == 0 means
 * "nothing finished yet, go on reading".  After gts_submit(..., eof=1) it waits for each remaining
 * result; n == 0 then means "done" (and ctx is ready for another stream).  Drain gts_next_output
 * after every gts_submit, as in the loop above (the library keeps a bounded number of chunks in
 * flight: three per GPU).  The other entry points of ctx must not be called while a stream is in
 * progress.  With several GPUs (gts_options.num_gpus) the chunks are dealt to them in turn.
 */
int gts_acquire_input(gts_ctx *ctx, uint8_t **buf, size_t *cap);
int gts_submit(gts_ctx *ctx, size_t nbytes, int eof);
int gts_next_output(gts_ctx *ctx, const uint8_t **buf, size_t *n);
int gts_release_output(gts_ctx *ctx);
/* gts_next_output that always waits: for a writer thread of its own (one thread may drive
 * gts_acquire_input / gts_submit while another drives gts_wait_output / gts_release_output).
 * n == 0: the stream has ended and everything was handed out. */
int gts_wait_output(gts_ctx *ctx, const uint8_t **buf, size_t *n);

/*
 * Cancellation and write errors (transeq.go:126-131,145-149,163-172; writer.go:171-179).  In the
 * reference the first failed out.Write puts "fail to write to output file: %v" on the error
 * channel and cancels the context; the reader and the other workers stop, Translate returns that
 * error.  Here:
 *   gts_report_write_error  the caller's write failed: every further call on the stream returns
 *                           GTS_ERR_WRITE and gts_last_error() is "fail to write to output file: <msg>"
 *   gts_cancel              stop the stream (safe from any thread): queued chunks are dropped,
 *                           blocked calls return GTS_ERR_CANCELLED
 *   gts_stream_reset        wait for the GPUs to go idle, drop everything queued, clear the error:
 *                           the context is ready for a new stream (also needed after any other
 *                           stream error)
 */
int gts_report_write_error(gts_ctx *ctx, const char *msg);
int gts_cancel(gts_ctx *ctx);
int gts_stream_reset(gts_ctx *ctx);

/* GPUs the context deals chunks to. */
int gts_num_devices(const gts_ctx *ctx);

/*
 * WARNING side channel (encodedSequence.go:36-40): the reference prints one line per nucleotide
 * that is none of ACGTUN (either case) and translates it as N.  With a callback installed the
 * kernels record those bytes and the library reports them in input order (with any_order: chunk by chunk,
 * in the order the chunks are handed back), on the caller's thread,
 * from gts_next_output / gts_wait_output / gts_translate_buffer before the chunk's output is handed
 * over.  header = the record's header line (without the line break; header_len == 0 for sequence
 * bytes in front of the first header line), pos = index of the nucleotide in its record.
 * gts_format_warning writes exactly the bytes the reference prints to stdout --
 *   "WARNING: invalid char in sequence %s: '%s' ( pos %d), replacing with 'N'\n"
 * where the first %s is s[:headerSize]: the four raw little-endian bytes of header_len + 4 followed
 * by the header line, and the second is Go's string(byte): the UTF-8 encoding of that code point --
 * into dst (capacity cap) and returns the length needed.  The device-resident entry points do not
 * report warnings.
 */
typedef void (*gts_warning_fn)(void *user, const uint8_t *header, size_t header_len, uint8_t byte, uint64_t pos);
int gts_set_warning_callback(gts_ctx *ctx, gts_warning_fn fn, void *user);
size_t gts_format_warning(const uint8_t *header, size_t header_len, uint8_t byte, uint64_t pos, char *dst, size_t cap);

/* Counters of the last device pass / since creation (diagnostics and bench accounting). */
typedef struct gts_stats {
    uint64_t records;        /* records emitted by the last pass                          */
    uint64_t nucleotides;    /* nucleotides translated by the last pass                   */
    uint64_t in_bytes;       /* FASTA bytes of the last pass                              */
    uint64_t out_bytes;      /* protein FASTA bytes of the last pass                      */
    uint64_t kernel_launches;/* kernels launched since gts_create                         */
    uint64_t h2d_bytes;      /* bytes copied host->device since gts_create                */
    uint64_t d2h_bytes;      /* bytes copied device->host since gts_create                */
    uint64_t passes;         /* device passes since gts_create                            */
} gts_stats;
int gts_get_stats(const gts_ctx *ctx, gts_stats *st);

/*
 * Per-kernel timing (bench.py's roofline numbers).  While enabled, CUDA events are recorded on
 * the pass's stream around each kernel.  gts_kernel_times waits for the recorded passes, adds
 * their durations up in ms -- [0] scratch reset + k_parse, [1] k_tile_scan + k_records + k_size
 * (+ k_warn), [2] k_lines, [3] what is left of k_headers and k_short (they run on side streams next
 * to k_lines) once k_lines has finished -- reports how many passes that covers and starts a new
 * accumulation.
 */
int gts_set_profiling(gts_ctx *ctx, int enable);
int gts_kernel_times(gts_ctx *ctx, double ms[4], uint64_t *passes);

/*
 * The reference reads lines through a bufio.Scanner limited to 100 MiB (transeq.go:27,186): at a
 * line of that length or more it silently stops reading and translates what it has seen so far.
 * The library reproduces this; this hook lowers the limit (>= 16384) so that tests can reach it,
 * like the oracle's to_set_max_line_length.
 */
int gts_set_max_line_length(gts_ctx *ctx, size_t n);

/*
 * One record split over several GPUs (BASELINE config 4: a chromosome-sized record; the reference
 * translates a record on ONE worker, writer.go:57-117, so this has no counterpart there).  The
 * caller cuts the record's FASTA bytes into pieces at line starts (the first piece begins with the
 * header line) and gives each piece to one GPU / context:
 *
 *   1. gts_piece_scan*       parses the piece: its nucleotide count, the codes of its first 192
 *                            and last two nucleotides, the header line's length (first piece)
 *   2. the caller adds the counts up (an all-gather of ~200 bytes between the ranks) and fills
 *      gts_piece for every piece: nucleotides before it, nucleotides of the whole record, the two
 *      nucleotides in front of it (carry), the 192 behind it (next_head), the header length
 *   3. gts_piece_translate*  writes the piece's share of the record's protein FASTA at its
 *                            FINAL offsets: at most 12 byte ranges (one per requested frame, plus
 *                            the header lines on the first piece), returned as segments
 *
 * The pieces' segments are disjoint and cover the record's output exactly.  Output lines are
 * never shared: a 60-residue line (180 nucleotides) is computed by the piece that holds the last
 * nucleotide of the line's lowest codon; carry and next_head supply what the line needs from the
 * neighbouring pieces, so no sequence bytes travel besides those ~200.
 *
 * With --trim (writer.go:141-163) the frame lengths depend on the residues at the record's ends:
 * between steps 2 and 3 the pieces walk them.  Start with len = gts_frame_lengths(nt_total) and
 * call gts_piece_trim* on the pieces until every requested frame is resolved: a piece checks the
 * residues whose codon starts in it, from len[f] - 1 backwards while they are 'X' or '*', and
 * reports where it stopped (resolved[f] = 1) or that the walk leaves the piece (resolved[f] = 0,
 * len[f] = what is left for the neighbouring piece).  Forward frames end in the last piece, reverse
 * frames in the first, so one call on each usually settles everything; the result goes into
 * gts_piece.trim_len (has_trim_len = 1) of every piece.  gts_piece_translate* of a buffer that was
 * scanned by the call before reuses that parse (no second k_parse).
 */
typedef struct gts_piece_info {
    uint64_t nucleotides; /* nucleotides in the piece                                          */
    uint64_t headers;     /* header lines seen: must be 1 for the first piece, 0 for the others */
    uint32_t header_len;  /* bytes of the header line, CR dropped (first piece)                */
    uint8_t tail[2];      /* codes (N=0 A=1 C=2 T=3 G=4) of its last / second-to-last nucleotide */
    uint8_t reserved[2];
    uint8_t head[192];    /* codes of its first 192 nucleotides (0 beyond the piece's end)      */
} gts_piece_info;

typedef struct gts_piece {
    uint64_t nt_before;   /* nucleotides of the record in the pieces before this one            */
    uint64_t nt_total;    /* nucleotides of the whole record                                    */
    uint32_t header_len;  /* gts_piece_info.header_len of the first piece                       */
    uint8_t carry[2];     /* codes of the two nucleotides before the piece: [0] last, [1] second to last */
    uint8_t first;        /* this piece starts with the record's header line                    */
    uint8_t last;         /* this piece ends the record                                         */
    uint8_t next_head[192]; /* codes of the 192 nucleotides behind the piece (0 beyond the record) */
    uint8_t has_trim_len; /* --trim: trim_len holds the record's trimmed frame lengths           */
    uint8_t reserved[7];
    uint64_t trim_len[6]; /* residues of frames 1,2,3,-1,-2,-3 after --trim (gts_piece_trim)      */
} gts_piece;

typedef struct gts_segment {
    uint64_t offset;      /* offset in the record's protein FASTA                               */
    uint64_t length;
} gts_segment;
#define GTS_MAX_SEGMENTS 12

/* Device-resident: d_in as for gts_translate_device.  d_out holds the WHOLE record's output
 * (*out_total bytes, reported also on GTS_ERR_CAPACITY); the piece writes its segments only. */
int gts_frame_lengths(const gts_ctx *ctx, uint64_t n, uint64_t len[6]);
int gts_piece_scan_device(gts_ctx *ctx, const void *d_in, size_t n, int first, gts_piece_info *info,
                          void *stream);
int gts_piece_trim_device(gts_ctx *ctx, const void *d_in, size_t n, const gts_piece *piece, uint64_t len[6],
                          uint8_t resolved[6], void *stream);
int gts_piece_translate_device(gts_ctx *ctx, const void *d_in, size_t n, const gts_piece *piece,
                               void *d_out, size_t cap, size_t *out_total,
                               gts_segment segs[GTS_MAX_SEGMENTS], int *nseg, void *stream);
/*
 * The same step without a host round trip between scan and translate (what bench.py's
 * chromosome_split runs): gts_piece_scan_async enqueues the scan and leaves the piece's 256-byte
 * summary in device memory (d_info); the caller all-gathers the summaries of all ranks on the same
 * stream (NCCL) into d_infos (world x 256 bytes, rank order); gts_piece_translate_async plans the
 * piece on the device from them and enqueues the translate kernels on the scan's slots;
 * gts_piece_result waits, checks, and returns the piece's segments (and, optionally, its gts_piece).
 * No --trim here (GTS_ERR_ARG): the walk over the pieces needs the host-driven calls.
 */
#define GTS_PIECE_INFO_BYTES 256
int gts_piece_scan_async(gts_ctx *ctx, const void *d_in, size_t n, void *d_info, void *stream);
int gts_piece_translate_async(gts_ctx *ctx, const void *d_in, size_t n, int rank, int world, const void *d_infos,
                              void *d_out, size_t cap, void *stream);
int gts_piece_result(gts_ctx *ctx, size_t *out_total, gts_segment segs[GTS_MAX_SEGMENTS], int *nseg,
                     gts_piece *piece_out);
/* Host buffers: gts_piece_scan copies the piece to the device and keeps it there for
 * gts_piece_translate, which hands segment i back as data[i] (pinned, owned by ctx, valid until
 * the next call on ctx). */
int gts_piece_scan(gts_ctx *ctx, const uint8_t *in, size_t n, int first, gts_piece_info *info);
int gts_piece_trim(gts_ctx *ctx, const gts_piece *piece, uint64_t len[6], uint8_t resolved[6]);
int gts_piece_translate(gts_ctx *ctx, const gts_piece *piece, size_t *out_total,
                        gts_segment segs[GTS_MAX_SEGMENTS], const uint8_t *data[GTS_MAX_SEGMENTS],
                        int *nseg);

/* Pinned host memory for callers of gts_translate_buffer that want full-speed PCIe copies. */
void *gts_host_alloc(size_t bytes);
void gts_host_free(void *p);

#ifdef __cplusplus
}
#endif
#endif /* GOTRANSEQ_B200_H */
