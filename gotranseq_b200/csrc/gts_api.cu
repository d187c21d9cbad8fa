// gts_api.cu -- the C ABI of include/gotranseq_b200.h: option decoding, per-GPU scratch, device
// passes, and the host chunk engine behind the one-shot and the streaming entry points.
//
// This file replaces the orchestration of transeq.Translate (transeq/transeq.go:114-173).  Where
// the reference runs one reader goroutine feeding NumWorker translate goroutines over a channel
// (transeq.go:126-160) and lets every worker flush to the writer (writer.go:171-181), a gts_ctx
//   * cuts the byte stream into chunks at record boundaries ('>' at a line start, transeq.go:198),
//   * deals the chunks to its GPUs in turn; every GPU has an issuer thread (H2D copy + the six
//     kernels of gts_kernels.h, enqueued without waiting) and a retirer thread (exact output size,
//     D2H copy), so that the copies of neighbouring chunks overlap the kernels of the current one,
//   * hands the results back strictly in input order (the reference's -n 1 order),
//   * carries the reference's failure semantics: the first write error cancels everything
//     (writer.go:171-179, transeq.go:145-149,200-204).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <functional>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/gotranseq_b200.h"
#include "gts_boundary.h"
#include "gts_codes.h"
#include "gts_device.h"
#include "gts_kernels.h"

using gts::u32;
using gts::u64;

static thread_local std::string g_create_error;

namespace {

constexpr int kSlots = 8;  // chunk slots per GPU (device input / output buffers allocated on first use)
// slots in use: one chunk in H2D, one in the kernels, up to two in (or queued for) D2H (GTS_SLOTS: A/B runs)
static const int g_slots = [] { const char *e = std::getenv("GTS_SLOTS"); const int v = e ? std::atoi(e) : 4; return v < 2 ? 2 : (v > kSlots ? kSlots : v); }();

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct HostBuf {  // pinned host memory of the chunk pools
    uint8_t *p = nullptr;
    size_t cap = 0;
};

struct WarnOut {  // one WARNING line of encodedSequence.go:39, ready for the callback
    std::string header;
    uint64_t pos = 0;
    uint8_t byte = 0;
};

struct XJob {  // one chunk on its way through the engine
    uint64_t seq = 0;
    const uint8_t *in = nullptr;
    size_t n = 0;
    bool last = false;         // the stream ends with this chunk
    HostBuf *inbuf = nullptr;  // pool buffer holding `in`; nullptr = caller memory
    int dev = 0, slot = -1;
    bool skipped = false;      // cancelled / behind the point where the reference stops reading
    // ---- result ----
    HostBuf *outbuf = nullptr;
    const uint8_t *out = nullptr;
    size_t out_n = 0;
    uint64_t records = 0, nucleotides = 0;
    bool truncated = false;    // a line of >= 100 MiB: the reference stops reading inside this chunk
    std::vector<WarnOut> warns;
    bool warned = false;
    int rc = 0;
    std::string err;
    bool done = false;
    // ---- retirer: the device-to-host copy is in flight (issued, not yet waited for) ----
    bool d2h_issued = false;
    uint8_t *d2h_dst = nullptr;
    size_t d2h_total = 0;
    std::vector<gts::WarnRec> wr_raw;
    double t_disp = 0, t_issue0 = 0, t_issue1 = 0, t_comp = 0, t_d2h0 = 0, t_d2h1 = 0;  // GTS_TRACE=1: host clock, ms
};

static const bool g_trace = std::getenv("GTS_TRACE") != nullptr;
static double now_ms() {
    return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

struct PassSlot {  // device side of one chunk in flight
    uint8_t *d_in = nullptr, *d_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;
    gts::Totals *h_totals = nullptr;  // pinned
    u64 *d_warn = nullptr;
    gts::WarnRec *d_wres = nullptr;
    size_t warn_cap = 0;
    cudaEvent_t ev_h2d = nullptr, ev_comp = nullptr, ev_d2h = nullptr;
    bool busy = false;
};

struct DevCtx {  // one GPU: streams, scratch of a device pass, chunk slots, worker threads
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;  // kernels of the host entry points
    cudaStream_t s_side = nullptr;  // k_headers runs here, next to k_lines
    cudaStream_t s_side2 = nullptr; // ... and k_short here
    cudaEvent_t ev_join2 = nullptr;
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // ---- scratch of a device pass (grows, never shrinks) ----
    u64 cap_n = 0;  // raw bytes the per-tile scratch is sized for
    u32 *d_sym = nullptr;
    gts::TileInfo *d_tile_info = nullptr;
    gts::TilePrefix *d_tile_pre = nullptr;
    uint8_t *d_zero = nullptr;  // per-pass zeroed block: scan status, tickets, totals
    size_t zero_cap = 0;
    u64 rec_cap = 0;
    gts::HeaderEvent *d_events = nullptr;
    u64 *d_hdr_off = nullptr, *d_dbase = nullptr, *d_loc = nullptr, *d_hinfo = nullptr;
    gts::RecInfo *d_rec = nullptr;
    u32 *d_trim = nullptr;
    gts::FrameTab *d_ftab = nullptr;
    gts::NlPart *d_scan_nl = nullptr;
    size_t scan_nl_cap = 0;
    gts::Totals *h_totals = nullptr;  // pinned (device-resident entry points)
    u64 *d_dbg = nullptr;             // phase timing sums (GTS_PHASE_TIMING builds)
    double rec_per_byte = 1.0 / 128;  // record table sizing, raised when a pass overflows

    // ---- buffers of the piece entry points ----
    uint8_t *d_in = nullptr, *d_out = nullptr, *d_head = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;

    // ---- last enqueued device pass (device-resident entry points) ----
    bool pending = false;
    gts::Job job;
    cudaStream_t job_stream = nullptr;
    bool job_scan_only = false;
    // the slots in d_sym hold the parse of this buffer (a piece scanned, not yet translated)
    const void *slots_of = nullptr;
    size_t slots_n = 0;
    u64 *d_pt = nullptr;  // k_piece_trim's results
    gts::PieceParams *d_pp = nullptr;     // k_piece_plan's result
    gts::PieceInfoDev *h_infos = nullptr; // pinned copy of the all-gathered piece summaries
    int piece_rank = 0, piece_world = 1;

    // ---- per-kernel timing ----
    std::vector<cudaEvent_t> prof_events;  // 5 per recorded pass
    size_t prof_used = 0;

    // ---- chunk engine ----
    std::mutex mu;  // enqueue / scratch reallocation: the issuer and the retirer both use the scratch
    PassSlot slots[kSlots];
    std::deque<XJob *> inq, inflight;
    std::thread issuer, retirer;
};

}  // namespace

struct gts_ctx {
    gts_options opt;
    std::string frame;
    uint32_t mask = 0;
    bool reverse = false;
    gts::Lut lut;
    gts::LutBytes lutb;
    std::vector<std::unique_ptr<DevCtx>> devs;
    size_t chunk = (size_t)16 << 20;
    u64 max_line = 100ull << 20;  // bufio.Scanner token limit of the reference (transeq.go:27)
    bool poison = false;          // GTS_POISON_OUTPUT=1: fill the output with 0xEE first (tests)
    bool short_on = true;         // GTS_NO_SHORT=1: short records go through k_lines / k_headers as well (A/B runs)
    bool profiling = false;

    // ---- piece of a split record (gts_piece_*) ----
    bool piece_on = false;
    gts::Job piece_job;  // only the piece fields are used
    size_t piece_n = 0;
    bool piece_first = false;
    bool last_truncated = false;

    // ---- chunk engine (streaming and one-shot host entry points) ----
    std::mutex mu;
    std::condition_variable cv;
    bool workers_started = false, exiting = false;
    std::vector<HostBuf *> in_free, out_free;
    size_t in_count = 0, out_count = 0, in_max = 0, out_max = 0;
    size_t out_want = 0;  // largest chunk result so far: output buffers are (re)allocated at this size, so the pool settles after one pass over it
    std::deque<XJob *> window;  // submitted and not yet consumed, in input order
    uint64_t next_seq = 0, consume_seq = 0;
    uint64_t trunc_seq = ~0ull;  // chunk in which the reference stops reading
    bool cancelled = false;
    bool any_order = false;  // gts_options.any_order: results are handed back as they complete
    int e_rc = 0;  // first error of the stream (sticky until gts_stream_reset)
    std::string e_err;
    XJob *handed = nullptr;  // result currently with the caller
    // producer side (touched by the submitting thread only)
    HostBuf *fill = nullptr;
    size_t fill_n = 0, scan_from = 1, last_boundary = 0;
    bool eof = false, first_job = true;
    // one-shot mode: results go straight to their final place in h_out
    bool direct = false;
    uint8_t *h_out = nullptr;
    size_t h_out_cap = 0;
    uint64_t direct_turn = 0;
    size_t direct_off = 0;
    int direct_active = 0;
    bool direct_growing = false;
    // piece entry points
    uint8_t *h_piece = nullptr;
    size_t h_piece_cap = 0;

    gts_warning_fn warn_fn = nullptr;
    void *warn_user = nullptr;

    gts_stats stats;
    std::string err;
};

namespace {

// The calling thread's current device is put back when an entry point that visits several GPUs returns.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) prev = -1; }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

int fail(gts_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}
int cuda_code(cudaError_t e) { return e == cudaErrorMemoryAllocation ? GTS_ERR_NOMEM : GTS_ERR_CUDA; }
int cuda_fail(gts_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, cuda_code(e), std::string(what) + ": " + cudaGetErrorString(e));
}
// CU: inside API functions (message goes to ctx->err); CUE: inside helpers with a `std::string *err`
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
    } while (0)
#define CUE(call)                                                                  \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            if (err) *err = std::string(#call) + ": " + cudaGetErrorString(e_);    \
            return cuda_code(e_);                                                  \
        }                                                                          \
    } while (0)

struct ZeroLayout {
    size_t t_status, s_status, totals, counters, bytes;
};
ZeroLayout zero_layout(u64 ntiles, u64 rec_cap) {
    ZeroLayout z;
    const u64 scan_tiles = (ntiles + gts::kScanTile - 1) / gts::kScanTile;
    const u64 size_tiles = (rec_cap + gts::kSizeTile - 1) / gts::kSizeTile;
    size_t off = 0;
    z.t_status = off; off += scan_tiles * 4 * sizeof(u64);
    z.s_status = off; off += size_tiles * 4 * sizeof(u64);
    z.totals = off;   off += sizeof(gts::Totals);
    z.counters = off; off += 32;
    z.bytes = round_up(off, 16);
    return z;
}

void free_rec(DevCtx *dev) {
    cudaFree(dev->d_events); cudaFree(dev->d_hdr_off); cudaFree(dev->d_dbase); cudaFree(dev->d_loc); cudaFree(dev->d_hinfo);
    cudaFree(dev->d_rec); cudaFree(dev->d_trim); cudaFree(dev->d_ftab);
    dev->d_ftab = nullptr;
    dev->d_events = nullptr;
    dev->d_hdr_off = dev->d_dbase = dev->d_loc = dev->d_hinfo = nullptr;
    dev->d_rec = nullptr;
    dev->d_trim = nullptr;
    dev->rec_cap = 0;
}

// (the caller has made sure that no pass is in flight on this device)
int alloc_rec(gts_ctx *ctx, DevCtx *dev, u64 rec_cap, std::string *err) {
    free_rec(dev);
    CUE(cudaMalloc(&dev->d_events, rec_cap * sizeof(gts::HeaderEvent)));
    CUE(cudaMalloc(&dev->d_hdr_off, rec_cap * sizeof(u64)));
    CUE(cudaMalloc(&dev->d_dbase, rec_cap * sizeof(u64)));
    CUE(cudaMalloc(&dev->d_loc, rec_cap * sizeof(u64)));
    CUE(cudaMalloc(&dev->d_hinfo, rec_cap * sizeof(u64)));
    CUE(cudaMalloc(&dev->d_rec, rec_cap * sizeof(gts::RecInfo)));
    if (ctx->opt.trim) CUE(cudaMalloc(&dev->d_trim, rec_cap * 6 * sizeof(u32)));
    CUE(cudaMalloc(&dev->d_ftab, rec_cap * sizeof(gts::FrameTab)));
    dev->rec_cap = rec_cap;
    return GTS_OK;
}

u64 parse_tiles(u64 n) { return std::max<u64>(1, (n + gts::kParseTile - 1) / gts::kParseTile); }

// Size the scratch for a pass over n raw bytes.  Growing it frees the old blocks, so every stream
// of the device is drained first: a pass enqueued earlier (on any stream) may still be using them.
int ensure_scratch(gts_ctx *ctx, DevCtx *dev, u64 n, std::string *err) {
    const u64 want_rec = (u64)((double)n * dev->rec_per_byte) + 4096;
    const bool grow_tiles = n > dev->cap_n || dev->d_sym == nullptr;
    const bool grow_rec = want_rec > dev->rec_cap;
    if (!grow_tiles && !grow_rec) return GTS_OK;
    CUE(cudaDeviceSynchronize());
    if (grow_tiles) {
        const u64 cap = std::max<u64>(n + n / 8, 1u << 20);
        const u64 tiles = parse_tiles(cap);
        cudaFree(dev->d_sym); cudaFree(dev->d_tile_info); cudaFree(dev->d_tile_pre);
        dev->d_sym = nullptr; dev->d_tile_info = nullptr; dev->d_tile_pre = nullptr;
        dev->cap_n = 0;
        const size_t sym_bytes = (tiles + 2) * gts::kSlotWords * sizeof(u32);
        CUE(cudaMalloc(&dev->d_sym, sym_bytes));
        // zero once: stale words past the end of a slot must still hold valid symbols (<= 4)
        CUE(cudaMemset(dev->d_sym, 0, sym_bytes));
        // the memset runs on the legacy stream, the passes on non-blocking streams: without this
        // it can land after the first k_parse has written its slots
        CUE(cudaDeviceSynchronize());
        CUE(cudaMalloc(&dev->d_tile_info, tiles * sizeof(gts::TileInfo)));
        CUE(cudaMalloc(&dev->d_tile_pre, tiles * sizeof(gts::TilePrefix)));
        dev->cap_n = cap;
    }
    if (grow_rec) {
        int rc = alloc_rec(ctx, dev, want_rec + want_rec / 8, err);
        if (rc) return rc;
    }
    const size_t nscan = (parse_tiles(dev->cap_n) + gts::kScanTile - 1) / gts::kScanTile;
    if (nscan > dev->scan_nl_cap) {
        cudaFree(dev->d_scan_nl);
        dev->d_scan_nl = nullptr; dev->scan_nl_cap = 0;
        CUE(cudaMalloc(&dev->d_scan_nl, nscan * sizeof(gts::NlPart)));
        dev->scan_nl_cap = nscan;
    }
    const ZeroLayout z = zero_layout(parse_tiles(dev->cap_n), dev->rec_cap);
    if (z.bytes > dev->zero_cap) {
        cudaFree(dev->d_zero);
        dev->d_zero = nullptr; dev->zero_cap = 0;
        CUE(cudaMalloc(&dev->d_zero, z.bytes));
        dev->zero_cap = z.bytes;
    }
    return GTS_OK;
}

struct PassArgs {
    const uint8_t *d_in = nullptr;
    u64 n = 0;
    uint8_t *d_out = nullptr;
    u64 cap = 0;
    cudaStream_t stream = nullptr;
    bool sizes_only = false;
    bool scan_only = false;        // k_parse + k_tile_scan only (piece scan)
    bool reuse_slots = false;      // the slots of the pass before are still valid: no k_parse (piece translate)
    gts::Totals *h_dst = nullptr;  // pinned destination of the totals (default: dev->h_totals)
    bool publish = false;          // the totals reach h_dst by a store from the device, not by a copy (chunk engine)
    bool more_follows = false;
    u64 *d_warn = nullptr;         // warning list (nullptr = off)
    gts::WarnRec *d_wres = nullptr;
    size_t warn_cap = 0;
};

// Enqueue one device pass on a.stream (no host synchronisation).
int enqueue_pass(gts_ctx *ctx, DevCtx *dev, const PassArgs &a, std::string *err) {
    if (((uintptr_t)a.d_in & 15) != 0) {
        if (err) *err = "device input must be 16-byte aligned";
        return GTS_ERR_ARG;
    }
    int rc = ensure_scratch(ctx, dev, a.n, err);
    if (rc) return rc;
    const u64 ntiles = parse_tiles(a.n);
    const ZeroLayout z = zero_layout(ntiles, dev->rec_cap);
    gts::Job &job = dev->job;
    job.in = a.d_in; job.n = a.n; job.out = a.d_out; job.out_cap = a.cap;
    job.sym = dev->d_sym;
    job.tile_info = dev->d_tile_info;
    job.tile_pre = dev->d_tile_pre;
    job.events = dev->d_events;
    job.hdr_off = dev->d_hdr_off; job.dbase = dev->d_dbase; job.loc = dev->d_loc; job.hinfo = dev->d_hinfo;
    job.rec = dev->d_rec; job.trim_len = dev->d_trim; job.ftab = dev->d_ftab; job.rec_cap = dev->rec_cap;
    job.t_status = reinterpret_cast<u64 *>(dev->d_zero + z.t_status);
    job.s_status = reinterpret_cast<u64 *>(dev->d_zero + z.s_status);
    job.totals = reinterpret_cast<gts::Totals *>(dev->d_zero + z.totals);
    job.counters = reinterpret_cast<u32 *>(dev->d_zero + z.counters);
    job.warn = a.warn_cap ? a.d_warn : nullptr;
    job.wres = a.d_wres;
    job.warn_cap = (u32)std::min<size_t>(a.warn_cap, 0xFFFFFFFFu);
    job.pad0 = 0;
    job.dbg = dev->d_dbg;
    job.scan_nl = dev->d_scan_nl;
    job.max_line = ctx->max_line;
    job.ntiles = (u32)ntiles;
    job.frame_mask = ctx->mask;
    job.alternative = ctx->opt.alternative ? 1 : 0;
    job.trim = ctx->opt.trim ? 1 : 0;
    job.more_follows = a.more_follows ? 1 : 0;
    job.short_on = ctx->short_on ? 1 : 0;
    job.pad2 = 0;
    std::memset(&job.pv, 0, sizeof(job.pv));
    job.pp = nullptr;
    job.info_out = nullptr;
    job.ov_trim = 0; job.pad1 = 0; job.pt_before = 0; job.pt_out = nullptr;
    std::memset(job.ov_len, 0, sizeof(job.ov_len));
    std::memset(job.pt_len, 0, sizeof(job.pt_len));
    if (ctx->piece_on) {
        const gts::Job &p = ctx->piece_job;
        job.pv = p.pv;
        job.pp = p.pp;
        job.info_out = p.info_out;
        job.ov_trim = p.ov_trim;
        std::memcpy(job.ov_len, p.ov_len, sizeof(job.ov_len));
    }
    job.lut = ctx->lut;
    job.lutb = ctx->lutb;
    dev->job_stream = a.stream;
    dev->job_scan_only = a.scan_only;
    dev->slots_of = nullptr;  // (set again by the piece scan once its pass has finished)
    cudaEvent_t *ev = nullptr;
    if (ctx->profiling && !a.sizes_only) {
        while (dev->prof_events.size() < dev->prof_used + 5) {
            cudaEvent_t e;
            CUE(cudaEventCreate(&e));
            dev->prof_events.push_back(e);
        }
        ev = dev->prof_events.data() + dev->prof_used;
        dev->prof_used += 5;
        CUE(cudaEventRecord(ev[0], a.stream));
    }
    CUE(cudaMemsetAsync(dev->d_zero, 0, z.bytes, a.stream));
    u64 launches = 0;
    if (a.reuse_slots) {
        CUE(gts::launch_piece_fix(job, a.stream));
    } else {
        CUE(gts::launch_parse(job, a.stream));
    }
    launches++;
    if (ev) CUE(cudaEventRecord(ev[1], a.stream));
    if (a.scan_only) {
        CUE(gts::launch_tile_scan(job, dev->sm_count, a.stream));
        CUE(cudaMemcpyAsync(a.h_dst ? a.h_dst : dev->h_totals, job.totals, sizeof(gts::Totals), cudaMemcpyDeviceToHost, a.stream));
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->stats.kernel_launches += launches + 1;
        ctx->stats.passes++;
        return GTS_OK;
    }
    CUE(gts::launch_sizes(job, dev->sm_count, a.stream));
    launches += 3;
    if (job.warn) {
        CUE(gts::launch_warn(job, dev->sm_count, a.stream));
        launches++;
    }
    if (ev) CUE(cudaEventRecord(ev[2], a.stream));
    if (!a.sizes_only) {
        gts::EmitStreams es;
        es.side = dev->s_side; es.side2 = dev->s_side2; es.fork = dev->ev_fork; es.join = dev->ev_join; es.join2 = dev->ev_join2;
        CUE(gts::launch_emit(job, dev->sm_count, a.stream, ev ? ev[3] : nullptr, es));
        if (ev) CUE(cudaEventRecord(ev[4], a.stream));
        launches += job.short_on ? 3 : 2;
    }
    if (a.publish && a.h_dst) {
        // A copy would queue up behind the device-to-host copy of the chunk before (one copy engine per
        // direction) and delay this chunk's completion event by the whole of it.
        CUE(gts::launch_publish(job.totals, a.h_dst, a.stream));
        launches++;
    } else {
        CUE(cudaMemcpyAsync(a.h_dst ? a.h_dst : dev->h_totals, job.totals, sizeof(gts::Totals), cudaMemcpyDeviceToHost, a.stream));
    }
    {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->stats.kernel_launches += launches;
        ctx->stats.passes++;
    }
    return GTS_OK;
}

// Device-resident entry points: wait for the enqueued pass; grow the record table and redo the
// pass if it was too small; redo it on the bytes in front of an over-long line.
int finish_pass(gts_ctx *ctx, DevCtx *dev, bool sizes_only) {
    if (!dev->pending) return fail(ctx, GTS_ERR_ARG, "no device pass enqueued");
    std::string emsg, *err = &emsg;
    for (int attempt = 0;; attempt++) {
        CU(cudaStreamSynchronize(dev->job_stream));
        const gts::Totals &t = *dev->h_totals;
        PassArgs a;
        a.d_in = dev->job.in; a.n = dev->job.n; a.d_out = dev->job.out; a.cap = dev->job.out_cap;
        a.stream = dev->job_stream; a.sizes_only = sizes_only;
        a.scan_only = dev->job_scan_only;  // (a retry parses again: the slots are not reused)
        if (t.flags & gts::kFlagLongLine) {
            // the reference stops reading at a line of >= 100 MiB (transeq.go:186): redo the pass on
            // the bytes in front of that line
            if (attempt >= 3) return fail(ctx, GTS_ERR_ARG, "over-long line handling did not converge");
            ctx->last_truncated = true;
            a.n = t.cut;
        } else if (t.flags & gts::kFlagRecOverflow) {
            if (attempt >= 3) return fail(ctx, GTS_ERR_NOMEM, "record table overflow persists");
            const u64 need = t.n_headers + 2;
            CU(cudaDeviceSynchronize());
            int rc = alloc_rec(ctx, dev, need + need / 16 + 1024, err);
            if (rc) return fail(ctx, rc, emsg);
            dev->rec_per_byte = std::max(dev->rec_per_byte, 1.1 * (double)need / (double)std::max<u64>(a.n, 1));
        } else {
            break;
        }
        int rc = enqueue_pass(ctx, dev, a, err);
        if (rc) return fail(ctx, rc, emsg);
    }
    dev->pending = false;
    const gts::Totals &t = *dev->h_totals;
    ctx->stats.records = t.records;
    ctx->stats.nucleotides = t.nucleotides;
    ctx->stats.in_bytes = dev->job.n;
    ctx->stats.out_bytes = t.out_total;
    return GTS_OK;
}

int ensure_dev_buffer(uint8_t **p, size_t *cap, size_t need, std::string *err) {
    if (need <= *cap && *p) return GTS_OK;
    cudaFree(*p);  // (waits for the device: nothing enqueued earlier still uses the block)
    *p = nullptr; *cap = 0;
    const size_t want = round_up(need + need / 8 + 4096, 4096);
    CUE(cudaMalloc(p, want));
    *cap = want;
    return GTS_OK;
}
int ensure_pinned(uint8_t **p, size_t *cap, size_t need, size_t keep, std::string *err) {
    if (need <= *cap && *p) return GTS_OK;
    const size_t want = round_up(need + need / 8 + 4096, 4096);
    uint8_t *q = nullptr;
    const double t0 = g_trace ? now_ms() : 0;
    CUE(cudaHostAlloc(&q, want, cudaHostAllocPortable));
    if (g_trace) std::fprintf(stderr, "[gts trace] pinned allocation of %zu bytes (was %zu): %.3f ms\n", want, *cap, now_ms() - t0);
    if (keep && *p) std::memcpy(q, *p, keep);
    if (*p) cudaFreeHost(*p);
    *p = q;
    *cap = want;
    return GTS_OK;
}

// ---- device contexts --------------------------------------------------------------------

void dev_destroy(DevCtx *dev) {
    cudaSetDevice(dev->device);
    cudaDeviceSynchronize();
    if (dev->d_dbg) {
        u64 h[64];
        cudaMemcpy(h, dev->d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "[gts phase timing, ns summed over tiles]");
        for (int i = 0; i < 24; i++) std::fprintf(stderr, " p%d=%llu", i, (unsigned long long)h[i]);
        std::fprintf(stderr, "\n");
        cudaFree(dev->d_dbg);
    }
    cudaFree(dev->d_sym); cudaFree(dev->d_tile_info); cudaFree(dev->d_tile_pre);
    cudaFree(dev->d_zero); cudaFree(dev->d_scan_nl);
    free_rec(dev);
    cudaFree(dev->d_in); cudaFree(dev->d_out); cudaFree(dev->d_head); cudaFree(dev->d_pt); cudaFree(dev->d_pp);
    if (dev->h_infos) cudaFreeHost(dev->h_infos);
    for (PassSlot &s : dev->slots) {
        cudaFree(s.d_in); cudaFree(s.d_out); cudaFree(s.d_warn); cudaFree(s.d_wres);
        if (s.h_totals) cudaFreeHost(s.h_totals);
        if (s.ev_h2d) cudaEventDestroy(s.ev_h2d);
        if (s.ev_comp) cudaEventDestroy(s.ev_comp);
        if (s.ev_d2h) cudaEventDestroy(s.ev_d2h);
    }
    if (dev->h_totals) cudaFreeHost(dev->h_totals);
    for (cudaEvent_t e : dev->prof_events) cudaEventDestroy(e);
    if (dev->stream) cudaStreamDestroy(dev->stream);
    if (dev->s_side) cudaStreamDestroy(dev->s_side);
    if (dev->s_side2) cudaStreamDestroy(dev->s_side2);
    if (dev->ev_join2) cudaEventDestroy(dev->ev_join2);
    if (dev->s_h2d) cudaStreamDestroy(dev->s_h2d);
    if (dev->s_d2h) cudaStreamDestroy(dev->s_d2h);
    if (dev->ev_fork) cudaEventDestroy(dev->ev_fork);
    if (dev->ev_join) cudaEventDestroy(dev->ev_join);
}

// No CPU fallback: without a usable sm_100 device creation fails.
int dev_create(int ordinal, std::unique_ptr<DevCtx> *out, std::string *err) {
    std::unique_ptr<DevCtx> dev(new DevCtx());
    int d = ordinal;
    if (d < 0) CUE(cudaGetDevice(&d)); else CUE(cudaSetDevice(d));
    CUE(cudaFree(nullptr));  // force context creation
    cudaDeviceProp prop;
    CUE(cudaGetDeviceProperties(&prop, d));
    if (prop.major < 10) {
        *err = "gotranseq_b200 needs an sm_100a device (found sm_" + std::to_string(prop.major) + std::to_string(prop.minor) + ")";
        return GTS_ERR_CUDA;
    }
    dev->device = d;
    dev->sm_count = prop.multiProcessorCount;
    std::memset(&dev->job, 0, sizeof(dev->job));
    *out = std::move(dev);
    DevCtx *p = out->get();
    CUE(gts::init_kernels());
    CUE(cudaStreamCreateWithFlags(&p->stream, cudaStreamNonBlocking));
    {
        // k_headers / k_short run next to the persistent k_lines grid.  GTS_SIDE_PRIO=1 (A/B runs) gives
        // their streams the highest priority: measured, it moves time from the tail into k_lines and
        // leaves the step where it was (profiles/r2_notes.md), so plain streams are the default
        int lo = 0, hi = 0;
        CUE(cudaDeviceGetStreamPriorityRange(&lo, &hi));
        const char *sp = std::getenv("GTS_SIDE_PRIO");
        const int prio = (sp && std::atoi(sp) != 0) ? hi : lo;
        CUE(cudaStreamCreateWithPriority(&p->s_side, cudaStreamNonBlocking, prio));
        CUE(cudaStreamCreateWithPriority(&p->s_side2, cudaStreamNonBlocking, prio));
    }
    CUE(cudaEventCreateWithFlags(&p->ev_join2, cudaEventDisableTiming));
    CUE(cudaStreamCreateWithFlags(&p->s_h2d, cudaStreamNonBlocking));
    CUE(cudaStreamCreateWithFlags(&p->s_d2h, cudaStreamNonBlocking));
    CUE(cudaEventCreateWithFlags(&p->ev_fork, cudaEventDisableTiming));
    CUE(cudaEventCreateWithFlags(&p->ev_join, cudaEventDisableTiming));
    CUE(cudaHostAlloc(&p->h_totals, sizeof(gts::Totals), cudaHostAllocPortable));
    for (PassSlot &s : p->slots) {
        CUE(cudaHostAlloc(&s.h_totals, sizeof(gts::Totals), cudaHostAllocPortable));
        CUE(cudaEventCreateWithFlags(&s.ev_h2d, cudaEventDisableTiming));
        CUE(cudaEventCreateWithFlags(&s.ev_comp, cudaEventDisableTiming));
        CUE(cudaEventCreateWithFlags(&s.ev_d2h, cudaEventDisableTiming));
    }
#ifdef GTS_PHASE_TIMING
    if (std::getenv("GTS_PHASE_TIMING")) {
        cudaMalloc(&p->d_dbg, 64 * sizeof(u64));
        cudaMemset(p->d_dbg, 0, 64 * sizeof(u64));
    }
#endif
    return GTS_OK;
}

// ---- chunk engine -----------------------------------------------------------------------

bool stream_dead(const gts_ctx *ctx) { return ctx->cancelled || ctx->e_rc != 0 || ctx->exiting; }

int stream_error_locked(gts_ctx *ctx) {
    if (ctx->e_rc) {
        ctx->err = ctx->e_err;
        return ctx->e_rc;
    }
    if (ctx->cancelled) {
        ctx->err = "translation cancelled";
        return GTS_ERR_CANCELLED;
    }
    return GTS_OK;
}

void set_stream_error_locked(gts_ctx *ctx, int rc, const std::string &msg) {
    if (ctx->e_rc == 0) {
        ctx->e_rc = rc;
        ctx->e_err = msg;
    }
    ctx->cv.notify_all();
}

// A pinned buffer from a pool (allocated on demand up to the pool's limit); nullptr when the stream died.
HostBuf *pool_take(gts_ctx *ctx, std::unique_lock<std::mutex> &lk, std::vector<HostBuf *> &pool, size_t &count,
                   size_t max_count, size_t min_cap, std::string *err, int *rc) {
    *rc = GTS_OK;
    ctx->cv.wait(lk, [&] { return stream_dead(ctx) || !pool.empty() || count < max_count; });
    if (stream_dead(ctx)) return nullptr;
    HostBuf *b = nullptr;
    if (!pool.empty()) {
        b = pool.back();
        pool.pop_back();
    } else {
        count++;
        b = new HostBuf();
    }
    if (b->cap < min_cap) {
        lk.unlock();
        *rc = ensure_pinned(&b->p, &b->cap, min_cap, 0, err);
        lk.lock();
        if (*rc) {
            pool.push_back(b);
            return nullptr;
        }
    }
    return b;
}

int free_slot(const DevCtx *dev) {
    for (int i = 0; i < g_slots; i++)
        if (!dev->slots[i].busy) return i;
    return -1;
}

size_t out_estimate(size_t n) { return n * 4 + 4096; }

// H2D copy and the kernels of one chunk, enqueued without waiting.
int issue_job(gts_ctx *ctx, DevCtx *dev, XJob *job) {
    std::string *err = &job->err;
    PassSlot &s = dev->slots[job->slot];
    std::lock_guard<std::mutex> dl(dev->mu);
    int rc = ensure_dev_buffer(&s.d_in, &s.d_in_cap, job->n + 64, err);
    if (rc) return rc;
    rc = ensure_dev_buffer(&s.d_out, &s.d_out_cap, out_estimate(job->n), err);
    if (rc) return rc;
    if (ctx->warn_fn) {
        const size_t want = job->n / 16 + 4096;
        if (want > s.warn_cap) {
            cudaFree(s.d_warn); cudaFree(s.d_wres);
            s.d_warn = nullptr; s.d_wres = nullptr; s.warn_cap = 0;
            CUE(cudaMalloc(&s.d_warn, want * sizeof(u64)));
            CUE(cudaMalloc(&s.d_wres, want * sizeof(gts::WarnRec)));
            s.warn_cap = want;
        }
    }
    rc = ensure_scratch(ctx, dev, job->n, err);  // before the copy: growing the scratch drains the device
    if (rc) return rc;
    if (job->n) CUE(cudaMemcpyAsync(s.d_in, job->in, job->n, cudaMemcpyHostToDevice, dev->s_h2d));
    CUE(cudaEventRecord(s.ev_h2d, dev->s_h2d));
    CUE(cudaStreamWaitEvent(dev->stream, s.ev_h2d, 0));
    if (ctx->poison) CUE(cudaMemsetAsync(s.d_out, 0xEE, s.d_out_cap, dev->stream));
    PassArgs a;
    a.d_in = s.d_in; a.n = job->n; a.d_out = s.d_out; a.cap = s.d_out_cap; a.stream = dev->stream;
    a.h_dst = s.h_totals; a.more_follows = !job->last; a.publish = true;
    if (ctx->warn_fn) { a.d_warn = s.d_warn; a.d_wres = s.d_wres; a.warn_cap = s.warn_cap; }
    rc = enqueue_pass(ctx, dev, a, err);
    if (rc) return rc;
    CUE(cudaEventRecord(s.ev_comp, dev->stream));
    return GTS_OK;
}

// Wait for a chunk's kernels, repair the rare overflows, start the copy of the result to the host.
// The copy is NOT waited for (retire_finish does that): the retirer goes on to the next chunk, whose
// copy queues up behind this one, so the device-to-host stream never idles between chunks.
// `drain` finishes the copies this retirer still has in flight; it is called before any wait that
// only the consumer of those chunks can end (an output buffer, the one-shot output growing).
struct PendingCopies {  // the retirer's copies in flight, seen from retire_issue
    std::function<void()> drain;
    std::function<bool()> any;
};
int retire_issue(gts_ctx *ctx, DevCtx *dev, XJob *job, const PendingCopies &pc) {
    const std::function<void()> &drain = pc.drain;
    std::string *err = &job->err;
    PassSlot &s = dev->slots[job->slot];
    CUE(cudaEventSynchronize(s.ev_comp));
    if (g_trace) job->t_comp = now_ms();
    size_t n_eff = job->n;
    bool more_follows = !job->last;
    for (int attempt = 0; s.h_totals->flags != 0; attempt++) {
        // redo the pass on this slot: over-long line (the reference stops reading there), record
        // table / output buffer / warning list too small.  The other chunks in flight on this GPU
        // are finished first (they share the scratch).
        const gts::Totals t = *s.h_totals;
        if (attempt >= 6) {
            *err = "device pass did not converge (flags " + std::to_string(t.flags) + ")";
            return GTS_ERR_NOMEM;
        }
        std::lock_guard<std::mutex> dl(dev->mu);
        CUE(cudaDeviceSynchronize());
        if (t.flags & gts::kFlagLongLine) {
            job->truncated = true;
            n_eff = t.cut;
            more_follows = false;
        } else if (t.flags & gts::kFlagRecOverflow) {
            const u64 need = t.n_headers + 2;
            int rc = alloc_rec(ctx, dev, need + need / 16 + 1024, err);
            if (rc) return rc;
            dev->rec_per_byte = std::max(dev->rec_per_byte, 1.1 * (double)need / (double)std::max<size_t>(job->n, 1));
        } else if (t.flags & gts::kFlagWarnOverflow) {
            const size_t want = t.n_warn + t.n_warn / 8 + 4096;
            cudaFree(s.d_warn); cudaFree(s.d_wres);
            s.d_warn = nullptr; s.d_wres = nullptr; s.warn_cap = 0;
            CUE(cudaMalloc(&s.d_warn, want * sizeof(u64)));
            CUE(cudaMalloc(&s.d_wres, want * sizeof(gts::WarnRec)));
            s.warn_cap = want;
        } else if (t.flags & gts::kFlagOutOverflow) {
            int rc = ensure_dev_buffer(&s.d_out, &s.d_out_cap, t.out_total + 64, err);
            if (rc) return rc;
        }
        PassArgs a;
        a.d_in = s.d_in; a.n = n_eff; a.d_out = s.d_out; a.cap = s.d_out_cap; a.stream = dev->stream;
        a.h_dst = s.h_totals; a.more_follows = more_follows; a.publish = true;
        if (ctx->warn_fn) { a.d_warn = s.d_warn; a.d_wres = s.d_wres; a.warn_cap = s.warn_cap; }
        if (ctx->poison) CUE(cudaMemsetAsync(s.d_out, 0xEE, s.d_out_cap, dev->stream));
        int rc = enqueue_pass(ctx, dev, a, err);
        if (rc) return rc;
        CUE(cudaStreamSynchronize(dev->stream));
    }
    const gts::Totals t = *s.h_totals;
    job->records = t.records;
    job->nucleotides = t.nucleotides;
    size_t total = t.out_total;

    // ---- where the result goes ----
    uint8_t *dst = nullptr;
    if (ctx->direct) {
        // one-shot call: straight to its final place behind the chunks before it
        std::unique_lock<std::mutex> lk(ctx->mu);
        while (true) {
            ctx->cv.wait(lk, [&] {
                return stream_dead(ctx) || ctx->direct_turn == job->seq || job->seq > ctx->trunc_seq || (ctx->direct_growing && pc.any());
            });
            if (stream_dead(ctx) || ctx->direct_turn == job->seq || job->seq > ctx->trunc_seq) break;
            // the one-shot output is being moved to a larger block: that waits for every copy in flight, ours too
            lk.unlock();
            drain();
            lk.lock();
        }
        if (stream_dead(ctx) || job->seq > ctx->trunc_seq) return GTS_OK;  // (the reference never read this far)
        if (job->truncated) ctx->trunc_seq = job->seq;
        const size_t off = ctx->direct_off;
        if (off + total > ctx->h_out_cap) {
            lk.unlock();
            drain();  // (our own copies count in direct_active)
            lk.lock();
            ctx->direct_growing = true;
            ctx->cv.notify_all();  // (retirers waiting for their turn finish their copies in flight)
            ctx->cv.wait(lk, [&] { return ctx->direct_active == 0; });
            lk.unlock();
            int rc = ensure_pinned(&ctx->h_out, &ctx->h_out_cap, (off + total) + (off + total) / 2, off, err);
            lk.lock();
            ctx->direct_growing = false;
            if (rc) {
                ctx->cv.notify_all();
                return rc;
            }
        }
        ctx->direct_off = off + total;
        ctx->direct_turn++;
        ctx->direct_active++;
        dst = ctx->h_out + off;
        ctx->cv.notify_all();
    } else {
        std::unique_lock<std::mutex> lk(ctx->mu);
        // (window rule: only the out_max oldest chunks may hold an output buffer, so the chunk the
        // caller waits for can always get one)
        // (any_order: the caller takes whatever chunk is complete, so whoever holds a buffer gets consumed --
        // except behind a chunk that is large enough to hold an over-long line and not done yet: nothing
        // behind it is handed over before it is (next_output), so nothing behind it may take a buffer
        // that it could be left waiting for)
        auto in_window = [&] {
            if (!ctx->any_order) return job->seq < ctx->consume_seq + ctx->out_max;
            for (const XJob *j : ctx->window) {
                if (j == job) break;
                if (!j->done && j->n >= ctx->max_line) return false;
            }
            return true;
        };
        auto ready = [&] {
            return stream_dead(ctx) || (in_window() && (!ctx->out_free.empty() || ctx->out_count < ctx->out_max));
        };
        if (!ready()) {
            // the caller may be waiting for a chunk whose copy this thread has not finished yet
            lk.unlock();
            drain();
            lk.lock();
        }
        ctx->cv.wait(lk, [&] { return stream_dead(ctx) || in_window(); });
        if (stream_dead(ctx)) return GTS_OK;
        int rc = GTS_OK;
        ctx->out_want = std::max(ctx->out_want, total + 64);
        HostBuf *b = pool_take(ctx, lk, ctx->out_free, ctx->out_count, ctx->out_max, ctx->out_want, err, &rc);
        if (!b) return rc;
        job->outbuf = b;
        dst = b->p;
    }
    cudaError_t ce = cudaSuccess;
    if (total) ce = cudaMemcpyAsync(dst, s.d_out, total, cudaMemcpyDeviceToHost, dev->s_d2h);
    if (ce == cudaSuccess && ctx->warn_fn && t.n_warn) {
        job->wr_raw.resize(t.n_warn);
        ce = cudaMemcpyAsync(job->wr_raw.data(), s.d_wres, t.n_warn * sizeof(gts::WarnRec), cudaMemcpyDeviceToHost, dev->s_d2h);
    }
    if (ce == cudaSuccess) ce = cudaEventRecord(s.ev_d2h, dev->s_d2h);
    if (g_trace) job->t_d2h0 = now_ms();
    job->d2h_issued = true;  // (also after a failed enqueue: retire_finish settles direct_active)
    job->d2h_dst = dst;
    job->d2h_total = total;
    if (ce != cudaSuccess) {
        *err = std::string("device to host copy: ") + cudaGetErrorString(ce);
        return cuda_code(ce);
    }
    return GTS_OK;
}

// Wait for the chunk's device-to-host copy; its result and warnings become visible to the caller.
int retire_finish(gts_ctx *ctx, DevCtx *dev, XJob *job) {
    std::string *err = &job->err;
    PassSlot &s = dev->slots[job->slot];
    int rc = job->rc;
    if (rc == GTS_OK) {
        const cudaError_t ce = cudaEventSynchronize(s.ev_d2h);
        if (g_trace) job->t_d2h1 = now_ms();
        if (ce != cudaSuccess) {
            *err = std::string("device to host copy: ") + cudaGetErrorString(ce);
            rc = cuda_code(ce);
        }
    } else {
        cudaStreamSynchronize(dev->s_d2h);  // whatever was enqueued of it
    }
    job->d2h_issued = false;
    if (ctx->direct) {
        std::lock_guard<std::mutex> lk(ctx->mu);
        ctx->direct_active--;
        ctx->cv.notify_all();
    }
    if (rc) return rc;
    job->out = job->d2h_dst;
    job->out_n = job->d2h_total;
    std::vector<gts::WarnRec> &wr = job->wr_raw;
    if (!wr.empty()) {
        // the reader goroutine prints the warnings in input order (encodedSequence.go:39)
        std::sort(wr.begin(), wr.end(), [](const gts::WarnRec &x, const gts::WarnRec &y) { return x.key < y.key; });
        job->warns.reserve(wr.size());
        for (const gts::WarnRec &w : wr) {
            WarnOut o;
            o.header.assign(reinterpret_cast<const char *>(job->in + w.hdr_off), w.H);
            o.pos = w.pos;
            o.byte = (uint8_t)w.byte;
            job->warns.push_back(std::move(o));
        }
        wr.clear();
        wr.shrink_to_fit();
    }
    return GTS_OK;
}

void issuer_main(gts_ctx *ctx, DevCtx *dev) {
    cudaSetDevice(dev->device);
    std::unique_lock<std::mutex> lk(ctx->mu);
    while (true) {
        ctx->cv.wait(lk, [&] { return ctx->exiting || (!dev->inq.empty() && free_slot(dev) >= 0); });
        if (ctx->exiting) return;
        XJob *job = dev->inq.front();
        dev->inq.pop_front();
        job->slot = free_slot(dev);
        dev->slots[job->slot].busy = true;
        job->skipped = ctx->cancelled || ctx->e_rc != 0 || job->seq > ctx->trunc_seq;
        lk.unlock();
        if (g_trace) job->t_issue0 = now_ms();
        if (!job->skipped) job->rc = issue_job(ctx, dev, job);
        if (g_trace) job->t_issue1 = now_ms();
        lk.lock();
        dev->inflight.push_back(job);
        ctx->cv.notify_all();
    }
}

void retirer_main(gts_ctx *ctx, DevCtx *dev) {
    cudaSetDevice(dev->device);
    std::deque<XJob *> pend;  // chunks whose device-to-host copy is in flight, oldest first (this thread only)
    std::unique_lock<std::mutex> lk(ctx->mu);
    // (called with the lock held) the chunk leaves the engine: slot and input buffer free, counters, `done`
    auto complete = [&](XJob *job, bool run) {
        dev->slots[job->slot].busy = false;
        if (job->inbuf) {
            ctx->in_free.push_back(job->inbuf);
            job->inbuf = nullptr;
        }
        if (job->rc) set_stream_error_locked(ctx, job->rc, job->err);
        if (job->truncated && job->seq < ctx->trunc_seq) ctx->trunc_seq = job->seq;
        ctx->stats.h2d_bytes += run ? job->n : 0;
        ctx->stats.d2h_bytes += job->out_n;
        if (!ctx->direct && run && job->rc == 0) {  // a stream's counters add up over its chunks
            ctx->stats.records += job->records;
            ctx->stats.nucleotides += job->nucleotides;
            ctx->stats.in_bytes += job->n;
            ctx->stats.out_bytes += job->out_n;
        }
        job->done = true;
        ctx->cv.notify_all();
    };
    // (called WITHOUT the lock) finish the oldest copy in flight
    auto finish_one = [&] {
        XJob *job = pend.front();
        pend.pop_front();
        job->rc = retire_finish(ctx, dev, job);
        std::lock_guard<std::mutex> g(ctx->mu);
        complete(job, true);
    };
    PendingCopies pc;
    pc.drain = [&] { while (!pend.empty()) finish_one(); };
    pc.any = [&] { return !pend.empty(); };
    const std::function<void()> &drain = pc.drain;
    while (true) {
        ctx->cv.wait(lk, [&] { return ctx->exiting || !dev->inflight.empty() || !pend.empty(); });
        if (ctx->exiting) return;
        if (!dev->inflight.empty() && pend.size() < 2) {
            // the next chunk: wait for its kernels and queue its copy behind the ones in flight
            XJob *job = dev->inflight.front();
            dev->inflight.pop_front();
            const bool run = !job->skipped && job->rc == 0;
            lk.unlock();
            if (run) {
                job->rc = retire_issue(ctx, dev, job, pc);
                if (job->d2h_issued) {
                    pend.push_back(job);
                    lk.lock();
                    continue;
                }
            } else if (!job->skipped) {
                cudaStreamSynchronize(dev->stream);  // a failed enqueue: let what was enqueued finish
            }
            drain();  // chunks leave in order
            lk.lock();
            complete(job, run);
        } else {
            lk.unlock();
            finish_one();
            lk.lock();
        }
    }
}

int start_workers(gts_ctx *ctx) {
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (ctx->workers_started) return GTS_OK;
    for (auto &d : ctx->devs) {
        d->issuer = std::thread(issuer_main, ctx, d.get());
        d->retirer = std::thread(retirer_main, ctx, d.get());
    }
    ctx->workers_started = true;
    return GTS_OK;
}

// Hand a chunk to its GPU (round robin over the chunks keeps every GPU's queue equally long).
void dispatch_locked(gts_ctx *ctx, XJob *job) {
    job->seq = ctx->next_seq++;
    if (g_trace) job->t_disp = now_ms();
    job->dev = (int)(job->seq % ctx->devs.size());
    ctx->window.push_back(job);
    ctx->devs[job->dev]->inq.push_back(job);
    ctx->cv.notify_all();
}

void recycle_job_locked(gts_ctx *ctx, XJob *job) {
    if (job->outbuf) ctx->out_free.push_back(job->outbuf);
    if (job->inbuf) ctx->in_free.push_back(job->inbuf);
    delete job;
}

// Everything queued is finished (or skipped) and forgotten; the pools get their buffers back.
void drain_and_clear_locked(gts_ctx *ctx, std::unique_lock<std::mutex> &lk) {
    const bool was_cancelled = ctx->cancelled;
    if (!ctx->window.empty()) ctx->cancelled = true;  // queued chunks are dropped, running ones finish
    ctx->cv.notify_all();
    ctx->cv.wait(lk, [&] {
        for (XJob *j : ctx->window)
            if (!j->done) return false;
        return true;
    });
    for (XJob *j : ctx->window) recycle_job_locked(ctx, j);
    ctx->window.clear();
    ctx->handed = nullptr;
    ctx->cancelled = was_cancelled;
}

void reset_stream_locked(gts_ctx *ctx) {
    ctx->consume_seq = ctx->next_seq;
    ctx->trunc_seq = ~0ull;
    ctx->eof = false;
    ctx->first_job = true;
    if (ctx->fill) {
        ctx->in_free.push_back(ctx->fill);
        ctx->fill = nullptr;
    }
    ctx->fill_n = 0;
    ctx->scan_from = 1;
    ctx->last_boundary = 0;
}

void fire_warnings(gts_ctx *ctx, XJob *job) {
    if (job->warned) return;
    job->warned = true;
    if (!ctx->warn_fn) return;
    for (const WarnOut &w : job->warns)
        ctx->warn_fn(ctx->warn_user, reinterpret_cast<const uint8_t *>(w.header.data()), w.header.size(), w.byte, w.pos);
    job->warns.clear();
    job->warns.shrink_to_fit();
}

// The next result in input order.  block: wait for it (after eof gts_next_output waits as well).
int next_output(gts_ctx *ctx, const uint8_t **buf, size_t *n, bool block) {
    *buf = nullptr;
    *n = 0;
    std::unique_lock<std::mutex> lk(ctx->mu);
    while (true) {
        int rc = stream_error_locked(ctx);
        if (rc) return rc;
        if (ctx->handed) {  // not released yet: the same result again
            *buf = ctx->handed->out;
            *n = ctx->handed->out_n;
            return GTS_OK;
        }
        // the oldest chunk -- or, with any_order, the oldest COMPLETE chunk that no chunk in front of it can
        // still silence: a line of >= max_line bytes (transeq.go:186) lies inside one record, hence inside one
        // chunk, so only a chunk of at least that size can turn out to be where the reference stops reading
        XJob *ready_job = nullptr;
        for (XJob *j : ctx->window) {
            if (j->done) { ready_job = j; break; }
            if (!ctx->any_order || j->n >= ctx->max_line) break;
        }
        if (ready_job) {
            XJob *job = ready_job;
            if (job->seq > ctx->trunc_seq) {  // the reference stopped reading before this chunk
                job->out_n = 0;
                job->warns.clear();
            }
            if (!job->warned && !job->warns.empty()) {
                lk.unlock();
                fire_warnings(ctx, job);  // on the caller's thread, without the engine lock
                lk.lock();
                continue;
            }
            if (job->out_n == 0) {
                ctx->window.erase(std::find(ctx->window.begin(), ctx->window.end(), job));
                ctx->consume_seq++;
                recycle_job_locked(ctx, job);
                ctx->cv.notify_all();
                continue;
            }
            ctx->handed = job;
            *buf = job->out;
            *n = job->out_n;
            return GTS_OK;
        }
        if (ctx->eof && ctx->window.empty()) {  // drained: ready for another stream
            reset_stream_locked(ctx);
            return GTS_OK;
        }
        if (!block && !ctx->eof) return GTS_OK;
        ctx->cv.wait(lk);
    }
}

// ---- split record: which bytes of the record's output a piece writes --------------------------
// Same frame arithmetic as the kernels (gts_common.cuh), for the whole piece at once.
u64 host_reverse_start(u64 n, int i, bool alternative) {
    if (alternative) return (u64)i;
    int r = (int)(n % 3) - i;
    return (u64)(r < 0 ? r + 3 : r);
}
void host_frame_lengths(u64 n, bool alternative, u64 L[6]) {
    for (int k = 0; k < 3; k++) L[k] = n >= (u64)k ? (n - k + 2) / 3 : 0;
    for (int i = 0; i < 3; i++) {
        const u64 p = host_reverse_start(n, i, alternative);
        L[3 + i] = n >= p ? (n - p + 2) / 3 : 0;
    }
}
u64 host_run_size(u64 H, u64 L) { return H + 3 + L + (L + 59) / 60; }

int piece_segments(uint32_t mask, bool alternative, const gts_piece &p, u64 n_piece, gts_segment *segs, u64 *total) {
    const u64 n = p.nt_total, H = p.header_len;
    u64 L[6];
    host_frame_lengths(n, alternative, L);
    if (p.has_trim_len)
        for (int f = 0; f < 6; f++) L[f] = p.trim_len[f];
    // Output line b of a strand's frames is one block of 180 nucleotides; the block belongs to the
    // piece that holds its anchor, the last symbol of its lowest codon (k_lines, gts_lines.cu).
    // Anchors as indices into the record's nucleotides (n, n + 1 = the two pads behind it):
    // forward 2 + 180 b; reverse n - 189 - 180 b, or 0 for the blocks b >= Lc at the record's start.
    const u64 a = p.nt_before, e = p.nt_before + n_piece + (p.last ? 2 : 0);
    const u64 f_lo = a > 2 ? (a - 2 + 179) / 180 : 0, f_hi = e > 2 ? (e - 2 + 179) / 180 : 0;
    const u64 Lc = n >= 189 ? (n - 189) / 180 + 1 : 0;
    u64 r_lo = (n >= 189 && n - 189 >= e) ? (n - 189 - e) / 180 + 1 : 0;
    u64 r_hi = (n >= 189 && n - 189 >= a) ? (n - 189 - a) / 180 + 1 : 0;
    if (r_hi > Lc) r_hi = Lc;
    if (a == 0 && e > 0) {  // the blocks at the record's start
        if (r_hi <= r_lo) r_lo = Lc;
        r_hi = ~0ull;
    }
    std::vector<gts_segment> v;
    u64 run_base = 0;
    for (int f = 0; f < 6; f++) {
        if (!((mask >> f) & 1u)) continue;
        const u64 Lf = L[f];
        const u64 body = run_base + H + 3;
        if (p.first) v.push_back({run_base, H + 3});
        run_base += host_run_size(H, Lf);
        const u64 NL = (Lf + 59) / 60;
        u64 lo = f < 3 ? f_lo : r_lo, hi = f < 3 ? f_hi : r_hi;
        if (hi > NL) hi = NL;
        if (hi <= lo) continue;
        const u64 bytes = 61 * (hi - lo) - (hi == NL ? 60 * NL - Lf : 0);
        v.push_back({body + 61 * lo, bytes});
    }
    *total = run_base;
    std::sort(v.begin(), v.end(), [](const gts_segment &x, const gts_segment &y) { return x.offset < y.offset; });
    int m = 0;
    for (const gts_segment &s : v) {
        if (s.length == 0) continue;
        if (m && segs[m - 1].offset + segs[m - 1].length == s.offset) segs[m - 1].length += s.length;
        else segs[m++] = s;
    }
    return m;
}

DevCtx *dev0(gts_ctx *ctx) { return ctx->devs[0].get(); }

}  // namespace

extern "C" {

int gts_frame_mask(const char *frame, uint32_t *mask, int *reverse) {
    uint32_t m = 0;
    bool r = false;
    if (!gts::frame_mask(frame, &m, &r)) return GTS_ERR_FRAME;
    if (mask) *mask = m;
    if (reverse) *reverse = r ? 1 : 0;
    return GTS_OK;
}

int gts_codon_table(int32_t table, char aa64[64]) {
    const char *aa = gts::find_codon_table(table);
    if (!aa) return GTS_ERR_TABLE;
    std::memcpy(aa64, aa, 64);
    return GTS_OK;
}

int gts_fused_lut(int32_t table, int clean, uint16_t lut[125]) {
    const char *aa = gts::find_codon_table(table);
    if (!aa) return GTS_ERR_TABLE;
    gts::build_fused_lut(aa, clean != 0, lut);
    return GTS_OK;
}

int gts_create(const gts_options *opt, gts_ctx **out) {
    if (!out) return fail(nullptr, GTS_ERR_ARG, "gts_create: out is NULL");
    *out = nullptr;
    if (!opt) return fail(nullptr, GTS_ERR_ARG, "gts_create: options are NULL");
    uint32_t mask = 0;
    bool reverse = false;
    const char *frame = opt->frame ? opt->frame : "";
    // same texts as transeq.go:107 and ncbicode/code.go:378
    if (!gts::frame_mask(frame, &mask, &reverse))
        return fail(nullptr, GTS_ERR_FRAME, std::string("wrong value for -f | --frame parameter: ") + frame);
    const char *aa = gts::find_codon_table(opt->table);
    if (!aa) return fail(nullptr, GTS_ERR_TABLE, "invalid table code: " + std::to_string(opt->table));
    if (opt->num_gpus > 8) return fail(nullptr, GTS_ERR_ARG, "num_gpus must be at most 8");

    std::unique_ptr<gts_ctx> ctx(new gts_ctx());
    ctx->opt = *opt;
    ctx->frame = frame;
    ctx->opt.frame = ctx->frame.c_str();
    ctx->mask = mask;
    ctx->reverse = reverse;
    std::memset(&ctx->lut, 0, sizeof(ctx->lut));
    gts::build_fused_lut(aa, opt->clean != 0, ctx->lut.v);
    {   // k_lines' tables: one residue per byte, forward and (read downwards) reverse strand
        uint8_t *bf = reinterpret_cast<uint8_t *>(ctx->lutb.f), *br = reinterpret_cast<uint8_t *>(ctx->lutb.r);
        for (int e = 0; e < 128; e++) {
            const int x = e < 125 ? e : 0;
            const int s0 = x % 5, s1 = (x / 5) % 5, s2 = x / 25;
            bf[e] = (uint8_t)(ctx->lut.v[x] & 0xFF);
            br[e] = (uint8_t)(ctx->lut.v[s2 + 5 * s1 + 25 * s0] >> 8);
        }
    }
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    if (opt->chunk_bytes) ctx->chunk = std::max<size_t>((size_t)opt->chunk_bytes, 4096);

    // ---- devices (no CPU fallback: without a usable sm_100 device creation fails) ----
    std::vector<int> ids;
    if (opt->num_gpus == 0) {
        ids.push_back(opt->device);
    } else if (opt->num_gpus > 0) {
        for (int i = 0; i < opt->num_gpus; i++) ids.push_back(opt->device_ids[i]);
    } else {
        int count = 0;
        cudaError_t e = cudaGetDeviceCount(&count);
        if (e != cudaSuccess || count <= 0)
            return fail(nullptr, GTS_ERR_CUDA, std::string("CUDA initialisation failed (no CPU fallback): ") +
                                                   (e != cudaSuccess ? cudaGetErrorString(e) : "no device"));
        for (int i = 0; i < count && i < 8; i++) ids.push_back(i);
    }
    DeviceGuard guard;
    const bool keep_current = ids.size() > 1;  // several GPUs: the caller's current device stays what it was
    for (int id : ids) {
        std::unique_ptr<DevCtx> dev;
        std::string emsg;
        int rc = dev_create(id, &dev, &emsg);
        if (rc) {
            if (dev) dev_destroy(dev.get());
            for (auto &d : ctx->devs) dev_destroy(d.get());
            return fail(nullptr, rc, "CUDA initialisation failed (no CPU fallback): " + emsg);
        }
        ctx->devs.push_back(std::move(dev));
    }
    if (!keep_current) guard.prev = ctx->devs[0]->device;  // one GPU: it becomes current (as cudaSetDevice(device) would)
    ctx->in_max = g_slots * ctx->devs.size() + 2;
    ctx->out_max = ctx->in_max + 1;
    ctx->any_order = opt->any_order != 0;
    ctx->poison = std::getenv("GTS_POISON_OUTPUT") != nullptr;
    ctx->short_on = std::getenv("GTS_NO_SHORT") == nullptr;
    *out = ctx.release();
    return GTS_OK;
}

void gts_destroy(gts_ctx *ctx) {
    if (!ctx) return;
    DeviceGuard guard;
    if (ctx->workers_started) {
        {
            std::unique_lock<std::mutex> lk(ctx->mu);
            ctx->cancelled = true;
            drain_and_clear_locked(ctx, lk);
            ctx->exiting = true;
        }
        ctx->cv.notify_all();
        for (auto &d : ctx->devs) {
            d->issuer.join();
            d->retirer.join();
        }
    }
    for (auto &d : ctx->devs) dev_destroy(d.get());
    if (ctx->fill) ctx->in_free.push_back(ctx->fill);
    for (HostBuf *b : ctx->in_free) { if (b->p) cudaFreeHost(b->p); delete b; }
    for (HostBuf *b : ctx->out_free) { if (b->p) cudaFreeHost(b->p); delete b; }
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_piece) cudaFreeHost(ctx->h_piece);
    delete ctx;
}

const char *gts_last_error(const gts_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gts_num_devices(const gts_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

// ---- device-resident entry points (first device) ------------------------------------------

int gts_translate_device_async(gts_ctx *ctx, const void *d_in, size_t n, void *d_out, size_t cap, void *stream) {
    if (!ctx) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    PassArgs a;
    a.d_in = static_cast<const uint8_t *>(d_in); a.n = n; a.d_out = static_cast<uint8_t *>(d_out); a.cap = cap;
    a.stream = static_cast<cudaStream_t>(stream);
    std::string emsg;
    std::lock_guard<std::mutex> dl(dev->mu);
    int rc = enqueue_pass(ctx, dev, a, &emsg);
    if (rc) return fail(ctx, rc, emsg);
    dev->pending = true;
    return GTS_OK;
}

int gts_device_result(gts_ctx *ctx, size_t *out_n) {
    if (!ctx) return GTS_ERR_ARG;
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    int rc = finish_pass(ctx, dev, false);
    if (rc) return rc;
    if (out_n) *out_n = dev->h_totals->out_total;
    if (dev->h_totals->flags & gts::kFlagOutOverflow)
        return fail(ctx, GTS_ERR_CAPACITY, "output buffer too small: need " +
                                               std::to_string(dev->h_totals->out_total) + " bytes");
    return GTS_OK;
}

int gts_translate_device(gts_ctx *ctx, const void *d_in, size_t n, void *d_out, size_t cap, size_t *out_n,
                         void *stream) {
    int rc = gts_translate_device_async(ctx, d_in, n, d_out, cap, stream);
    if (rc) return rc;
    return gts_device_result(ctx, out_n);
}

size_t gts_output_bound_hint(const gts_ctx *ctx, size_t n) {
    // long 6-frame records produce ~2.04 output bytes per input byte; short records are dominated
    // by the six copies of the header.  This is a hint, not a bound: gts_translate_device reports
    // the exact size when the buffer is too small.
    (void)ctx;
    return out_estimate(n);
}

// ---- one-shot host call: the chunk engine in direct mode ------------------------------------

int gts_translate_buffer(gts_ctx *ctx, const uint8_t *in, size_t n, const uint8_t **out, size_t *out_n) {
    if (!ctx) return GTS_ERR_ARG;
    if ((n && !in) || !out || !out_n) return fail(ctx, GTS_ERR_ARG, "NULL argument");
    *out = nullptr;
    *out_n = 0;
    start_workers(ctx);
    {
        std::unique_lock<std::mutex> lk(ctx->mu);
        if (!ctx->window.empty() || ctx->fill_n || ctx->eof)
            return fail(ctx, GTS_ERR_ARG, "gts_translate_buffer: a stream is in progress on this context");
        int rc = stream_error_locked(ctx);
        if (rc) return rc;
    }
    // a first guess of the output size (grown on demand: the chunks report their exact sizes)
    {
        std::string emsg;
        const size_t guess = n * 2 + n / 4 + ((size_t)1 << 20);
        if (guess > ctx->h_out_cap) {
            int rc = ensure_pinned(&ctx->h_out, &ctx->h_out_cap, guess, 0, &emsg);
            if (rc) return fail(ctx, rc, emsg);
        }
    }
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->direct = true;
    ctx->direct_turn = ctx->next_seq;
    ctx->direct_off = 0;
    ctx->direct_active = 0;
    ctx->direct_growing = false;
    ctx->trunc_seq = ~0ull;
    const uint64_t first_seq = ctx->next_seq;
    // ---- cut at record boundaries and deal the chunks out ----
    const size_t chunk = ctx->chunk;
    size_t pos = 0;
    uint64_t records = 0, nucleotides = 0;
    do {
        size_t end = n;
        if (n - pos > chunk + chunk / 2) {
            end = gts_host::boundary_below(in, pos, pos + chunk);
            if (end == 0) end = gts_host::boundary_above(in, pos + chunk, n);  // one record larger than a chunk
        }
        XJob *job = new XJob();
        job->in = in + pos;
        job->n = end - pos;
        job->last = end == n;
        dispatch_locked(ctx, job);
        pos = end;
        // (bounded backlog: the window holds at most a few chunks per GPU)
        ctx->cv.wait(lk, [&] {
            return stream_dead(ctx) || ctx->trunc_seq != ~0ull || ctx->next_seq - ctx->direct_turn < 4 * ctx->devs.size() + 4;
        });
    } while (pos < n && !stream_dead(ctx) && ctx->trunc_seq == ~0ull);
    ctx->cv.wait(lk, [&] {
        for (XJob *j : ctx->window)
            if (!j->done) return false;
        return true;
    });
    int rc = stream_error_locked(ctx);
    size_t total = ctx->direct_off;
    ctx->last_truncated = ctx->trunc_seq != ~0ull;
    std::vector<XJob *> jobs(ctx->window.begin(), ctx->window.end());
    ctx->window.clear();
    ctx->consume_seq = ctx->next_seq;
    ctx->direct = false;
    lk.unlock();
    if (g_trace) {
        const double t0 = jobs.empty() ? 0 : jobs[0]->t_disp, t_end = now_ms();
        for (XJob *j : jobs)
            std::fprintf(stderr, "[gts trace] chunk %llu dev %d n %zu out %zu: dispatched %.3f issue %.3f..%.3f kernels done %.3f d2h issued %.3f done %.3f\n",
                         (unsigned long long)j->seq, j->dev, j->n, j->out_n, j->t_disp - t0, j->t_issue0 - t0, j->t_issue1 - t0,
                         j->t_comp - t0, j->t_d2h0 - t0, j->t_d2h1 - t0);
        std::fprintf(stderr, "[gts trace] call returns at %.3f\n", t_end - t0);
    }
    for (XJob *j : jobs) {
        if (rc == GTS_OK && j->seq <= ctx->trunc_seq) {
            fire_warnings(ctx, j);
            records += j->records;
            nucleotides += j->nucleotides;
        }
    }
    lk.lock();
    for (XJob *j : jobs) recycle_job_locked(ctx, j);
    ctx->trunc_seq = ~0ull;
    (void)first_seq;
    if (rc) return rc;
    ctx->stats.records = records;
    ctx->stats.nucleotides = nucleotides;
    ctx->stats.in_bytes = n;
    ctx->stats.out_bytes = total;
    *out = ctx->h_out;
    *out_n = total;
    return GTS_OK;
}

// ---- streaming host interface ----------------------------------------------------------

int gts_acquire_input(gts_ctx *ctx, uint8_t **buf, size_t *cap) {
    if (!ctx || !buf || !cap) return GTS_ERR_ARG;
    start_workers(ctx);
    std::unique_lock<std::mutex> lk(ctx->mu);
    int rc = stream_error_locked(ctx);
    if (rc) return rc;
    if (ctx->eof) return fail(ctx, GTS_ERR_ARG, "gts_acquire_input after eof");
    const size_t chunk = ctx->chunk;
    const size_t min_read = std::max<size_t>(chunk / 8, 4096);
    std::string emsg;
    if (!ctx->fill) {
        HostBuf *b = pool_take(ctx, lk, ctx->in_free, ctx->in_count, ctx->in_max, chunk + chunk / 4 + 65536, &emsg, &rc);
        if (!b) return rc ? fail(ctx, rc, emsg) : stream_error_locked(ctx);
        ctx->fill = b;
        ctx->fill_n = 0;
        ctx->scan_from = 1;
        ctx->last_boundary = 0;
    }
    if (ctx->fill->cap - ctx->fill_n < min_read) {
        // one record larger than the buffer: grow geometrically (the bytes read so far move along)
        lk.unlock();
        rc = ensure_pinned(&ctx->fill->p, &ctx->fill->cap, 2 * ctx->fill->cap, ctx->fill_n, &emsg);
        lk.lock();
        if (rc) return fail(ctx, rc, emsg);
    }
    *buf = ctx->fill->p + ctx->fill_n;
    *cap = ctx->fill->cap - ctx->fill_n;
    return GTS_OK;
}

int gts_submit(gts_ctx *ctx, size_t nbytes, int eof) {
    if (!ctx) return GTS_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    int rc = stream_error_locked(ctx);
    if (rc) return rc;
    if (ctx->eof) return fail(ctx, GTS_ERR_ARG, "gts_submit after eof");
    if (!ctx->fill) {
        if (nbytes) return fail(ctx, GTS_ERR_ARG, "gts_submit without gts_acquire_input");
        if (!eof) return GTS_OK;
        lk.unlock();
        uint8_t *b = nullptr;
        size_t c = 0;
        rc = gts_acquire_input(ctx, &b, &c);
        if (rc) return rc;
        lk.lock();
    }
    HostBuf *cur = ctx->fill;
    if (ctx->fill_n + nbytes > cur->cap) return fail(ctx, GTS_ERR_ARG, "gts_submit: more bytes than acquired");
    const size_t old_fill = ctx->fill_n;
    ctx->fill_n += nbytes;
    // last record boundary among the new bytes: a '>' that starts a line (transeq.go:198); every
    // byte is looked at once, however often a large record makes the buffer grow
    {
        const size_t from = std::max<size_t>(std::max<size_t>(ctx->scan_from, old_fill), 1);
        if (ctx->fill_n > from) {
            const size_t b = gts_host::boundary_below(cur->p, from - 1, ctx->fill_n);
            if (b) ctx->last_boundary = b;
        }
        ctx->scan_from = ctx->fill_n;
    }
    size_t cut = 0;
    if (eof) {
        ctx->eof = true;
        cut = ctx->fill_n;
        if (cut == 0 && !ctx->first_job) {  // nothing left behind the last record boundary
            ctx->in_free.push_back(cur);
            ctx->fill = nullptr;
            ctx->cv.notify_all();
            return GTS_OK;
        }
    } else {
        if (ctx->fill_n < ctx->chunk || ctx->last_boundary == 0) return GTS_OK;  // keep filling
        cut = ctx->last_boundary;
    }
    // the unfinished record behind the cut moves to the front of a fresh buffer
    const size_t rest = ctx->fill_n - cut;
    HostBuf *next = nullptr;
    if (!eof) {
        std::string emsg;
        const size_t need = std::max(ctx->chunk + ctx->chunk / 4 + 65536, rest + ctx->chunk / 4);
        next = pool_take(ctx, lk, ctx->in_free, ctx->in_count, ctx->in_max, need, &emsg, &rc);
        if (!next) return rc ? fail(ctx, rc, emsg) : stream_error_locked(ctx);
        if (rest) std::memcpy(next->p, cur->p + cut, rest);
    }
    XJob *job = new XJob();
    job->in = cur->p;
    job->n = cut;
    job->last = eof != 0;
    job->inbuf = cur;
    if (ctx->first_job) ctx->stats.records = ctx->stats.nucleotides = ctx->stats.in_bytes = ctx->stats.out_bytes = 0;
    ctx->first_job = false;
    ctx->fill = next;
    ctx->fill_n = next ? rest : 0;
    ctx->scan_from = 1;  // (the carried bytes hold no boundary: position 0 does not count)
    ctx->last_boundary = 0;
    if (next && rest > 1) ctx->scan_from = rest;
    dispatch_locked(ctx, job);
    return GTS_OK;
}

int gts_next_output(gts_ctx *ctx, const uint8_t **buf, size_t *n) {
    if (!ctx || !buf || !n) return GTS_ERR_ARG;
    return next_output(ctx, buf, n, false);
}

int gts_wait_output(gts_ctx *ctx, const uint8_t **buf, size_t *n) {
    if (!ctx || !buf || !n) return GTS_ERR_ARG;
    return next_output(ctx, buf, n, true);
}

int gts_release_output(gts_ctx *ctx) {
    if (!ctx) return GTS_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    XJob *job = ctx->handed;
    if (!job) return GTS_OK;
    ctx->handed = nullptr;
    auto it = std::find(ctx->window.begin(), ctx->window.end(), job);  // (the front, unless any_order)
    if (it != ctx->window.end()) {
        ctx->window.erase(it);
        ctx->consume_seq++;
        recycle_job_locked(ctx, job);
    }
    ctx->cv.notify_all();
    return GTS_OK;
}

int gts_report_write_error(gts_ctx *ctx, const char *msg) {
    if (!ctx) return GTS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    // writer.go:175: fmt.Errorf("fail to write to output file: %v", err); the first error wins
    set_stream_error_locked(ctx, GTS_ERR_WRITE, std::string("fail to write to output file: ") + (msg ? msg : ""));
    ctx->cancelled = true;
    ctx->cv.notify_all();
    return GTS_OK;
}

int gts_cancel(gts_ctx *ctx) {
    if (!ctx) return GTS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    ctx->cancelled = true;
    ctx->cv.notify_all();
    return GTS_OK;
}

int gts_stream_reset(gts_ctx *ctx) {
    if (!ctx) return GTS_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->mu);
    ctx->cancelled = true;
    drain_and_clear_locked(ctx, lk);
    reset_stream_locked(ctx);
    ctx->cancelled = false;
    ctx->e_rc = 0;
    ctx->e_err.clear();
    ctx->direct = false;
    ctx->cv.notify_all();
    return GTS_OK;
}

int gts_set_warning_callback(gts_ctx *ctx, gts_warning_fn fn, void *user) {
    if (!ctx) return GTS_ERR_ARG;
    std::lock_guard<std::mutex> lk(ctx->mu);
    if (!ctx->window.empty()) return fail(ctx, GTS_ERR_ARG, "gts_set_warning_callback: a stream is in progress");
    ctx->warn_fn = fn;
    ctx->warn_user = user;
    return GTS_OK;
}

size_t gts_format_warning(const uint8_t *header, size_t header_len, uint8_t byte, uint64_t pos, char *dst, size_t cap) {
    // encodedSequence.go:39: fmt.Printf("WARNING: invalid char in sequence %s: '%s' ( pos %d), replacing
    // with 'N'\n", string(s[:headerSize]), string(n), i) -- s[:headerSize] starts with the four raw
    // little-endian bytes of headerSize (= header_len + 4); string(n) of a byte is the rune's UTF-8 form
    std::string s = "WARNING: invalid char in sequence ";
    const uint32_t hs = (uint32_t)(header_len + 4);
    for (int i = 0; i < 4; i++) s.push_back((char)((hs >> (8 * i)) & 0xFF));
    s.append(reinterpret_cast<const char *>(header), header_len);
    s += ": '";
    if (byte < 0x80) {
        s.push_back((char)byte);
    } else {
        s.push_back((char)(0xC0 | (byte >> 6)));
        s.push_back((char)(0x80 | (byte & 0x3F)));
    }
    s += "' ( pos " + std::to_string(pos) + "), replacing with 'N'\n";
    if (dst && cap) {
        const size_t m = std::min(s.size(), cap);
        std::memcpy(dst, s.data(), m);
    }
    return s.size();
}

int gts_set_max_line_length(gts_ctx *ctx, size_t n) {
    if (!ctx) return GTS_ERR_ARG;
    if (n < 2 * (size_t)gts::kParseTile) return fail(ctx, GTS_ERR_ARG, "line limit must be at least 16384 bytes");
    ctx->max_line = n;
    return GTS_OK;
}

int gts_set_profiling(gts_ctx *ctx, int enable) {
    if (!ctx) return GTS_ERR_ARG;
    ctx->profiling = enable != 0;
    return GTS_OK;
}

int gts_kernel_times(gts_ctx *ctx, double ms[4], uint64_t *passes) {
    if (!ctx || !ms) return GTS_ERR_ARG;
    for (int k = 0; k < 4; k++) ms[k] = 0.0;
    size_t np_all = 0;
    DeviceGuard guard;
    for (auto &d : ctx->devs) {
        CU(cudaSetDevice(d->device));
        const size_t np = d->prof_used / 5;
        for (size_t p = 0; p < np; p++) {
            cudaEvent_t *ev = d->prof_events.data() + 5 * p;
            CU(cudaEventSynchronize(ev[4]));
            for (int k = 0; k < 4; k++) {
                float t = 0.f;
                CU(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
                ms[k] += t;
            }
        }
        d->prof_used = 0;
        np_all += np;
    }
    if (passes) *passes = np_all;
    return GTS_OK;
}

// ---- one record split over several GPUs ---------------------------------------------------

// untrimmed residues of each frame of a record with n nucleotides (writer.go:97-112)
int gts_frame_lengths(const gts_ctx *ctx, uint64_t n, uint64_t len[6]) {
    if (!ctx || !len) return GTS_ERR_ARG;
    u64 L[6];
    host_frame_lengths(n, ctx->opt.alternative != 0, L);
    for (int f = 0; f < 6; f++) len[f] = ((ctx->mask >> f) & 1u) ? L[f] : 0;
    return GTS_OK;
}

int gts_piece_scan_device(gts_ctx *ctx, const void *d_in, size_t n, int first, gts_piece_info *info, void *stream) {
    if (!ctx || !info) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    ctx->piece_on = true;
    ctx->piece_job = gts::Job();
    ctx->piece_job.pv.no_end_pads = 1;  // count only: no pass of its own end
    if (!dev->d_head) CU(cudaMalloc(&dev->d_head, sizeof(gts::PieceInfoDev)));
    ctx->piece_job.info_out = reinterpret_cast<gts::PieceInfoDev *>(dev->d_head);
    ctx->last_truncated = false;
    PassArgs a;
    a.d_in = static_cast<const uint8_t *>(d_in); a.n = n; a.stream = static_cast<cudaStream_t>(stream);
    a.scan_only = true;  // k_parse + k_tile_scan: the counts, the head codes, the header line's length
    std::string emsg;
    int rc = enqueue_pass(ctx, dev, a, &emsg);
    if (rc) fail(ctx, rc, emsg);
    gts::PieceInfoDev inf;
    if (!rc) {
        // (the piece's summary travels with the totals: one synchronisation for the whole scan)
        cudaError_t ce = cudaMemcpyAsync(&inf, dev->d_head, sizeof(inf), cudaMemcpyDeviceToHost, a.stream);
        if (ce != cudaSuccess) rc = cuda_fail(ctx, ce, "copy of the piece summary");
    }
    if (!rc) {
        dev->pending = true;
        rc = finish_pass(ctx, dev, true);
    }
    ctx->piece_on = false;
    if (rc) return rc;
    const gts::Totals &t = *dev->h_totals;
    if (ctx->last_truncated) return fail(ctx, GTS_ERR_ARG, "a split record cannot hold an over-long line");
    info->headers = inf.headers;
    info->nucleotides = inf.nucleotides;
    info->tail[0] = (uint8_t)(inf.tail & 0xF);
    info->tail[1] = (uint8_t)((inf.tail >> 4) & 0xF);
    info->reserved[0] = info->reserved[1] = 0;
    info->header_len = 0;
    std::memcpy(info->head, inf.head, sizeof(info->head));
    if (first) {
        if (t.n_headers != 1) return fail(ctx, GTS_ERR_ARG, "the first piece must hold exactly one header line");
        if (inf.header_len == ~0ull) return fail(ctx, GTS_ERR_ARG, "the first piece must start with the header line");
        info->header_len = (uint32_t)inf.header_len;
    } else if (t.n_headers != 0) {
        return fail(ctx, GTS_ERR_ARG, "only the first piece of a split record may hold a header line");
    }
    dev->slots_of = d_in;  // the translate pass of the same buffer reuses the parse
    dev->slots_n = n;
    return GTS_OK;
}

// The piece fields of a pass from the caller's description of the piece.
static void fill_piece_job(gts::Job &j, const gts_piece *piece) {
    j = gts::Job();
    gts::PieceParams &p = j.pv;
    p.piece = 1;
    p.no_end_pads = piece->last ? 0 : 1;
    p.ov_last = piece->last ? 1 : 0;
    p.no_headers = piece->first ? 0 : 1;
    p.ov_H = piece->header_len;
    p.ov_n = piece->nt_total;
    p.ov_dbase = 4;  // stream of the whole record: slot 0's pads, the record's pads, its nucleotides
    std::memcpy(p.head, piece->next_head, sizeof(p.head));
    for (uint8_t &c : p.head) c = c > 4 ? 0 : c;
    if (piece->first) {
        p.ov_slot = 1;
    } else {
        // the buffer's first two symbols stand in for the two nucleotides before the piece
        p.ov_slot = 0;
        p.has_carry = 1;
        p.carry = (u32)(piece->carry[0] & 0xF) | (u32)(piece->carry[1] & 0xF) << 4;
        p.t0_base = piece->nt_before + 2;
        p.win_lo = p.t0_base + 2;
    }
    if (piece->has_trim_len) {
        j.ov_trim = 1;
        for (int f = 0; f < 6; f++) j.ov_len[f] = (u32)piece->trim_len[f];
    }
}

int gts_piece_trim_device(gts_ctx *ctx, const void *d_in, size_t n, const gts_piece *piece, uint64_t len[6],
                          uint8_t resolved[6], void *stream) {
    if (!ctx || !piece || !len || !resolved) return GTS_ERR_ARG;
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    if (dev->slots_of != d_in || dev->slots_n != n)
        return fail(ctx, GTS_ERR_ARG, "gts_piece_trim: the piece must have been scanned by the call before (gts_piece_scan)");
    if (!dev->d_pt) CU(cudaMalloc(&dev->d_pt, 12 * sizeof(u64)));
    gts::Job job = dev->job;  // the scan pass's tables: slots, tile records, tile offsets, totals
    gts::Job p;
    fill_piece_job(p, piece);
    job.pv = p.pv;
    job.pp = nullptr;
    job.pt_before = piece->nt_before;
    job.pt_out = dev->d_pt;
    for (int f = 0; f < 6; f++) job.pt_len[f] = len[f];
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    CU(gts::launch_piece_trim(job, st));
    u64 out[12];
    CU(cudaMemcpyAsync(out, dev->d_pt, sizeof(out), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    for (int f = 0; f < 6; f++) {
        len[f] = out[f];
        resolved[f] = (uint8_t)out[6 + f];
    }
    ctx->stats.kernel_launches++;
    return GTS_OK;
}

int gts_piece_translate_device(gts_ctx *ctx, const void *d_in, size_t n, const gts_piece *piece, void *d_out,
                               size_t cap, size_t *out_total, gts_segment segs[GTS_MAX_SEGMENTS], int *nseg,
                               void *stream) {
    if (!ctx || !piece || !segs || !nseg) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    if (ctx->opt.trim && !piece->has_trim_len)
        return fail(ctx, GTS_ERR_ARG, "--trim on a split record needs the trimmed frame lengths (gts_piece_trim)");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    fill_piece_job(ctx->piece_job, piece);
    ctx->piece_on = true;
    PassArgs a;
    a.d_in = static_cast<const uint8_t *>(d_in); a.n = n; a.d_out = static_cast<uint8_t *>(d_out); a.cap = cap;
    a.stream = static_cast<cudaStream_t>(stream);
    a.reuse_slots = dev->slots_of == d_in && dev->slots_n == n;  // parsed by gts_piece_scan: no second k_parse
    std::string emsg;
    int rc = enqueue_pass(ctx, dev, a, &emsg);
    if (rc) fail(ctx, rc, emsg);
    if (!rc) {
        dev->pending = true;
        rc = finish_pass(ctx, dev, false);
    }
    ctx->piece_on = false;
    if (rc) return rc;
    const gts::Totals &t = *dev->h_totals;
    if (out_total) *out_total = t.out_total;
    if (t.flags & gts::kFlagOutOverflow)
        return fail(ctx, GTS_ERR_CAPACITY, "output buffer too small: need " + std::to_string(t.out_total) + " bytes");
    const u64 n_piece = t.total_sym - 2 - (piece->first ? 2 : 0);
    u64 total = 0;
    *nseg = piece_segments(ctx->mask, ctx->opt.alternative != 0, *piece, n_piece, segs, &total);
    if (total != t.out_total) return fail(ctx, GTS_ERR_ARG, "split record: host and device output sizes differ");
    return GTS_OK;
}

// ---- the same step without a host round trip between scan and translate ---------------------
// scan (k_parse + k_tile_scan) -> the caller all-gathers the 256-byte summaries on the same stream
// (NCCL) -> plan on the device (k_piece_plan) + the translate kernels on the reused slots -> one
// synchronisation at the end (gts_piece_result).

int gts_piece_scan_async(gts_ctx *ctx, const void *d_in, size_t n, void *d_info, void *stream) {
    if (!ctx || !d_info) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    ctx->piece_on = true;
    ctx->piece_job = gts::Job();
    ctx->piece_job.pv.no_end_pads = 1;
    ctx->piece_job.info_out = static_cast<gts::PieceInfoDev *>(d_info);
    PassArgs a;
    a.d_in = static_cast<const uint8_t *>(d_in); a.n = n; a.stream = static_cast<cudaStream_t>(stream);
    a.scan_only = true;
    std::string emsg;
    int rc = enqueue_pass(ctx, dev, a, &emsg);
    ctx->piece_on = false;
    if (rc) return fail(ctx, rc, emsg);
    dev->slots_of = d_in;  // (valid once the stream has run the pass; the translate call is enqueued behind it)
    dev->slots_n = n;
    return GTS_OK;
}

int gts_piece_translate_async(gts_ctx *ctx, const void *d_in, size_t n, int rank, int world, const void *d_infos,
                              void *d_out, size_t cap, void *stream) {
    if (!ctx || !d_infos || rank < 0 || world < 1 || rank >= world || world > 64) return GTS_ERR_ARG;
    if (ctx->opt.trim) return fail(ctx, GTS_ERR_ARG, "--trim on a split record goes through gts_piece_scan / gts_piece_trim / gts_piece_translate");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    if (dev->slots_of != d_in || dev->slots_n != n)
        return fail(ctx, GTS_ERR_ARG, "gts_piece_translate_async: the piece must have been scanned by the call before");
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (!dev->d_pp) CU(cudaMalloc(&dev->d_pp, sizeof(gts::PieceParams)));
    if (!dev->h_infos) CU(cudaHostAlloc(&dev->h_infos, 64 * sizeof(gts::PieceInfoDev), cudaHostAllocPortable));
    CU(gts::launch_piece_plan(static_cast<const gts::PieceInfoDev *>(d_infos), rank, world, dev->d_pp, st));
    CU(cudaMemcpyAsync(dev->h_infos, d_infos, (size_t)world * sizeof(gts::PieceInfoDev), cudaMemcpyDeviceToHost, st));
    dev->piece_rank = rank;
    dev->piece_world = world;
    ctx->piece_on = true;
    ctx->piece_job = gts::Job();
    ctx->piece_job.pp = dev->d_pp;
    PassArgs a;
    a.d_in = static_cast<const uint8_t *>(d_in); a.n = n; a.d_out = static_cast<uint8_t *>(d_out); a.cap = cap;
    a.stream = st;
    a.reuse_slots = true;
    std::string emsg;
    int rc = enqueue_pass(ctx, dev, a, &emsg);
    ctx->piece_on = false;
    if (rc) return fail(ctx, rc, emsg);
    dev->pending = true;
    ctx->stats.kernel_launches++;
    return GTS_OK;
}

int gts_piece_result(gts_ctx *ctx, size_t *out_total, gts_segment segs[GTS_MAX_SEGMENTS], int *nseg, gts_piece *piece_out) {
    if (!ctx || !segs || !nseg) return GTS_ERR_ARG;
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::lock_guard<std::mutex> dl(dev->mu);
    if (!dev->pending || !dev->h_infos) return fail(ctx, GTS_ERR_ARG, "gts_piece_result: no piece enqueued");
    CU(cudaStreamSynchronize(dev->job_stream));
    dev->pending = false;
    const gts::Totals &t = *dev->h_totals;
    if (t.flags & (gts::kFlagRecOverflow | gts::kFlagLongLine))
        return fail(ctx, GTS_ERR_ARG, "split record: the piece needs the host-driven calls (record table / line length)");
    // the same plan on the host, for the segments
    const gts::PieceInfoDev *all = dev->h_infos;
    const int rank = dev->piece_rank, world = dev->piece_world;
    for (int k = 0; k < world; k++) {
        if (all[k].headers != (k == 0 ? 1u : 0u) || (k == 0 && all[0].header_len == ~0ull))
            return fail(ctx, GTS_ERR_ARG, "a split record is ONE record: the first piece starts with its header line, the others hold none");
    }
    gts_piece p;
    std::memset(&p, 0, sizeof(p));
    uint64_t total_nt = 0, before = 0;
    for (int k = 0; k < world; k++) {
        if (k < rank) before += all[k].nucleotides;
        total_nt += all[k].nucleotides;
    }
    p.nt_before = before;
    p.nt_total = total_nt;
    p.header_len = (uint32_t)all[0].header_len;
    p.first = rank == 0;
    p.last = rank == world - 1;
    if (piece_out) *piece_out = p;
    if (out_total) *out_total = t.out_total;
    if (t.flags & gts::kFlagOutOverflow)
        return fail(ctx, GTS_ERR_CAPACITY, "output buffer too small: need " + std::to_string(t.out_total) + " bytes");
    u64 total = 0;
    *nseg = piece_segments(ctx->mask, ctx->opt.alternative != 0, p, all[rank].nucleotides, segs, &total);
    if (total != t.out_total) return fail(ctx, GTS_ERR_ARG, "split record: host and device output sizes differ");
    return GTS_OK;
}

int gts_piece_scan(gts_ctx *ctx, const uint8_t *in, size_t n, int first, gts_piece_info *info) {
    if (!ctx || !info) return GTS_ERR_ARG;
    if (n && !in) return fail(ctx, GTS_ERR_ARG, "NULL argument");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    std::string emsg;
    int rc = ensure_dev_buffer(&dev->d_in, &dev->d_in_cap, n + 64, &emsg);
    if (rc) return fail(ctx, rc, emsg);
    if (n) CU(cudaMemcpyAsync(dev->d_in, in, n, cudaMemcpyHostToDevice, dev->stream));
    ctx->stats.h2d_bytes += n;
    ctx->piece_n = n;
    ctx->piece_first = first != 0;
    return gts_piece_scan_device(ctx, dev->d_in, n, first, info, dev->stream);
}

int gts_piece_translate(gts_ctx *ctx, const gts_piece *piece, size_t *out_total, gts_segment segs[GTS_MAX_SEGMENTS],
                        const uint8_t *data[GTS_MAX_SEGMENTS], int *nseg) {
    if (!ctx || !piece || !segs || !data || !nseg) return GTS_ERR_ARG;
    if ((piece->first != 0) != ctx->piece_first) return fail(ctx, GTS_ERR_ARG, "gts_piece_translate: not the piece last scanned");
    DevCtx *dev = dev0(ctx);
    CU(cudaSetDevice(dev->device));
    // the whole record's output size is known from the counts alone
    u64 L[6], total = 0;
    host_frame_lengths(piece->nt_total, ctx->opt.alternative != 0, L);
    if (piece->has_trim_len)
        for (int f = 0; f < 6; f++) L[f] = piece->trim_len[f];
    for (int f = 0; f < 6; f++)
        if ((ctx->mask >> f) & 1u) total += host_run_size(piece->header_len, L[f]);
    std::string emsg;
    int rc = ensure_dev_buffer(&dev->d_out, &dev->d_out_cap, total + 64, &emsg);
    if (rc) return fail(ctx, rc, emsg);
    if (ctx->poison && total) CU(cudaMemsetAsync(dev->d_out, 0xEE, total, dev->stream));
    size_t tot = 0;
    rc = gts_piece_translate_device(ctx, dev->d_in, ctx->piece_n, piece, dev->d_out, dev->d_out_cap, &tot, segs, nseg,
                                    dev->stream);
    if (rc) return rc;
    size_t bytes = 0;
    for (int i = 0; i < *nseg; i++) bytes += round_up(segs[i].length, 16);
    rc = ensure_pinned(&ctx->h_piece, &ctx->h_piece_cap, bytes + 64, 0, &emsg);
    if (rc) return fail(ctx, rc, emsg);
    size_t off = 0;
    for (int i = 0; i < *nseg; i++) {
        CU(cudaMemcpyAsync(ctx->h_piece + off, dev->d_out + segs[i].offset, segs[i].length, cudaMemcpyDeviceToHost, dev->stream));
        data[i] = ctx->h_piece + off;
        off += round_up(segs[i].length, 16);
        ctx->stats.d2h_bytes += segs[i].length;
    }
    CU(cudaStreamSynchronize(dev->stream));
    if (out_total) *out_total = tot;
    return GTS_OK;
}

int gts_piece_trim(gts_ctx *ctx, const gts_piece *piece, uint64_t len[6], uint8_t resolved[6]) {
    if (!ctx || !piece) return GTS_ERR_ARG;
    if ((piece->first != 0) != ctx->piece_first) return fail(ctx, GTS_ERR_ARG, "gts_piece_trim: not the piece last scanned");
    DevCtx *dev = dev0(ctx);
    return gts_piece_trim_device(ctx, dev->d_in, ctx->piece_n, piece, len, resolved, dev->stream);
}

void *gts_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) return nullptr;
    return p;
}
void gts_host_free(void *p) {
    if (p) cudaFreeHost(p);
}

int gts_get_stats(const gts_ctx *ctx, gts_stats *st) {
    if (!ctx || !st) return GTS_ERR_ARG;
    *st = ctx->stats;
    return GTS_OK;
}

}  // extern "C"
