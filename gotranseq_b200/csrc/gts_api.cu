// gts_api.cu -- the C ABI of include/gotranseq_b200.h: option decoding, scratch management,
// device passes, the one-shot host call and the streaming host interface.
//
// This file replaces the orchestration of transeq.Translate (transeq/transeq.go:114-173): where
// the reference runs one reader goroutine feeding NumWorker translate goroutines, a gts_ctx cuts
// the byte stream at record boundaries and runs one device pass (six kernels, gts_kernels.h)
// per cut, handing the results back strictly in input order.
#include <cuda_runtime.h>

#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <thread>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/gotranseq_b200.h"
#include "gts_codes.h"
#include "gts_device.h"
#include "gts_kernels.h"

using gts::u32;
using gts::u64;

static thread_local std::string g_create_error;

struct gts_ctx {
    gts_options opt;
    std::string frame;
    uint32_t mask = 0;
    bool reverse = false;
    gts::Lut lut;
    int device = 0;
    int sm_count = 148;
    cudaStream_t stream = nullptr;  // used by the host-buffer entry points
    cudaStream_t s_side = nullptr;  // k_headers runs here, next to k_lines
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;

    // ---- device scratch (grows, never shrinks) ----
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
    gts::Totals *h_totals = nullptr;  // pinned
    u64 *d_dbg = nullptr;             // phase timing sums (GTS_PHASE_TIMING builds with GTS_PHASE_TIMING=1 set)

    // ---- buffers of the host entry points ----
    uint8_t *d_in = nullptr, *d_out = nullptr;
    size_t d_in_cap = 0, d_out_cap = 0;
    uint8_t *h_out = nullptr;
    size_t h_out_cap = 0;
    // ---- streaming host interface: the caller fills one pinned input buffer while a worker thread
    // translates the other; two output buffers, handed back in order ----
    struct StreamJob {
        int in_idx = 0;    // input buffer
        size_t n = 0;      // bytes to translate: [0, n) of it
    };
    uint8_t *s_in[2] = {nullptr, nullptr};   // pinned
    size_t s_in_cap[2] = {0, 0};
    int s_fill_idx = 0;                      // the buffer the caller is filling
    size_t s_fill = 0;                       // bytes in it
    const uint8_t *s_rest_src = nullptr;     // unfinished record of the buffer submitted last: copied
    size_t s_rest_len = 0;                   //   to the front of the next buffer when it is acquired
    uint8_t *s_out[2] = {nullptr, nullptr};  // pinned
    size_t s_out_cap[2] = {0, 0}, s_out_n[2] = {0, 0};
    int s_out_state[2] = {0, 0};             // 0 free, 1 being written, 2 ready for the caller
    int s_out_head = 0, s_out_tail = 0;      // next buffer for the caller / for the worker
    bool s_has_running = false, s_has_pending = false;
    StreamJob s_running, s_pending;
    int s_running_out = 0;
    std::thread s_worker;
    std::mutex s_mu;
    std::condition_variable s_cv;
    bool s_worker_started = false, s_worker_exit = false;
    int s_rc = 0;                            // first error of the worker
    std::string s_err;
    bool stream_first_pass = true, stream_eof = false;

    // ---- chunk pipeline of the host entry points: H2D / kernels / D2H overlap ----
    cudaStream_t s_h2d = nullptr, s_d2h = nullptr;  // `stream` above runs the kernels
    uint8_t *p_in[2] = {nullptr, nullptr}, *p_out[2] = {nullptr, nullptr};
    size_t p_in_cap[2] = {0, 0}, p_out_cap[2] = {0, 0};
    cudaEvent_t ev_h2d[2] = {nullptr, nullptr}, ev_comp[2] = {nullptr, nullptr}, ev_d2h[2] = {nullptr, nullptr};
    gts::Totals *h_totals2 = nullptr;  // pinned, one per pipeline slot

    // ---- last enqueued device pass ----
    bool pending = false;
    gts::Job job;
    cudaStream_t job_stream = nullptr;

    // ---- piece of a split record (gts_piece_*) ----
    bool piece_on = false;
    gts::Job piece_job;      // only the piece fields are used
    size_t piece_n = 0;      // bytes of the piece kept in d_in by gts_piece_scan
    bool piece_first = false;
    uint8_t *d_head = nullptr;  // 192 codes written by the scan pass of a piece

    // ---- per-kernel timing ----
    bool profiling = false;
    std::vector<cudaEvent_t> prof_events;  // 5 per recorded pass
    size_t prof_used = 0;

    u64 max_line = 100ull << 20;  // bufio.Scanner token limit of the reference (transeq.go:27)
    gts::NlPart *d_scan_nl = nullptr;
    size_t scan_nl_cap = 0;
    bool last_truncated = false;  // the last host translation stopped at an over-long line
    bool stream_truncated = false;
    bool poison = false;  // GTS_POISON_OUTPUT=1: fill the output with 0xEE first (tests: no byte left unwritten)

    gts_stats stats;
    std::string err;
};

namespace {

int fail(gts_ctx *ctx, int code, const std::string &msg) {
    if (ctx) ctx->err = msg; else g_create_error = msg;
    return code;
}
int cuda_fail(gts_ctx *ctx, cudaError_t e, const char *what) {
    return fail(ctx, e == cudaErrorMemoryAllocation ? GTS_ERR_NOMEM : GTS_ERR_CUDA,
                std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(call)                                                   \
    do {                                                           \
        cudaError_t e_ = (call);                                   \
        if (e_ != cudaSuccess) return cuda_fail(ctx, e_, #call);   \
    } while (0)

size_t round_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

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
    z.counters = off; off += 16;
    z.bytes = round_up(off, 16);
    return z;
}

int free_rec(gts_ctx *ctx) {
    cudaFree(ctx->d_events); cudaFree(ctx->d_hdr_off); cudaFree(ctx->d_dbase); cudaFree(ctx->d_loc); cudaFree(ctx->d_hinfo);
    cudaFree(ctx->d_rec); cudaFree(ctx->d_trim); cudaFree(ctx->d_ftab);
    ctx->d_ftab = nullptr;
    ctx->d_events = nullptr;
    ctx->d_hdr_off = ctx->d_dbase = ctx->d_loc = ctx->d_hinfo = nullptr;
    ctx->d_rec = nullptr;
    ctx->d_trim = nullptr;
    ctx->rec_cap = 0;
    return 0;
}

int alloc_rec(gts_ctx *ctx, u64 rec_cap) {
    free_rec(ctx);
    CU(cudaMalloc(&ctx->d_events, rec_cap * sizeof(gts::HeaderEvent)));
    CU(cudaMalloc(&ctx->d_hdr_off, rec_cap * sizeof(u64)));
    CU(cudaMalloc(&ctx->d_dbase, rec_cap * sizeof(u64)));
    CU(cudaMalloc(&ctx->d_loc, rec_cap * sizeof(u64)));
    CU(cudaMalloc(&ctx->d_hinfo, rec_cap * sizeof(u64)));
    CU(cudaMalloc(&ctx->d_rec, rec_cap * sizeof(gts::RecInfo)));
    if (ctx->opt.trim) CU(cudaMalloc(&ctx->d_trim, rec_cap * 6 * sizeof(u32)));
    CU(cudaMalloc(&ctx->d_ftab, rec_cap * sizeof(gts::FrameTab)));
    ctx->rec_cap = rec_cap;
    return GTS_OK;
}

u64 parse_tiles(u64 n) { return std::max<u64>(1, (n + gts::kParseTile - 1) / gts::kParseTile); }

// Size the scratch for a pass over n raw bytes.
int ensure_scratch(gts_ctx *ctx, u64 n) {
    if (n > ctx->cap_n || ctx->d_sym == nullptr) {
        const u64 cap = std::max<u64>(n + n / 8, 1u << 20);
        const u64 tiles = parse_tiles(cap);
        cudaFree(ctx->d_sym); cudaFree(ctx->d_tile_info); cudaFree(ctx->d_tile_pre);
        ctx->d_sym = nullptr; ctx->d_tile_info = nullptr; ctx->d_tile_pre = nullptr;
        ctx->cap_n = 0;
        const size_t sym_bytes = (tiles + 2) * gts::kSlotWords * sizeof(u32);
        CU(cudaMalloc(&ctx->d_sym, sym_bytes));
        // zero once: stale words past the end of a slot must still hold valid symbols (<= 4)
        CU(cudaMemset(ctx->d_sym, 0, sym_bytes));
        // the memset runs on the legacy stream, the passes on non-blocking streams: without this
        // it can land after the first k_parse has written its slots
        CU(cudaDeviceSynchronize());
        CU(cudaMalloc(&ctx->d_tile_info, tiles * sizeof(gts::TileInfo)));
        CU(cudaMalloc(&ctx->d_tile_pre, tiles * sizeof(gts::TilePrefix)));
        ctx->cap_n = cap;
    }
    const u64 want_rec = n / 128 + 4096;
    if (want_rec > ctx->rec_cap) {
        int rc = alloc_rec(ctx, want_rec + want_rec / 8);
        if (rc) return rc;
    }
    const size_t nscan = (parse_tiles(ctx->cap_n) + gts::kScanTile - 1) / gts::kScanTile;
    if (nscan > ctx->scan_nl_cap) {
        cudaFree(ctx->d_scan_nl);
        ctx->d_scan_nl = nullptr; ctx->scan_nl_cap = 0;
        CU(cudaMalloc(&ctx->d_scan_nl, nscan * sizeof(gts::NlPart)));
        ctx->scan_nl_cap = nscan;
    }
    const ZeroLayout z = zero_layout(parse_tiles(ctx->cap_n), ctx->rec_cap);
    if (z.bytes > ctx->zero_cap) {
        cudaFree(ctx->d_zero);
        ctx->d_zero = nullptr; ctx->zero_cap = 0;
        CU(cudaMalloc(&ctx->d_zero, z.bytes));
        ctx->zero_cap = z.bytes;
    }
    return GTS_OK;
}

// Enqueue one device pass on `stream` (no host synchronisation).
int enqueue_pass(gts_ctx *ctx, const uint8_t *d_in, u64 n, uint8_t *d_out, u64 cap, cudaStream_t stream,
                 bool sizes_only, gts::Totals *h_dst = nullptr, bool more_follows = false) {
    if (((uintptr_t)d_in & 15) != 0) return fail(ctx, GTS_ERR_ARG, "device input must be 16-byte aligned");
    int rc = ensure_scratch(ctx, n);
    if (rc) return rc;
    const u64 ntiles = parse_tiles(n);
    const ZeroLayout z = zero_layout(ntiles, ctx->rec_cap);
    gts::Job &job = ctx->job;
    job.in = d_in; job.n = n; job.out = d_out; job.out_cap = cap;
    job.sym = ctx->d_sym;
    job.tile_info = ctx->d_tile_info;
    job.tile_pre = ctx->d_tile_pre;
    job.events = ctx->d_events;
    job.hdr_off = ctx->d_hdr_off; job.dbase = ctx->d_dbase; job.loc = ctx->d_loc; job.hinfo = ctx->d_hinfo;
    job.rec = ctx->d_rec; job.trim_len = ctx->d_trim; job.ftab = ctx->d_ftab; job.rec_cap = ctx->rec_cap;
    job.t_status = reinterpret_cast<u64 *>(ctx->d_zero + z.t_status);
    job.s_status = reinterpret_cast<u64 *>(ctx->d_zero + z.s_status);
    job.totals = reinterpret_cast<gts::Totals *>(ctx->d_zero + z.totals);
    job.counters = reinterpret_cast<u32 *>(ctx->d_zero + z.counters);
    job.dbg = ctx->d_dbg;
    job.scan_nl = ctx->d_scan_nl;
    job.max_line = ctx->max_line;
    job.ntiles = (u32)ntiles;
    job.frame_mask = ctx->mask;
    job.alternative = ctx->opt.alternative ? 1 : 0;
    job.trim = ctx->opt.trim ? 1 : 0;
    job.more_follows = more_follows ? 1 : 0;
    job.piece = 0; job.no_end_pads = 0; job.has_carry = 0; job.carry = 0; job.no_headers = 0; job.ov_H = 0;
    job.t0_base = 0; job.win_lo = 0; job.ov_slot = 0; job.ov_n = 0; job.ov_dbase = 0;
    job.head_out = nullptr;
    std::memset(job.head, 0, sizeof(job.head));
    if (ctx->piece_on) {
        const gts::Job &p = ctx->piece_job;
        job.piece = p.piece; job.no_end_pads = p.no_end_pads; job.has_carry = p.has_carry; job.carry = p.carry;
        job.no_headers = p.no_headers; job.ov_H = p.ov_H; job.t0_base = p.t0_base; job.win_lo = p.win_lo;
        job.ov_slot = p.ov_slot; job.ov_n = p.ov_n; job.ov_dbase = p.ov_dbase;
        job.head_out = p.head_out;
        std::memcpy(job.head, p.head, sizeof(job.head));
    }
    job.lut = ctx->lut;
    {
        uint8_t *bf = reinterpret_cast<uint8_t *>(job.lutb.f), *br = reinterpret_cast<uint8_t *>(job.lutb.r);
        for (int e = 0; e < 128; e++) {
            const int x = e < 125 ? e : 0;
            const int s0 = x % 5, s1 = (x / 5) % 5, s2 = x / 25;
            bf[e] = (uint8_t)(ctx->lut.v[x] & 0xFF);
            br[e] = (uint8_t)(ctx->lut.v[s2 + 5 * s1 + 25 * s0] >> 8);
        }
    }
    ctx->job_stream = stream;
    cudaEvent_t *ev = nullptr;
    if (ctx->profiling && !sizes_only) {
        while (ctx->prof_events.size() < ctx->prof_used + 5) {
            cudaEvent_t e;
            CU(cudaEventCreate(&e));
            ctx->prof_events.push_back(e);
        }
        ev = ctx->prof_events.data() + ctx->prof_used;
        ctx->prof_used += 5;
        CU(cudaEventRecord(ev[0], stream));
    }
    CU(cudaMemsetAsync(ctx->d_zero, 0, z.bytes, stream));
    CU(gts::launch_parse(job, stream));
    if (ev) CU(cudaEventRecord(ev[1], stream));
    CU(gts::launch_sizes(job, ctx->sm_count, stream));
    if (ev) CU(cudaEventRecord(ev[2], stream));
    ctx->stats.kernel_launches += 4;
    if (!sizes_only) {
        CU(gts::launch_emit(job, ctx->sm_count, stream, ev ? ev[3] : nullptr, ctx->s_side, ctx->ev_fork, ctx->ev_join));
        if (ev) CU(cudaEventRecord(ev[4], stream));
        ctx->stats.kernel_launches += 2;
    }
    CU(cudaMemcpyAsync(h_dst ? h_dst : ctx->h_totals, job.totals, sizeof(gts::Totals), cudaMemcpyDeviceToHost, stream));
    ctx->pending = true;
    ctx->stats.passes++;
    return GTS_OK;
}

// Wait for the enqueued pass; grow the record table and redo the pass if it was too small.
int finish_pass(gts_ctx *ctx, bool sizes_only) {
    if (!ctx->pending) return fail(ctx, GTS_ERR_ARG, "no device pass enqueued");
    for (int attempt = 0;; attempt++) {
        CU(cudaStreamSynchronize(ctx->job_stream));
        const gts::Totals &t = *ctx->h_totals;
        if (t.flags & gts::kFlagLongLine) {
            // the reference stops reading at a line of >= 100 MiB (transeq.go:186): redo the pass on
            // the bytes in front of that line
            if (attempt >= 3) return fail(ctx, GTS_ERR_ARG, "over-long line handling did not converge");
            const gts::Job j = ctx->job;
            ctx->last_truncated = true;
            int rc = enqueue_pass(ctx, j.in, t.cut, j.out, j.out_cap, ctx->job_stream, sizes_only);
            if (rc) return rc;
            continue;
        }
        if (t.flags & gts::kFlagRecOverflow) {
            if (attempt >= 3) return fail(ctx, GTS_ERR_NOMEM, "record table overflow persists");
            const u64 need = t.n_headers + 2;
            int rc = alloc_rec(ctx, need + need / 16 + 1024);
            if (rc) return rc;
            const gts::Job j = ctx->job;
            rc = enqueue_pass(ctx, j.in, j.n, j.out, j.out_cap, ctx->job_stream, sizes_only);
            if (rc) return rc;
            continue;
        }
        break;
    }
    ctx->pending = false;
    const gts::Totals &t = *ctx->h_totals;
    ctx->stats.records = t.records;
    ctx->stats.nucleotides = t.nucleotides;
    ctx->stats.in_bytes = ctx->job.n;
    ctx->stats.out_bytes = t.out_total;
    return GTS_OK;
}

int ensure_dev_buffer(gts_ctx *ctx, uint8_t **p, size_t *cap, size_t need) {
    if (need <= *cap && *p) return GTS_OK;
    cudaFree(*p);
    *p = nullptr; *cap = 0;
    const size_t want = round_up(need + need / 8 + 4096, 4096);
    CU(cudaMalloc(p, want));
    *cap = want;
    return GTS_OK;
}
int ensure_pinned(gts_ctx *ctx, uint8_t **p, size_t *cap, size_t need, size_t keep) {
    if (need <= *cap && *p) return GTS_OK;
    const size_t want = round_up(need + need / 8 + 4096, 4096);
    uint8_t *q = nullptr;
    CU(cudaHostAlloc(&q, want, cudaHostAllocDefault));
    if (keep && *p) std::memcpy(q, *p, keep);
    if (*p) cudaFreeHost(*p);
    *p = q;
    *cap = want;
    return GTS_OK;
}

// Host bytes in pinned or pageable memory -> device pass -> pinned ctx->h_out.
int translate_host_simple(gts_ctx *ctx, const uint8_t *in, size_t n, size_t *out_n) {
    ctx->last_truncated = false;
    int rc = ensure_dev_buffer(ctx, &ctx->d_in, &ctx->d_in_cap, n + 64);
    if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(ctx->d_in, in, n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n;
    // pass 1: parse + sizes, so that the output buffers can be sized exactly
    rc = enqueue_pass(ctx, ctx->d_in, n, nullptr, 0, ctx->stream, true);
    if (rc) return rc;
    rc = finish_pass(ctx, true);
    if (rc) return rc;
    const size_t total = ctx->h_totals->out_total;
    rc = ensure_dev_buffer(ctx, &ctx->d_out, &ctx->d_out_cap, total + 64);
    if (rc) return rc;
    rc = ensure_pinned(ctx, &ctx->h_out, &ctx->h_out_cap, total + 64, 0);
    if (rc) return rc;
    // pass 2: residues and headers (the parse results are still in the scratch)
    if (ctx->poison && total) CU(cudaMemsetAsync(ctx->d_out, 0xEE, total, ctx->stream));
    ctx->job.out = ctx->d_out;
    ctx->job.out_cap = ctx->d_out_cap;
    // clear the "output too small" flag left by the sizing pass
    ctx->h_totals->flags &= ~gts::kFlagOutOverflow;
    CU(cudaMemcpyAsync(ctx->job.totals, ctx->h_totals, sizeof(gts::Totals), cudaMemcpyHostToDevice, ctx->stream));
    CU(gts::launch_emit(ctx->job, ctx->sm_count, ctx->stream, nullptr, ctx->s_side, ctx->ev_fork, ctx->ev_join));
    ctx->stats.kernel_launches += 2;
    if (total) CU(cudaMemcpyAsync(ctx->h_out, ctx->d_out, total, cudaMemcpyDeviceToHost, ctx->stream));
    ctx->stats.d2h_bytes += total;
    CU(cudaStreamSynchronize(ctx->stream));
    *out_n = total;
    return GTS_OK;
}


// Record boundary (a '>' that starts a line) closest below `pos`, searched back to `lo`; 0 = none.
size_t boundary_below(const uint8_t *p, size_t lo, size_t pos) {
    for (size_t i = pos; i > lo + 1; i--)
        if (p[i - 1] == '>' && p[i - 2] == '\n') return i - 1;
    return 0;
}

// Host bytes -> pinned ctx->h_out through a two-slot pipeline of record-aligned chunks: the copy
// in of chunk i+1 and the copy out of chunk i-1 run while chunk i is in the kernels.  Falls back
// to the simple two-pass path when the input is small, has a record larger than a chunk, or a
// chunk overflows one of the (generously) estimated buffers.
int translate_host(gts_ctx *ctx, const uint8_t *in, size_t n, size_t *out_n) {
    const size_t chunk = ctx->opt.chunk_bytes ? (size_t)ctx->opt.chunk_bytes : (size_t)16 << 20;
    ctx->last_truncated = false;
    // (a chunk shorter than the line limit cannot hold an over-long line; otherwise use the simple path)
    if (n < 2 * chunk || chunk + chunk / 2 >= ctx->max_line) return translate_host_simple(ctx, in, n, out_n);
    // ---- cut points ----
    std::vector<size_t> cuts;
    cuts.push_back(0);
    while (n - cuts.back() > chunk + chunk / 2) {
        const size_t b = boundary_below(in, cuts.back(), cuts.back() + chunk);
        if (b == 0) return translate_host_simple(ctx, in, n, out_n);  // a record larger than a chunk
        cuts.push_back(b);
    }
    cuts.push_back(n);
    const size_t nchunks = cuts.size() - 1;
    // ---- resources ----
    if (!ctx->s_h2d) {
        CU(cudaStreamCreateWithFlags(&ctx->s_h2d, cudaStreamNonBlocking));
        CU(cudaStreamCreateWithFlags(&ctx->s_d2h, cudaStreamNonBlocking));
        for (int b = 0; b < 2; b++) {
            CU(cudaEventCreateWithFlags(&ctx->ev_h2d[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_comp[b], cudaEventDisableTiming));
            CU(cudaEventCreateWithFlags(&ctx->ev_d2h[b], cudaEventDisableTiming));
        }
        CU(cudaHostAlloc(&ctx->h_totals2, 2 * sizeof(gts::Totals), cudaHostAllocDefault));
    }
    const size_t in_need = chunk + chunk / 2 + 64, out_need = 4 * in_need;
    for (int b = 0; b < 2; b++) {
        int rc = ensure_dev_buffer(ctx, &ctx->p_in[b], &ctx->p_in_cap[b], in_need);
        if (rc) return rc;
        rc = ensure_dev_buffer(ctx, &ctx->p_out[b], &ctx->p_out_cap[b], out_need);
        if (rc) return rc;
    }
    int rc = ensure_pinned(ctx, &ctx->h_out, &ctx->h_out_cap, n * 2 + n / 4 + (1 << 20), 0);
    if (rc) return rc;
    rc = ensure_scratch(ctx, in_need);  // no reallocation while chunks are in flight
    if (rc) return rc;

    size_t out_off = 0;
    bool fallback = false;
    // retire chunk j: wait for its kernels, read its exact size, copy it out behind the others
    auto retire = [&](size_t j) -> int {
        const int b = (int)(j & 1);
        CU(cudaEventSynchronize(ctx->ev_comp[b]));
        const gts::Totals &t = ctx->h_totals2[b];
        if (t.flags != 0) {  // record table or output estimate too small: redo everything the simple way
            fallback = true;
            return GTS_OK;
        }
        if (out_off + t.out_total > ctx->h_out_cap) {
            CU(cudaStreamSynchronize(ctx->s_d2h));
            int r2 = ensure_pinned(ctx, &ctx->h_out, &ctx->h_out_cap, (out_off + t.out_total) * 3 / 2, out_off);
            if (r2) return r2;
        }
        CU(cudaStreamWaitEvent(ctx->s_d2h, ctx->ev_comp[b], 0));
        if (t.out_total)
            CU(cudaMemcpyAsync(ctx->h_out + out_off, ctx->p_out[b], t.out_total, cudaMemcpyDeviceToHost, ctx->s_d2h));
        CU(cudaEventRecord(ctx->ev_d2h[b], ctx->s_d2h));
        out_off += t.out_total;
        ctx->stats.d2h_bytes += t.out_total;
        ctx->stats.records += t.records;
        ctx->stats.nucleotides += t.nucleotides;
        return GTS_OK;
    };
    ctx->stats.records = 0;
    ctx->stats.nucleotides = 0;
    for (size_t i = 0; i < nchunks && !fallback; i++) {
        const int b = (int)(i & 1);
        const size_t lo = cuts[i], len = cuts[i + 1] - cuts[i];
        if (i >= 2) CU(cudaStreamWaitEvent(ctx->s_h2d, ctx->ev_comp[b], 0));  // chunk i-2 has left p_in[b]
        CU(cudaMemcpyAsync(ctx->p_in[b], in + lo, len, cudaMemcpyHostToDevice, ctx->s_h2d));
        CU(cudaEventRecord(ctx->ev_h2d[b], ctx->s_h2d));
        ctx->stats.h2d_bytes += len;
        CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_h2d[b], 0));
        if (i >= 2) CU(cudaStreamWaitEvent(ctx->stream, ctx->ev_d2h[b], 0));  // chunk i-2 has left p_out[b]
        if (ctx->poison) CU(cudaMemsetAsync(ctx->p_out[b], 0xEE, ctx->p_out_cap[b], ctx->stream));
        rc = enqueue_pass(ctx, ctx->p_in[b], len, ctx->p_out[b], ctx->p_out_cap[b], ctx->stream, false,
                          &ctx->h_totals2[b], i + 1 < nchunks);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_comp[b], ctx->stream));
        if (i >= 1) {
            rc = retire(i - 1);
            if (rc) return rc;
        }
    }
    if (!fallback) {
        rc = retire(nchunks - 1);
        if (rc) return rc;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    CU(cudaStreamSynchronize(ctx->s_d2h));
    ctx->pending = false;
    if (fallback) return translate_host_simple(ctx, in, n, out_n);
    ctx->stats.in_bytes = n;
    ctx->stats.out_bytes = out_off;
    *out_n = out_off;
    return GTS_OK;
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
long long floordiv3(long long x) { return x >= 0 ? x / 3 : -((-x + 2) / 3); }
long long ceildiv3(long long x) { return -floordiv3(-x); }

int piece_segments(uint32_t mask, bool alternative, const gts_piece &p, u64 n_piece, gts_segment *segs, u64 *total) {
    const u64 n = p.nt_total, H = p.header_len;
    u64 L[6];
    host_frame_lengths(n, alternative, L);
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

    gts_ctx *ctx = new gts_ctx();
    ctx->opt = *opt;
    ctx->frame = frame;
    ctx->opt.frame = ctx->frame.c_str();
    ctx->mask = mask;
    ctx->reverse = reverse;
    std::memset(&ctx->lut, 0, sizeof(ctx->lut));
    gts::build_fused_lut(aa, opt->clean != 0, ctx->lut.v);
    std::memset(&ctx->stats, 0, sizeof(ctx->stats));
    std::memset(&ctx->job, 0, sizeof(ctx->job));

    // No CPU fallback: without a usable CUDA device creation fails.
    cudaError_t e;
    int dev = opt->device;
    if (dev < 0) {
        e = cudaGetDevice(&dev);
    } else {
        e = cudaSetDevice(dev);
    }
    if (e == cudaSuccess) e = cudaFree(nullptr);  // force context creation
    cudaDeviceProp prop;
    if (e == cudaSuccess) e = cudaGetDeviceProperties(&prop, dev);
    if (e == cudaSuccess && prop.major < 10) {
        delete ctx;
        return fail(nullptr, GTS_ERR_CUDA, "gotranseq_b200 needs an sm_100a device (found sm_" +
                                               std::to_string(prop.major) + std::to_string(prop.minor) + ")");
    }
    if (e == cudaSuccess) e = gts::init_kernels();
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->s_side, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaHostAlloc(&ctx->h_totals, sizeof(gts::Totals), cudaHostAllocDefault);
    if (e != cudaSuccess) {
        std::string msg = std::string("CUDA initialisation failed (no CPU fallback): ") + cudaGetErrorString(e);
        if (ctx->stream) cudaStreamDestroy(ctx->stream);
        delete ctx;
        return fail(nullptr, GTS_ERR_CUDA, msg);
    }
    ctx->device = dev;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->poison = std::getenv("GTS_POISON_OUTPUT") != nullptr;
#ifdef GTS_PHASE_TIMING
    if (std::getenv("GTS_PHASE_TIMING")) {
        cudaMalloc(&ctx->d_dbg, 64 * sizeof(u64));
        cudaMemset(ctx->d_dbg, 0, 64 * sizeof(u64));
    }
#endif
    *out = ctx;
    return GTS_OK;
}

void gts_destroy(gts_ctx *ctx) {
    if (!ctx) return;
    if (ctx->s_worker_started) {
        {
            std::lock_guard<std::mutex> lk(ctx->s_mu);
            ctx->s_worker_exit = true;
        }
        ctx->s_cv.notify_all();
        ctx->s_worker.join();
    }
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->d_dbg) {
        u64 h[64];
        cudaDeviceSynchronize();
        cudaMemcpy(h, ctx->d_dbg, sizeof(h), cudaMemcpyDeviceToHost);
        std::fprintf(stderr, "[gts phase timing, ns summed over tiles, %llu passes]", (unsigned long long)ctx->stats.passes);
        for (int i = 0; i < 24; i++) std::fprintf(stderr, " p%d=%llu", i, (unsigned long long)h[i]);
        std::fprintf(stderr, "\n");
        cudaFree(ctx->d_dbg);
    }
    cudaFree(ctx->d_sym); cudaFree(ctx->d_tile_info); cudaFree(ctx->d_tile_pre);
    cudaFree(ctx->d_zero); cudaFree(ctx->d_scan_nl);
    free_rec(ctx);
    cudaFree(ctx->d_in); cudaFree(ctx->d_out); cudaFree(ctx->d_head);
    for (int b = 0; b < 2; b++) {
        cudaFree(ctx->p_in[b]); cudaFree(ctx->p_out[b]);
        if (ctx->ev_h2d[b]) cudaEventDestroy(ctx->ev_h2d[b]);
        if (ctx->ev_comp[b]) cudaEventDestroy(ctx->ev_comp[b]);
        if (ctx->ev_d2h[b]) cudaEventDestroy(ctx->ev_d2h[b]);
    }
    if (ctx->h_totals2) cudaFreeHost(ctx->h_totals2);
    if (ctx->s_h2d) cudaStreamDestroy(ctx->s_h2d);
    if (ctx->s_d2h) cudaStreamDestroy(ctx->s_d2h);
    for (int b = 0; b < 2; b++) {
        if (ctx->s_in[b]) cudaFreeHost(ctx->s_in[b]);
        if (ctx->s_out[b] && ctx->s_out[b] != ctx->h_out) cudaFreeHost(ctx->s_out[b]);
    }
    if (ctx->h_out) cudaFreeHost(ctx->h_out);
    if (ctx->h_totals) cudaFreeHost(ctx->h_totals);
    for (cudaEvent_t e : ctx->prof_events) cudaEventDestroy(e);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->s_side) cudaStreamDestroy(ctx->s_side);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
}

const char *gts_last_error(const gts_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int gts_translate_device_async(gts_ctx *ctx, const void *d_in, size_t n, void *d_out, size_t cap, void *stream) {
    if (!ctx) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    return enqueue_pass(ctx, static_cast<const uint8_t *>(d_in), n, static_cast<uint8_t *>(d_out), cap,
                        static_cast<cudaStream_t>(stream), false);
}

int gts_device_result(gts_ctx *ctx, size_t *out_n) {
    if (!ctx) return GTS_ERR_ARG;
    int rc = finish_pass(ctx, false);
    if (rc) return rc;
    if (out_n) *out_n = ctx->h_totals->out_total;
    if (ctx->h_totals->flags & gts::kFlagOutOverflow)
        return fail(ctx, GTS_ERR_CAPACITY, "output buffer too small: need " +
                                               std::to_string(ctx->h_totals->out_total) + " bytes");
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
    return n * 4 + 4096;
}

int gts_translate_buffer(gts_ctx *ctx, const uint8_t *in, size_t n, const uint8_t **out, size_t *out_n) {
    if (!ctx) return GTS_ERR_ARG;
    if ((n && !in) || !out || !out_n) return fail(ctx, GTS_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    size_t total = 0;
    int rc = translate_host(ctx, in, n, &total);
    if (rc) return rc;
    *out = ctx->h_out;
    *out_n = total;
    return GTS_OK;
}

// ---- streaming host interface ----------------------------------------------------------

// The worker: translates the running job into its output buffer, then promotes the pending job.
static void stream_try_start_locked(gts_ctx *ctx) {
    if (ctx->s_has_running || !ctx->s_has_pending) return;
    if (ctx->s_out_state[ctx->s_out_tail] != 0) return;  // both output buffers are still with the caller
    ctx->s_running = ctx->s_pending;
    ctx->s_has_pending = false;
    ctx->s_has_running = true;
    ctx->s_running_out = ctx->s_out_tail;
    ctx->s_out_state[ctx->s_out_tail] = 1;
    ctx->s_out_tail ^= 1;
}

static void stream_worker(gts_ctx *ctx) {
    cudaSetDevice(ctx->device);
    std::unique_lock<std::mutex> lk(ctx->s_mu);
    while (true) {
        ctx->s_cv.wait(lk, [&] { return ctx->s_worker_exit || ctx->s_has_running; });
        if (ctx->s_worker_exit) return;
        const gts_ctx::StreamJob job = ctx->s_running;
        const int o = ctx->s_running_out;
        const bool skip = ctx->stream_truncated || ctx->s_rc != 0;
        lk.unlock();
        size_t total = 0;
        int rc = GTS_OK;
        if (!skip) {
            // translate_host writes into ctx->h_out: lend it this job's output buffer
            uint8_t *keep = ctx->h_out;
            const size_t keep_cap = ctx->h_out_cap;
            ctx->h_out = ctx->s_out[o];
            ctx->h_out_cap = ctx->s_out_cap[o];
            rc = translate_host(ctx, ctx->s_in[job.in_idx], job.n, &total);
            ctx->s_out[o] = ctx->h_out;
            ctx->s_out_cap[o] = ctx->h_out_cap;
            ctx->h_out = keep;
            ctx->h_out_cap = keep_cap;
        }
        lk.lock();
        if (rc != GTS_OK && ctx->s_rc == 0) {
            ctx->s_rc = rc;
            ctx->s_err = ctx->err;
        }
        if (!skip && rc == GTS_OK && ctx->last_truncated) ctx->stream_truncated = true;  // the reference stopped reading here
        ctx->s_out_n[o] = rc == GTS_OK ? total : 0;
        if (ctx->s_out_n[o]) {
            ctx->s_out_state[o] = 2;
        } else {
            // nothing to hand back: the buffer is free again (it was the last one taken)
            ctx->s_out_state[o] = 0;
            ctx->s_out_tail = o;
        }
        ctx->s_has_running = false;
        stream_try_start_locked(ctx);
        ctx->s_cv.notify_all();
    }
}

static int stream_error_locked(gts_ctx *ctx) {
    if (ctx->s_rc == 0) return GTS_OK;
    ctx->err = ctx->s_err;
    return ctx->s_rc;
}

int gts_acquire_input(gts_ctx *ctx, uint8_t **buf, size_t *cap) {
    if (!ctx || !buf || !cap) return GTS_ERR_ARG;
    if (ctx->stream_eof) return fail(ctx, GTS_ERR_ARG, "gts_acquire_input after eof");
    CU(cudaSetDevice(ctx->device));
    const int k = ctx->s_fill_idx;
    {
        // the buffer may still be the input of a job (the one submitted two submits ago)
        std::unique_lock<std::mutex> lk(ctx->s_mu);
        ctx->s_cv.wait(lk, [&] {
            return !(ctx->s_has_running && ctx->s_running.in_idx == k) && !(ctx->s_has_pending && ctx->s_pending.in_idx == k);
        });
        int rc = stream_error_locked(ctx);
        if (rc) return rc;
    }
    const size_t chunk = ctx->opt.chunk_bytes ? (size_t)ctx->opt.chunk_bytes : (size_t)16 << 20;
    // keep at least half a chunk of free space; a record larger than the buffer makes it grow
    size_t need = ctx->s_fill + ctx->s_rest_len + std::max<size_t>(chunk / 2, 4096);
    if (need < chunk) need = chunk;
    int rc = ensure_pinned(ctx, &ctx->s_in[k], &ctx->s_in_cap[k], need, ctx->s_fill);
    if (rc) return rc;
    if (ctx->s_rest_len) {  // the unfinished record of the buffer submitted last (that buffer is only read by its job)
        std::memcpy(ctx->s_in[k], ctx->s_rest_src, ctx->s_rest_len);
        ctx->s_fill = ctx->s_rest_len;
        ctx->s_rest_len = 0;
        ctx->s_rest_src = nullptr;
    }
    *buf = ctx->s_in[k] + ctx->s_fill;
    *cap = ctx->s_in_cap[k] - ctx->s_fill;
    return GTS_OK;
}

int gts_submit(gts_ctx *ctx, size_t nbytes, int eof) {
    if (!ctx) return GTS_ERR_ARG;
    if (ctx->stream_eof) return fail(ctx, GTS_ERR_ARG, "gts_submit after eof");
    const int k = ctx->s_fill_idx;
    if (ctx->s_fill + nbytes > ctx->s_in_cap[k]) return fail(ctx, GTS_ERR_ARG, "gts_submit: more bytes than acquired");
    ctx->s_fill += nbytes;
    if (eof) ctx->stream_eof = true;
    const size_t chunk = ctx->opt.chunk_bytes ? (size_t)ctx->opt.chunk_bytes : (size_t)16 << 20;
    size_t cut = 0;
    if (eof) {
        cut = ctx->s_fill;
        if (cut == 0 && !ctx->stream_first_pass) return GTS_OK;  // nothing left after the last record boundary
    } else {
        if (ctx->s_fill < chunk) return GTS_OK;  // keep filling
        // last record boundary: a '>' that starts a line (transeq.go:198)
        const uint8_t *p = ctx->s_in[k];
        for (size_t i = ctx->s_fill; i-- > 1;) {
            if (p[i] == '>' && p[i - 1] == '\n') {
                cut = i;
                break;
            }
        }
        if (cut == 0) return GTS_OK;  // one record larger than the buffer: keep reading (buffer grows)
    }
    ctx->stream_first_pass = false;
    std::unique_lock<std::mutex> lk(ctx->s_mu);
    int rc = stream_error_locked(ctx);
    if (rc) return rc;
    if (!ctx->s_worker_started) {
        ctx->s_worker = std::thread(stream_worker, ctx);
        ctx->s_worker_started = true;
    }
    ctx->s_cv.wait(lk, [&] { return !ctx->s_has_pending; });  // at most one job waits behind the running one
    ctx->s_pending.in_idx = k;
    ctx->s_pending.n = cut;
    ctx->s_has_pending = true;
    stream_try_start_locked(ctx);
    ctx->s_cv.notify_all();
    // the caller goes on with the other buffer; the unfinished record moves there when it is acquired
    ctx->s_rest_src = ctx->s_in[k] + cut;
    ctx->s_rest_len = ctx->s_fill - cut;
    ctx->s_fill_idx = k ^ 1;
    ctx->s_fill = 0;
    return GTS_OK;
}

int gts_next_output(gts_ctx *ctx, const uint8_t **buf, size_t *n) {
    if (!ctx || !buf || !n) return GTS_ERR_ARG;
    *buf = nullptr;
    *n = 0;
    std::unique_lock<std::mutex> lk(ctx->s_mu);
    const int h = ctx->s_out_head;
    if (ctx->stream_eof)  // drain: wait for the next result, or for the worker to run dry
        ctx->s_cv.wait(lk, [&] { return ctx->s_out_state[h] == 2 || (!ctx->s_has_running && !ctx->s_has_pending); });
    int rc = stream_error_locked(ctx);
    if (rc) return rc;
    if (ctx->s_out_state[h] == 2) {
        *buf = ctx->s_out[h];
        *n = ctx->s_out_n[h];
        return GTS_OK;
    }
    if (ctx->stream_eof) {  // drained: ready for another stream
        ctx->stream_truncated = false;
        ctx->stream_eof = false;
        ctx->stream_first_pass = true;
        ctx->s_fill = 0;
        ctx->s_rest_len = 0;
        ctx->s_rest_src = nullptr;
    }
    return GTS_OK;
}

int gts_release_output(gts_ctx *ctx) {
    if (!ctx) return GTS_ERR_ARG;
    std::unique_lock<std::mutex> lk(ctx->s_mu);
    const int h = ctx->s_out_head;
    if (ctx->s_out_state[h] != 2) return GTS_OK;
    ctx->s_out_state[h] = 0;
    ctx->s_out_n[h] = 0;
    ctx->s_out_head ^= 1;
    stream_try_start_locked(ctx);
    ctx->s_cv.notify_all();
    return GTS_OK;
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
    const size_t np = ctx->prof_used / 5;
    for (size_t p = 0; p < np; p++) {
        cudaEvent_t *ev = ctx->prof_events.data() + 5 * p;
        CU(cudaEventSynchronize(ev[4]));
        for (int k = 0; k < 4; k++) {
            float t = 0.f;
            CU(cudaEventElapsedTime(&t, ev[k], ev[k + 1]));
            ms[k] += t;
        }
    }
    if (passes) *passes = np;
    ctx->prof_used = 0;
    return GTS_OK;
}

// ---- one record split over several GPUs ---------------------------------------------------

int gts_piece_scan_device(gts_ctx *ctx, const void *d_in, size_t n, int first, gts_piece_info *info, void *stream) {
    if (!ctx || !info) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    if (ctx->opt.trim) return fail(ctx, GTS_ERR_ARG, "a split record cannot be trimmed");
    CU(cudaSetDevice(ctx->device));
    ctx->piece_on = true;
    ctx->piece_job = gts::Job();
    ctx->piece_job.no_end_pads = 1;  // count only: no pass of its own end
    if (!ctx->d_head) CU(cudaMalloc(&ctx->d_head, 256));
    ctx->piece_job.head_out = ctx->d_head;
    ctx->last_truncated = false;
    int rc = enqueue_pass(ctx, static_cast<const uint8_t *>(d_in), n, nullptr, 0, static_cast<cudaStream_t>(stream), true);
    if (!rc) rc = finish_pass(ctx, true);
    ctx->piece_on = false;
    if (rc) return rc;
    const gts::Totals &t = *ctx->h_totals;
    if (ctx->last_truncated) return fail(ctx, GTS_ERR_ARG, "a split record cannot hold an over-long line");
    info->headers = t.n_headers;
    info->nucleotides = t.total_sym - 2 - 2 * t.n_headers;
    info->tail[0] = (uint8_t)(t.tail & 0xF);
    info->tail[1] = (uint8_t)((t.tail >> 4) & 0xF);
    info->reserved[0] = info->reserved[1] = 0;
    info->header_len = 0;
    CU(cudaMemcpy(info->head, ctx->d_head, sizeof(info->head), cudaMemcpyDeviceToHost));
    if (first) {
        if (t.n_headers != 1) return fail(ctx, GTS_ERR_ARG, "the first piece must hold exactly one header line");
        u64 hoff = 0, hinfo = 0;
        CU(cudaMemcpy(&hoff, ctx->d_hdr_off + 1, sizeof(u64), cudaMemcpyDeviceToHost));
        CU(cudaMemcpy(&hinfo, ctx->d_hinfo + 1, sizeof(u64), cudaMemcpyDeviceToHost));
        if (hoff != 0) return fail(ctx, GTS_ERR_ARG, "the first piece must start with the header line");
        info->header_len = (uint32_t)(hinfo & 0xFFFFFFFFull);
    } else if (t.n_headers != 0) {
        return fail(ctx, GTS_ERR_ARG, "only the first piece of a split record may hold a header line");
    }
    return GTS_OK;
}

int gts_piece_translate_device(gts_ctx *ctx, const void *d_in, size_t n, const gts_piece *piece, void *d_out,
                               size_t cap, size_t *out_total, gts_segment segs[GTS_MAX_SEGMENTS], int *nseg,
                               void *stream) {
    if (!ctx || !piece || !segs || !nseg) return GTS_ERR_ARG;
    if (n && !d_in) return fail(ctx, GTS_ERR_ARG, "d_in is NULL");
    if (ctx->opt.trim) return fail(ctx, GTS_ERR_ARG, "a split record cannot be trimmed");
    CU(cudaSetDevice(ctx->device));
    gts::Job &p = ctx->piece_job;
    p = gts::Job();
    p.piece = 1;
    p.no_end_pads = piece->last ? 0 : 1;
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
    ctx->piece_on = true;
    int rc = enqueue_pass(ctx, static_cast<const uint8_t *>(d_in), n, static_cast<uint8_t *>(d_out), cap,
                          static_cast<cudaStream_t>(stream), false);
    if (!rc) rc = finish_pass(ctx, false);
    ctx->piece_on = false;
    if (rc) return rc;
    const gts::Totals &t = *ctx->h_totals;
    if (out_total) *out_total = t.out_total;
    if (t.flags & gts::kFlagOutOverflow)
        return fail(ctx, GTS_ERR_CAPACITY, "output buffer too small: need " + std::to_string(t.out_total) + " bytes");
    const u64 n_piece = t.total_sym - 2 - (piece->first ? 2 : 0);
    u64 total = 0;
    *nseg = piece_segments(ctx->mask, ctx->opt.alternative != 0, *piece, n_piece, segs, &total);
    if (total != t.out_total) return fail(ctx, GTS_ERR_ARG, "split record: host and device output sizes differ");
    return GTS_OK;
}

int gts_piece_scan(gts_ctx *ctx, const uint8_t *in, size_t n, int first, gts_piece_info *info) {
    if (!ctx || !info) return GTS_ERR_ARG;
    if (n && !in) return fail(ctx, GTS_ERR_ARG, "NULL argument");
    CU(cudaSetDevice(ctx->device));
    int rc = ensure_dev_buffer(ctx, &ctx->d_in, &ctx->d_in_cap, n + 64);
    if (rc) return rc;
    if (n) CU(cudaMemcpyAsync(ctx->d_in, in, n, cudaMemcpyHostToDevice, ctx->stream));
    ctx->stats.h2d_bytes += n;
    ctx->piece_n = n;
    ctx->piece_first = first != 0;
    return gts_piece_scan_device(ctx, ctx->d_in, n, first, info, ctx->stream);
}

int gts_piece_translate(gts_ctx *ctx, const gts_piece *piece, size_t *out_total, gts_segment segs[GTS_MAX_SEGMENTS],
                        const uint8_t *data[GTS_MAX_SEGMENTS], int *nseg) {
    if (!ctx || !piece || !segs || !data || !nseg) return GTS_ERR_ARG;
    if ((piece->first != 0) != ctx->piece_first) return fail(ctx, GTS_ERR_ARG, "gts_piece_translate: not the piece last scanned");
    CU(cudaSetDevice(ctx->device));
    // the whole record's output size is known from the counts alone
    u64 L[6], total = 0;
    host_frame_lengths(piece->nt_total, ctx->opt.alternative != 0, L);
    for (int f = 0; f < 6; f++)
        if ((ctx->mask >> f) & 1u) total += host_run_size(piece->header_len, L[f]);
    int rc = ensure_dev_buffer(ctx, &ctx->d_out, &ctx->d_out_cap, total + 64);
    if (rc) return rc;
    if (ctx->poison && total) CU(cudaMemsetAsync(ctx->d_out, 0xEE, total, ctx->stream));
    size_t tot = 0;
    rc = gts_piece_translate_device(ctx, ctx->d_in, ctx->piece_n, piece, ctx->d_out, ctx->d_out_cap, &tot, segs, nseg,
                                    ctx->stream);
    if (rc) return rc;
    size_t bytes = 0;
    for (int i = 0; i < *nseg; i++) bytes += round_up(segs[i].length, 16);
    rc = ensure_pinned(ctx, &ctx->h_out, &ctx->h_out_cap, bytes + 64, 0);
    if (rc) return rc;
    size_t off = 0;
    for (int i = 0; i < *nseg; i++) {
        CU(cudaMemcpyAsync(ctx->h_out + off, ctx->d_out + segs[i].offset, segs[i].length, cudaMemcpyDeviceToHost, ctx->stream));
        data[i] = ctx->h_out + off;
        off += round_up(segs[i].length, 16);
        ctx->stats.d2h_bytes += segs[i].length;
    }
    CU(cudaStreamSynchronize(ctx->stream));
    if (out_total) *out_total = tot;
    return GTS_OK;
}

void *gts_host_alloc(size_t bytes) {
    void *p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess) return nullptr;
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
