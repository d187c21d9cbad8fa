// standin_pass.cpp -- TEST INFRASTRUCTURE ONLY: the device pass of gotranseq_b200 (k_parse .. k_headers) restated on
// the host behind the launchers of gts_kernels.h, for tests/test_engine_standin.py (see standin_cuda.cpp).  It is
// NOT the oracle and not a product path: an independent, straightforward walk over the lines of the buffer
// (transeq.go:183-216, encodedSequence.go:17-43, writer.go:57-169) that fills the same Totals / output / warning
// records the kernels fill, so that the host engine in gts_api.cu sees what it would see from the GPU -- including
// the rare flags (record table / output / warning list too small, a line of >= max_line bytes).
//   GTS_STANDIN_SLOW_BYTES / GTS_STANDIN_SLOW_MS: passes over at least that many bytes take that long (to put a
//   large chunk behind small ones in time).
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "gts_kernels.h"

namespace gts {

namespace {

struct Result {
    std::string out;
    std::vector<WarnRec> warns;
};
std::mutex g_mu;
std::map<const Totals *, Result> g_results;  // by the pass's totals block (one per device context)

int code_of(uint8_t b) {
    switch (b) {
    case 'A': case 'a': return 1;
    case 'C': case 'c': return 2;
    case 'T': case 't': case 'U': case 'u': return 3;
    case 'G': case 'g': return 4;
    default: return 0;
    }
}
bool is_n(uint8_t b) { return b == 'N' || b == 'n'; }
int reverse_start_host(u64 n, int i, bool alternative) {  // writer.go:64-79
    if (alternative) return i;
    int r = (int)(n % 3) - i;
    return r < 0 ? r + 3 : r;
}

struct Record {
    u64 hdr_off = 0;
    u32 H = 0;
    std::vector<uint8_t> codes;
};

void emit_record(const Job &job, const uint8_t *in, const Record &r, std::string &out) {
    const u64 n = r.codes.size();
    std::vector<uint8_t> rc;
    if (job.frame_mask & 0x38u) {  // encodedSequence.go:71-97
        rc.resize(n);
        for (u64 i = 0; i < n; i++) {
            const uint8_t c = r.codes[n - 1 - i];
            rc[i] = c == 0 ? 0 : (uint8_t)(((c - 1) ^ 2) + 1);
        }
    }
    const uint8_t *hdr = in + r.hdr_off;
    const void *spp = r.H ? std::memchr(hdr, ' ', r.H) : nullptr;
    const u32 sp = spp ? (u32)(static_cast<const uint8_t *>(spp) - hdr) : r.H;
    for (int f = 0; f < 6; f++) {
        if (!((job.frame_mask >> f) & 1u)) continue;
        const uint8_t *s = f < 3 ? r.codes.data() : rc.data();
        const long start = f < 3 ? f : reverse_start_host(n, f - 3, job.alternative != 0);
        // writer.go:120-131
        out.append(reinterpret_cast<const char *>(hdr), sp);
        out.push_back('_');
        out.push_back((char)('1' + f));
        out.append(reinterpret_cast<const char *>(hdr) + sp, r.H - sp);
        out.push_back('\n');
        // writer.go:97-112
        std::string aa;
        for (u64 pos = (u64)start; pos + 2 < n; pos += 3) aa.push_back((char)(job.lut.v[s[pos] + 5 * s[pos + 1] + 25 * s[pos + 2]] & 0xFF));
        const long rem = ((long)n - start) % 3;  // truncates toward zero, like Go's
        if (rem == 2) aa.push_back((char)(job.lut.v[s[n - 2] + 5 * s[n - 1]] & 0xFF));
        else if (rem == 1) aa.push_back('X');
        if (job.trim)
            while (!aa.empty() && (aa.back() == 'X' || aa.back() == '*')) aa.pop_back();  // writer.go:141-163
        for (size_t i = 0; i < aa.size(); i += 60) {
            out.append(aa, i, 60);
            out.push_back('\n');
        }
    }
}

void run_pass(const Job &job) {
    Totals t;
    std::memset(&t, 0, sizeof(t));
    Result res;
    const uint8_t *in = job.in;
    const u64 n = job.n;
    std::vector<Record> recs(1);  // slot 0: the bytes in front of the first header line
    u64 sym = 0;
    bool too_long = false;
    for (u64 pos = 0; pos < n;) {
        const void *nlp = std::memchr(in + pos, '\n', n - pos);
        const u64 end = nlp ? (u64)(static_cast<const uint8_t *>(nlp) - in) : n;
        if (end - pos >= job.max_line) {  // transeq.go:27,186
            t.flags |= kFlagLongLine;
            t.cut = pos;
            too_long = true;
            break;
        }
        u64 len = end - pos;
        const u64 a = pos;
        pos = nlp ? end + 1 : n;
        if (len > 0 && in[a + len - 1] == '\r') len--;
        if (len == 0) continue;
        if (in[a] == '>') {
            Record r;
            r.hdr_off = a;
            r.H = (u32)len;
            recs.push_back(std::move(r));
            continue;
        }
        Record &r = recs.back();
        for (u64 i = 0; i < len; i++) {
            const uint8_t b = in[a + i];
            const int c = code_of(b);
            if (c == 0 && !is_n(b) && job.warn) {
                WarnRec w;
                w.key = sym;
                w.hdr_off = r.hdr_off;
                w.pos = r.codes.size();
                w.H = r.H;
                w.byte = b;
                res.warns.push_back(w);
            }
            r.codes.push_back((uint8_t)c);
            sym++;
        }
    }
    const u64 R = recs.size() - 1;
    t.n_headers = R;
    t.total_sym = sym + 2 * (R + 1);
    if (!too_long) {
        if (R + 2 > job.rec_cap) {
            t.flags |= kFlagRecOverflow;
        } else {
            for (u64 s = 0; s <= R; s++) {
                const Record &r = recs[s];
                const bool emitted = s >= 1 || !r.codes.empty() || (R == 0 && !job.more_follows);  // transeq.go:198-213
                if (!emitted) continue;
                t.records++;
                t.nucleotides += r.codes.size();
                emit_record(job, in, r, res.out);
            }
            t.out_total = res.out.size();
            if (job.out && t.out_total > job.out_cap) t.flags |= kFlagOutOverflow;
            if (job.warn) {
                t.n_warn = res.warns.size();
                if (t.n_warn > job.warn_cap) t.flags |= kFlagWarnOverflow;
            }
        }
    }
    *job.totals = t;
    std::lock_guard<std::mutex> lk(g_mu);
    g_results[job.totals] = std::move(res);
}

void maybe_slow(const Job &job) {
    static const u64 slow_bytes = std::getenv("GTS_STANDIN_SLOW_BYTES") ? std::strtoull(std::getenv("GTS_STANDIN_SLOW_BYTES"), nullptr, 10) : 0;
    static const int slow_ms = std::getenv("GTS_STANDIN_SLOW_MS") ? std::atoi(std::getenv("GTS_STANDIN_SLOW_MS")) : 0;
    if (slow_bytes && job.n >= slow_bytes && slow_ms > 0) std::this_thread::sleep_for(std::chrono::milliseconds(slow_ms));
}

}  // namespace

cudaError_t init_kernels() { return cudaSuccess; }
cudaError_t launch_parse(const Job &, cudaStream_t) { return cudaSuccess; }
cudaError_t launch_sizes(const Job &job, int, cudaStream_t) {
    run_pass(job);
    return cudaSuccess;
}
cudaError_t launch_warn(const Job &job, int, cudaStream_t) {
    std::lock_guard<std::mutex> lk(g_mu);
    const Result &r = g_results[job.totals];
    if (job.totals->flags == 0 && job.wres && !r.warns.empty() && r.warns.size() <= job.warn_cap)
        std::memcpy(job.wres, r.warns.data(), r.warns.size() * sizeof(WarnRec));
    return cudaSuccess;
}
cudaError_t launch_emit(const Job &job, int, cudaStream_t, cudaEvent_t, const EmitStreams &) {
    maybe_slow(job);
    std::lock_guard<std::mutex> lk(g_mu);
    const Result &r = g_results[job.totals];
    if (job.totals->flags == 0 && job.out && !r.out.empty() && r.out.size() <= job.out_cap) std::memcpy(job.out, r.out.data(), r.out.size());
    return cudaSuccess;
}
cudaError_t launch_publish(const Totals *d_totals, Totals *h_totals, cudaStream_t) {
    *h_totals = *d_totals;
    return cudaSuccess;
}
// one record split over several GPUs: not part of the stand-in
cudaError_t launch_tile_scan(const Job &, int, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_piece_plan(const PieceInfoDev *, int, int, PieceParams *, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_piece_fix(const Job &, cudaStream_t) { return cudaErrorNotSupported; }
cudaError_t launch_piece_trim(const Job &, cudaStream_t) { return cudaErrorNotSupported; }

}  // namespace gts
