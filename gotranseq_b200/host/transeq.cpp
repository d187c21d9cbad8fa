// transeq.cpp -- transeq::Translate over the streaming C ABI (replaces the reader goroutine /
// worker pool / flush loop of transeq/transeq.go:126-163, writer.go:171-181).
#include "transeq.hpp"

#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cerrno>
#include <chrono>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/gotranseq_b200.h"

namespace transeq {

namespace {

// encodedSequence.go:39: fmt.Printf on stdout, one line per invalid nucleotide
void print_warning(void *, const uint8_t *header, size_t header_len, uint8_t byte, uint64_t pos) {
    char small[512];
    const size_t need = gts_format_warning(header, header_len, byte, pos, small, sizeof(small));
    if (need <= sizeof(small)) {
        std::fwrite(small, 1, need, stdout);
    } else {
        std::vector<char> big(need);
        gts_format_warning(header, header_len, byte, pos, big.data(), need);
        std::fwrite(big.data(), 1, need, stdout);
    }
}

gts_ctx *create(const Options &o, std::string *err) {
    gts_options opt{};
    opt.frame = o.Frame.c_str();
    opt.table = o.Table;
    opt.clean = o.Clean;
    opt.alternative = o.Alternative;
    opt.trim = o.Trim;
    opt.num_worker = o.NumWorker;
    opt.device = o.Device;
    opt.chunk_bytes = o.ChunkBytes;
    opt.num_gpus = o.NumGpus < 0 ? -1 : o.NumGpus;
    opt.any_order = o.AnyOrder ? 1 : 0;
    for (int i = 0; i < 8; i++) opt.device_ids[i] = i;
    gts_ctx *ctx = nullptr;
    if (gts_create(&opt, &ctx) != GTS_OK) {
        *err = gts_last_error(nullptr);
        return nullptr;
    }
    if (o.Warnings) gts_set_warning_callback(ctx, print_warning, nullptr);
    return ctx;
}

}  // namespace

std::string Translate(std::istream &in, std::ostream &out, const Options &o) {
    std::string err;
    gts_ctx *ctx = create(o, &err);
    if (!ctx) return err;
    bool eof = false;
    while (!eof && err.empty()) {
        uint8_t *buf = nullptr;
        size_t cap = 0;
        if (gts_acquire_input(ctx, &buf, &cap) != GTS_OK) { err = gts_last_error(ctx); break; }
        in.read(reinterpret_cast<char *>(buf), (std::streamsize)cap);
        const size_t got = (size_t)in.gcount();
        eof = got == 0 || in.eof();  // read errors end the stream silently, like transeq.go:192
        if (gts_submit(ctx, got, eof ? 1 : 0) != GTS_OK) { err = gts_last_error(ctx); break; }
        while (true) {
            const uint8_t *p = nullptr;
            size_t n = 0;
            if (gts_next_output(ctx, &p, &n) != GTS_OK) { err = gts_last_error(ctx); break; }
            if (n == 0) break;
            out.write(reinterpret_cast<const char *>(p), (std::streamsize)n);
            if (!out) {
                // writer.go:171-179: the first failed write cancels the stream and becomes the error
                gts_report_write_error(ctx, "stream error");
                gts_next_output(ctx, &p, &n);
                err = gts_last_error(ctx);
                break;
            }
            gts_release_output(ctx);
        }
    }
    gts_destroy(ctx);
    return err;
}

namespace {

// n bytes at file offset off, split over a few threads (page-cache reads and writes are memcpys: one
// thread moves ~5-10 GB/s, the PCIe link next to it 25-50 GB/s)
template <typename F>
bool parallel_io(size_t n, int threads, F io) {  // io(lo, hi) -> bool
    if (n < ((size_t)4 << 20) || threads <= 1) return io((size_t)0, n);
    std::vector<std::thread> ts;
    std::vector<char> ok((size_t)threads, 1);
    const size_t per = (n + (size_t)threads - 1) / (size_t)threads;
    for (int t = 0; t < threads; t++) {
        const size_t lo = std::min(n, per * (size_t)t), hi = std::min(n, lo + per);
        ts.emplace_back([&, t, lo, hi] { ok[(size_t)t] = io(lo, hi) ? 1 : 0; });
    }
    for (auto &t : ts) t.join();
    for (char c : ok)
        if (!c) return false;
    return true;
}

bool is_regular(int fd) {
    struct stat st;
    return fstat(fd, &st) == 0 && S_ISREG(st.st_mode);
}

}  // namespace

std::string TranslateFd(int fd_in, int fd_out, const Options &o, const char *out_path) {
    std::string err;
    const bool timing = std::getenv("GOTRANSEQ_B200_TIMING") != nullptr;
    const auto t0 = std::chrono::steady_clock::now();
    gts_ctx *ctx = create(o, &err);
    if (!ctx) return err;
    const auto t1 = std::chrono::steady_clock::now();
    const int io_threads = o.IoThreads > 0 ? o.IoThreads : 4;
    const bool in_reg = is_regular(fd_in), out_reg = is_regular(fd_out);
    // the writer: every finished chunk, in input order (flush, writer.go:171-181)
    std::thread writer([&] {
        off_t woff = 0;
        while (true) {
            const uint8_t *p = nullptr;
            size_t n = 0;
            if (gts_wait_output(ctx, &p, &n) != GTS_OK || n == 0) return;
            int werr = 0;
            const bool ok = parallel_io(n, out_reg ? io_threads : 1, [&](size_t lo, size_t hi) {
                while (lo < hi) {
                    const ssize_t w = out_reg ? ::pwrite(fd_out, p + lo, hi - lo, woff + (off_t)lo) : ::write(fd_out, p + lo, hi - lo);
                    if (w < 0) {
                        if (errno == EINTR) continue;
                        werr = errno;
                        return false;
                    }
                    lo += (size_t)w;
                }
                return true;
            });
            if (!ok) {
                // (*os.File).Write returns a *PathError: "write <path>: <errno text, lower case>"
                std::string m = std::strerror(werr);
                if (!m.empty() && m[0] >= 'A' && m[0] <= 'Z') m[0] = (char)(m[0] - 'A' + 'a');
                gts_report_write_error(ctx, (std::string("write ") + out_path + ": " + m).c_str());
                return;
            }
            woff += (off_t)n;
            gts_release_output(ctx);
        }
    });
    // the reader: this thread (readSequenceFromFasta runs on the caller's goroutine, transeq.go:161)
    bool eof = false;
    off_t roff = 0;
    off_t in_size = 0;
    if (in_reg) {
        struct stat st;
        fstat(fd_in, &st);
        in_size = st.st_size;
    }
    while (!eof) {
        uint8_t *buf = nullptr;
        size_t cap = 0;
        if (gts_acquire_input(ctx, &buf, &cap) != GTS_OK) break;
        size_t got = 0;
        if (in_reg) {
            const size_t want = (size_t)std::min<off_t>((off_t)cap, in_size - roff);
            std::vector<size_t> done((size_t)io_threads + 1, 0);
            parallel_io(want, io_threads, [&](size_t lo, size_t hi) {
                while (lo < hi) {
                    const ssize_t r = ::pread(fd_in, buf + lo, hi - lo, roff + (off_t)lo);
                    if (r < 0 && errno == EINTR) continue;
                    if (r <= 0) return false;  // read errors end the stream silently (transeq.go:192-215)
                    lo += (size_t)r;
                }
                return true;
            });
            got = want;
            roff += (off_t)want;
            eof = roff >= in_size;
        } else {
            while (got < cap) {
                const ssize_t r = ::read(fd_in, buf + got, cap - got);
                if (r < 0 && errno == EINTR) continue;
                if (r <= 0) { eof = true; break; }
                got += (size_t)r;
            }
        }
        if (gts_submit(ctx, got, eof ? 1 : 0) != GTS_OK) break;
    }
    if (!eof) gts_cancel(ctx);  // a failed call: make sure the writer wakes up
    writer.join();
    {
        const uint8_t *p = nullptr;
        size_t n = 0;
        if (gts_next_output(ctx, &p, &n) != GTS_OK) err = gts_last_error(ctx);
    }
    const auto t2 = std::chrono::steady_clock::now();
    gts_destroy(ctx);
    if (timing) {
        const auto t3 = std::chrono::steady_clock::now();
        auto ms = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
            return std::chrono::duration<double, std::milli>(b - a).count();
        };
        std::fprintf(stderr, "{\"create_ms\": %.1f, \"stream_ms\": %.1f, \"destroy_ms\": %.1f, \"gpus\": %d}\n", ms(t0, t1), ms(t1, t2),
                     ms(t2, t3), 0);
    }
    return err;
}

}  // namespace transeq
