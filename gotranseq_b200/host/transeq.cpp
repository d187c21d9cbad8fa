// transeq.cpp -- transeq::Translate over the streaming C ABI (replaces the reader goroutine /
// worker pool / flush loop of transeq/transeq.go:126-163, writer.go:171-181).
#include "transeq.hpp"

#include <unistd.h>

#include <cerrno>
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

std::string TranslateFd(int fd_in, int fd_out, const Options &o, const char *out_path) {
    std::string err;
    gts_ctx *ctx = create(o, &err);
    if (!ctx) return err;
    // the writer: every finished chunk, in input order (flush, writer.go:171-181)
    std::thread writer([&] {
        while (true) {
            const uint8_t *p = nullptr;
            size_t n = 0;
            if (gts_wait_output(ctx, &p, &n) != GTS_OK || n == 0) return;
            size_t done = 0;
            while (done < n) {
                const ssize_t w = ::write(fd_out, p + done, n - done);
                if (w < 0) {
                    if (errno == EINTR) continue;
                    // (*os.File).Write returns a *PathError: "write <path>: <errno text, lower case>"
                    std::string m = std::strerror(errno);
                    if (!m.empty() && m[0] >= 'A' && m[0] <= 'Z') m[0] = (char)(m[0] - 'A' + 'a');
                    gts_report_write_error(ctx, (std::string("write ") + out_path + ": " + m).c_str());
                    return;
                }
                done += (size_t)w;
            }
            gts_release_output(ctx);
        }
    });
    // the reader: this thread (readSequenceFromFasta runs on the caller's goroutine, transeq.go:161)
    bool eof = false;
    while (!eof) {
        uint8_t *buf = nullptr;
        size_t cap = 0;
        if (gts_acquire_input(ctx, &buf, &cap) != GTS_OK) break;
        size_t got = 0;
        while (got < cap) {
            const ssize_t r = ::read(fd_in, buf + got, cap - got);
            if (r < 0 && errno == EINTR) continue;
            if (r <= 0) { eof = true; break; }  // read errors end the stream silently (transeq.go:192-215)
            got += (size_t)r;
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
    gts_destroy(ctx);
    return err;
}

}  // namespace transeq
