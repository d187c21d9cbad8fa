// transeq.cpp -- transeq::Translate over the streaming C ABI (replaces the reader goroutine /
// worker pool / flush loop of transeq/transeq.go:126-163, writer.go:171-181).
#include "transeq.hpp"

#include "../../include/gotranseq_b200.h"

namespace transeq {

std::string Translate(std::istream &in, std::ostream &out, const Options &o) {
    gts_options opt{};
    opt.frame = o.Frame.c_str();
    opt.table = o.Table;
    opt.clean = o.Clean;
    opt.alternative = o.Alternative;
    opt.trim = o.Trim;
    opt.num_worker = o.NumWorker;
    opt.device = o.Device;
    opt.chunk_bytes = o.ChunkBytes;
    gts_ctx *ctx = nullptr;
    if (gts_create(&opt, &ctx) != GTS_OK) return gts_last_error(nullptr);
    std::string err;
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
            gts_release_output(ctx);
            if (!out) { err = "fail to write to output file: stream error"; break; }  // writer.go:175
        }
    }
    gts_destroy(ctx);
    return err;
}

}  // namespace transeq
