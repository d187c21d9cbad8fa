// engine_tsan.cpp -- TEST INFRASTRUCTURE: the chunk engine of gts_api.cu under ThreadSanitizer, on the stand-in
// device (tests/c/standin/).  Built with -fsanitize=thread and run by tests/test_engine_standin.py; exercises the
// calls that cross threads: reader thread + writer thread (gts_wait_output), several device contexts, any_order,
// gts_cancel from a third thread, a reported write error, gts_stream_reset and reuse, the one-shot call, the
// warning callback, destroy with chunks in flight.  Prints "tsan driver ok"; ThreadSanitizer's reports go to stderr.
//   engine_tsan IN.fna WANT.faa      (WANT = the oracle's output for -f 6 -T)
#include <algorithm>
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "gotranseq_b200.h"

static std::string slurp(const char *path) {
    std::string s;
    FILE *f = std::fopen(path, "rb");
    if (!f) { std::perror(path); std::exit(2); }
    char buf[1 << 16];
    size_t k;
    while ((k = std::fread(buf, 1, sizeof(buf), f)) > 0) s.append(buf, k);
    std::fclose(f);
    return s;
}
#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed line %d: %s (%s)\n", __LINE__, #c, ctx ? gts_last_error(ctx) : ""); std::exit(1); } } while (0)

static std::atomic<uint64_t> g_warns{0};
static void on_warn(void *, const uint8_t *, size_t, uint8_t, uint64_t) { g_warns++; }

static gts_ctx *make(int ngpu, size_t chunk, int any_order) {
    gts_options o;
    std::memset(&o, 0, sizeof(o));
    o.frame = "6"; o.trim = 1; o.device = 0; o.chunk_bytes = chunk; o.num_gpus = ngpu; o.any_order = any_order;
    for (int i = 0; i < 8; i++) o.device_ids[i] = i & 1;
    gts_ctx *ctx = nullptr;
    if (gts_create(&o, &ctx) != GTS_OK) { std::fprintf(stderr, "create: %s\n", gts_last_error(nullptr)); std::exit(1); }
    return ctx;
}

// reader = this thread, writer = a second thread; returns the stream's return code and the output
static int stream(gts_ctx *ctx, const std::string &in, size_t read_size, std::string *out, long fail_after, long cancel_after) {
    std::atomic<int> wrc{0};
    std::thread writer([&] {
        long writes = 0;
        while (true) {
            const uint8_t *p = nullptr;
            size_t m = 0;
            const int rc = gts_wait_output(ctx, &p, &m);
            if (rc != GTS_OK) { wrc = rc; return; }
            if (m == 0) return;
            if (fail_after >= 0 && writes == fail_after) { gts_report_write_error(ctx, "disk full"); continue; }
            out->append(reinterpret_cast<const char *>(p), m);
            writes++;
            gts_release_output(ctx);
        }
    });
    int rc = GTS_OK;
    size_t pos = 0;
    long reads = 0;
    std::thread canceller;
    while (rc == GTS_OK) {
        uint8_t *buf = nullptr;
        size_t cap = 0;
        rc = gts_acquire_input(ctx, &buf, &cap);
        if (rc != GTS_OK) break;
        size_t k = in.size() - pos < cap ? in.size() - pos : cap;
        if (read_size && k > read_size) k = read_size;
        std::memcpy(buf, in.data() + pos, k);
        pos += k;
        if (cancel_after >= 0 && reads++ == cancel_after) canceller = std::thread([ctx] { gts_cancel(ctx); });
        rc = gts_submit(ctx, k, k == 0);
        if (k == 0) break;
    }
    if (rc != GTS_OK && wrc == 0) gts_cancel(ctx);  // let the writer's wait return
    writer.join();
    if (canceller.joinable()) canceller.join();
    return rc != GTS_OK ? rc : wrc.load();
}

static std::vector<std::string> entries(const std::string &s) {  // per-frame entries, sorted
    std::vector<std::string> v;
    size_t a = 0;
    while (a < s.size()) {
        size_t b = s.find("\n>", a);
        if (b == std::string::npos) b = s.size() - 1;
        v.push_back(s.substr(a, b + 1 - a));
        a = b + 1;
    }
    std::sort(v.begin(), v.end());
    return v;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    const std::string in = slurp(argv[1]), want = slurp(argv[2]);
    gts_ctx *ctx = nullptr;
    for (int ngpu : {1, 3}) {
        for (int any : {0, 1}) {
            ctx = make(ngpu, 1 << 15, any);
            CHECK(gts_set_warning_callback(ctx, on_warn, nullptr) == GTS_OK);
            std::string out;
            CHECK(stream(ctx, in, 20000, &out, -1, -1) == GTS_OK);
            CHECK(any ? entries(out) == entries(want) : out == want);
            // a failing write: the stream's error is GTS_ERR_WRITE, the context streams again after a reset
            out.clear();
            CHECK(stream(ctx, in, 20000, &out, 3, -1) == GTS_ERR_WRITE);
            CHECK(std::strstr(gts_last_error(ctx), "fail to write to output file: disk full") != nullptr);
            CHECK(gts_stream_reset(ctx) == GTS_OK);
            // cancel from a third thread
            out.clear();
            CHECK(stream(ctx, in, 20000, &out, -1, 5) == GTS_ERR_CANCELLED);
            CHECK(gts_stream_reset(ctx) == GTS_OK);
            out.clear();
            CHECK(stream(ctx, in, 0, &out, -1, -1) == GTS_OK);
            CHECK(any ? entries(out) == entries(want) : out == want);
            // the one-shot call on the same context
            const uint8_t *p = nullptr;
            size_t m = 0;
            CHECK(gts_translate_buffer(ctx, reinterpret_cast<const uint8_t *>(in.data()), in.size(), &p, &m) == GTS_OK);
            CHECK(m == want.size() && std::memcmp(p, want.data(), m) == 0);
            gts_destroy(ctx);
        }
    }
    // destroy with chunks queued and running
    ctx = make(2, 1 << 14, 0);
    size_t pos = 0;
    for (int i = 0; i < 6 && pos < in.size(); i++) {
        uint8_t *buf = nullptr;
        size_t cap = 0;
        CHECK(gts_acquire_input(ctx, &buf, &cap) == GTS_OK);
        const size_t k = in.size() - pos < cap ? in.size() - pos : cap;
        std::memcpy(buf, in.data() + pos, k);
        pos += k;
        CHECK(gts_submit(ctx, k, 0) == GTS_OK);
    }
    gts_destroy(ctx);
    std::printf("tsan driver ok (%llu warnings)\n", (unsigned long long)g_warns.load());
    return 0;
}
