// stream_harness.cpp -- bench.py's streaming legs without the Python interpreter in the loop.
//
// BENCH INFRASTRUCTURE (not part of the product): the call sequence of the Go shim
// (go/transeq/translate_b200.go; the reference's reader goroutine transeq.go:183-216 and its writers
// writer.go:171-181) over a FASTA that already sits in host memory:
//   reader = the calling thread: gts_acquire_input -> memcpy into the pinned buffer (with a few helper
//            threads, what an io.Reader over a file or a buffer is free to do) -> gts_submit
//   writer = a second thread: gts_wait_output -> (count the bytes) -> gts_release_output
// The library's entry points come in as function pointers, so that the harness drives whatever build
// of the library bench.py has loaded (GOTRANSEQ_B200_LIB A/B runs) and links against nothing.
// Built by bench.py on first use: g++ -O2 -std=c++17 -pthread -shared -fPIC.
#include <atomic>
#include <condition_variable>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <thread>
#include <vector>

extern "C" {

struct gts_ctx;
struct stream_api {
    int (*acquire_input)(gts_ctx *, uint8_t **, size_t *);
    int (*submit)(gts_ctx *, size_t, int);
    int (*wait_output)(gts_ctx *, const uint8_t **, size_t *);
    int (*release_output)(gts_ctx *);
    int (*cancel)(gts_ctx *);
};
struct stream_result {
    uint64_t out_bytes, chunks;
    int rc;  // first non-zero return code of a library call (0 = the stream ran to its end)
};

}  // extern "C"

namespace {

// memcpy by `threads` threads: helpers that live as long as one harness call
class CopyPool {
public:
    explicit CopyPool(int helpers) {
        for (int i = 0; i < helpers; i++) workers_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void copy(uint8_t *dst, const uint8_t *src, size_t n) {
        const size_t parts = workers_.size() + 1;
        if (workers_.empty() || n < ((size_t)4 << 20)) {
            std::memcpy(dst, src, n);
            return;
        }
        const size_t per = (n + parts - 1) / parts;
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = dst; src_ = src; n_ = n; per_ = per;
            pending_ = (int)workers_.size();
            gen_++;
        }
        cv_.notify_all();
        std::memcpy(dst, src, per < n ? per : n);  // the caller takes the first part
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [this] { return pending_ == 0; });
    }

private:
    void run(int i) {
        uint64_t seen = 0;
        std::unique_lock<std::mutex> lk(mu_);
        while (true) {
            cv_.wait(lk, [&] { return stop_ || gen_ != seen; });
            if (stop_) return;
            seen = gen_;
            const size_t lo = per_ * (size_t)(i + 1);
            const size_t hi = lo + per_ < n_ ? lo + per_ : n_;
            uint8_t *d = dst_;
            const uint8_t *s = src_;
            lk.unlock();
            if (lo < hi) std::memcpy(d + lo, s + lo, hi - lo);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    bool stop_ = false;
    uint64_t gen_ = 0;
    int pending_ = 0;
    uint8_t *dst_ = nullptr;
    const uint8_t *src_ = nullptr;
    size_t n_ = 0, per_ = 0;
};

}  // namespace

// The FASTA in[0, n) streamed `repeats` times back to back as ONE stream (in must end with a whole record, so that the
// periods join at a record boundary), read_size > 0: at most that many bytes per gts_submit.
extern "C" int stream_harness_run(const stream_api *api, gts_ctx *ctx, const uint8_t *in, size_t n, int repeats,
                                  int reader_threads, size_t read_size, stream_result *res) {
    res->out_bytes = 0;
    res->chunks = 0;
    res->rc = 0;
    std::atomic<int> wrc{0};
    uint64_t out_bytes = 0, chunks = 0;
    std::thread writer([&] {
        const uint8_t *p = nullptr;
        size_t m = 0;
        while (true) {
            const int rc = api->wait_output(ctx, &p, &m);
            if (rc != 0) {
                wrc.store(rc);
                return;
            }
            if (m == 0) return;  // the stream is drained
            out_bytes += m;
            chunks++;
            api->release_output(ctx);
        }
    });
    int rc = 0;
    {
        CopyPool pool(reader_threads > 1 ? reader_threads - 1 : 0);
        for (int r = 0; r < repeats && rc == 0; r++) {
            size_t pos = 0;
            while (pos < n && rc == 0) {
                uint8_t *buf = nullptr;
                size_t cap = 0;
                rc = api->acquire_input(ctx, &buf, &cap);
                if (rc != 0) break;
                size_t k = cap < n - pos ? cap : n - pos;
                if (read_size && k > read_size) k = read_size;
                pool.copy(buf, in + pos, k);
                pos += k;
                rc = api->submit(ctx, k, 0);
            }
        }
        if (rc == 0) rc = api->submit(ctx, 0, 1);
    }
    if (rc != 0) api->cancel(ctx);  // the writer's wait returns
    writer.join();
    res->out_bytes = out_bytes;
    res->chunks = chunks;
    res->rc = rc != 0 ? rc : wrc.load();
    return res->rc;
}
