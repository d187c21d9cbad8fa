// gts_boundary.h -- record boundaries in a host buffer (plain C++, no CUDA): where the chunk engine
// (gts_api.cu) may cut the byte stream.
//
// Reference behaviour reproduced (feliixx/gotranseq): a record starts at a line whose first byte is
// '>' (transeq/transeq.go:198), i.e. at a '>' that follows a '\n' (or starts the input).
#pragma once
#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace gts_host {

// Record boundary (a '>' that starts a line, transeq.go:198) closest below `pos`, searched back to
// `lo` (exclusive); 0 = none.
inline size_t boundary_below_serial(const uint8_t *p, size_t lo, size_t pos) {
    size_t i = pos;
    while (i > lo + 1) {
        const void *q = memrchr(p + lo + 1, '>', i - (lo + 1));
        if (!q) return 0;
        i = (size_t)(static_cast<const uint8_t *>(q) - p);
        if (p[i - 1] == '\n') return i;
    }
    return 0;
}
// (Ordinary FASTA has a boundary a few KB to a few MB below `pos`: the top 16 MB are searched by the
// calling thread alone, which stops at the first hit.  Only what lies below them -- the new bytes of
// a record much larger than a chunk -- goes to a few threads, one segment each.)
inline size_t boundary_below(const uint8_t *p, size_t lo, size_t pos) {
    const size_t len = pos > lo ? pos - lo : 0;
    const size_t top = (size_t)16 << 20;
    if (len <= 2 * top) return boundary_below_serial(p, lo, pos);
    const size_t near = boundary_below_serial(p, pos - top, pos);
    if (near) return near;
    const size_t hi_all = pos - top + 1;  // a '>' at hi_all - 1 or below has not been looked at yet
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = std::min<size_t>(std::min<size_t>(8, hw ? hw : 1), (hi_all - lo) >> 23);  // >= 8 MB per thread
    if (nthreads < 2) return boundary_below_serial(p, lo, hi_all);
    std::vector<size_t> found(nthreads, 0);
    std::vector<std::thread> th;
    const size_t seg = (hi_all - lo + nthreads - 1) / nthreads;
    for (size_t t = 0; t < nthreads; t++) {
        th.emplace_back([&, t] {
            // '>' positions in (a, b): boundary_below_serial(p, a, b) looks at p[a + 1 .. b - 1]
            const size_t a = lo + t * seg, b = std::min(hi_all, a + seg + 1);
            found[t] = a + 1 < b ? boundary_below_serial(p, a, b) : 0;
        });
    }
    for (auto &x : th) x.join();
    for (size_t t = nthreads; t-- > 0;)
        if (found[t]) return found[t];
    return 0;
}
// First record boundary at or above `pos` (pos >= 1), n = none.
inline size_t boundary_above_serial(const uint8_t *p, size_t pos, size_t n) {
    size_t i = pos;
    while (i < n) {
        const void *q = memchr(p + i, '>', n - i);
        if (!q) return n;
        i = (size_t)(static_cast<const uint8_t *>(q) - p);
        if (p[i - 1] == '\n') return i;
        i++;
    }
    return n;
}
// A record larger than a chunk (a chromosome) makes this the only work of the calling thread while
// the GPUs wait for the chunk: long ranges are searched by a few threads, one segment each.
inline size_t boundary_above(const uint8_t *p, size_t pos, size_t n) {
    const size_t len = n - pos;
    const unsigned hw = std::thread::hardware_concurrency();
    const size_t nthreads = std::min<size_t>(std::min<size_t>(8, hw ? hw : 1), len >> 24);  // >= 16 MB per thread
    if (nthreads < 2) return boundary_above_serial(p, pos, n);
    std::vector<size_t> found(nthreads, n);
    std::vector<std::thread> th;
    const size_t seg = (len + nthreads - 1) / nthreads;
    for (size_t t = 0; t < nthreads; t++) {
        th.emplace_back([&, t] {
            const size_t lo = pos + t * seg, hi = std::min(n, lo + seg);
            // (a boundary is looked at by the segment that holds its '>'; the byte in front of it may lie in the one before)
            const size_t r = lo < hi ? boundary_above_serial(p, lo, hi) : hi;
            found[t] = r < hi ? r : n;
        });
    }
    for (auto &x : th) x.join();
    for (size_t t = 0; t < nthreads; t++)
        if (found[t] < n) return found[t];
    return n;
}


}  // namespace gts_host
