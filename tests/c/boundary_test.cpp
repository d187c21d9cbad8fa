// Record-boundary search of the chunk engine (gts_boundary.h) against a plain loop, on buffers long
// enough for the multi-threaded paths.  Prints "ok" and exits 0.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../gotranseq_b200/csrc/gts_boundary.h"

static size_t ref_below(const uint8_t *p, size_t lo, size_t pos) {  // largest x in (lo, pos) with p[x-1..x] == "\n>"
    for (size_t x = pos; x-- > lo + 1;)
        if (p[x] == '>' && p[x - 1] == '\n') return x;
    return 0;
}
static size_t ref_above(const uint8_t *p, size_t pos, size_t n) {  // smallest x in [pos, n) with p[x-1..x] == "\n>"
    for (size_t x = pos; x < n; x++)
        if (p[x] == '>' && p[x - 1] == '\n') return x;
    return n;
}

int main() {
    const size_t n = (size_t)96 << 20;
    std::vector<uint8_t> buf(n, 'A');
    unsigned long long s = 12345;
    auto rnd = [&] { s = s * 6364136223846793005ull + 1442695040888963407ull; return (size_t)(s >> 20); };
    for (size_t i = 60; i < n; i += 61) buf[i] = '\n';
    int checks = 0;
    for (int round = 0; round < 40; round++) {
        // a few boundaries, lone '>' inside lines, at random places (also near segment borders)
        std::vector<size_t> touched;
        const int k = round < 5 ? 0 : (int)(rnd() % 4);
        for (int j = 0; j < k; j++) {
            size_t x = 1 + rnd() % (n - 2);
            if (round % 3 == 0) x = ((rnd() % 12) * (n / 12)) + (rnd() % 5);  // around multiples of n/12 (segment borders for several thread counts)
            if (x < 1) x = 1;
            if (x >= n) x = n - 1;
            touched.push_back(x - 1); touched.push_back(x);
            buf[x - 1] = '\n'; buf[x] = '>';
        }
        for (int j = 0; j < 3; j++) { const size_t x = 2 + rnd() % (n - 3); if (buf[x - 1] != '\n') { touched.push_back(x); buf[x] = '>'; } }
        for (int q = 0; q < 6; q++) {
            size_t lo = rnd() % (n / 2), pos = lo + 2 + rnd() % (n - lo - 2);
            if (q == 0) { lo = 0; pos = n; }
            if (q == 1) { lo = rnd() % 100; pos = n - rnd() % 100; }
            if (gts_host::boundary_below(buf.data(), lo, pos) != ref_below(buf.data(), lo, pos)) { std::printf("below mismatch round %d lo %zu pos %zu\n", round, lo, pos); return 1; }
            const size_t from = lo + 1;
            if (gts_host::boundary_above(buf.data(), from, pos) != ref_above(buf.data(), from, pos)) { std::printf("above mismatch round %d from %zu n %zu\n", round, from, pos); return 1; }
            checks += 2;
        }
        for (size_t x : touched) buf[x] = (x % 61 == 60) ? '\n' : 'A';
    }
    std::printf("ok %d\n", checks);
    return 0;
}
