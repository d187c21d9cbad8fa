/* workloads_gen.c -- counter-based synthetic FASTA of the benchmark shapes (SURVEY.md 8d /
 * BASELINE.json configs).  BENCH INFRASTRUCTURE, not product: built by __graft_entry__.build() and
 * by workloads.py on first use (gcc), loaded through ctypes.
 *
 * Every record is a pure function of (kind, seed, record index): splitmix64 over a counter, so any
 * range of records -- hence any shard of the stream -- can be generated independently, on any
 * rank, by any number of threads, and always gives the same bytes.
 *
 *   kind 0  readme:      n ~ U[300,3000], ">seq%07d len=%d", 60-column lines, 0.1 % N
 *   kind 1  contigs:     n ~ U[4.5,5.5] Mbp, ">contig_%05d", 60-column lines, 0.01 % N
 *   kind 2  short reads: 150 nt on one line, 52-byte Illumina-style header line
 *   kind 3  chromosome:  one record, n = 250 000 000 (scaled by `scale_nt`), ">chr1", 60-column lines
 */
#include <stdint.h>
#include <stdio.h>
#include <string.h>

static inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
static inline uint64_t key(uint64_t seed, uint64_t rec, uint64_t ctr) {
    return splitmix64(splitmix64(seed * 0x100000001B3ull + rec) ^ (ctr * 0xD6E8FEB86659FD93ull));
}

static uint64_t rec_nt(int kind, uint64_t seed, uint64_t i, uint64_t scale_nt) {
    switch (kind) {
    case 0: return 300 + key(seed, i, ~0ull) % 2701;
    case 1: return 4500000 + key(seed, i, ~0ull) % 1000001;
    case 2: return 150;
    default: return scale_nt ? scale_nt : 250000000ull;
    }
}
static int rec_header(int kind, uint64_t seed, uint64_t i, uint64_t n, char *dst) {
    switch (kind) {
    case 0: return sprintf(dst, ">seq%07llu len=%llu\n", (unsigned long long)i, (unsigned long long)n);
    case 1: return sprintf(dst, ">contig_%05llu\n", (unsigned long long)i);
    case 2: {
        const uint64_t h = key(seed, i, ~1ull);
        return sprintf(dst, ">A00123:45:HXXXXXXXX:1:1101:%05u:%05u 1:N:0:ATCACG\n", (unsigned)(h % 100000), (unsigned)((h >> 32) % 100000));
    }
    default: return sprintf(dst, ">chr1\n");
    }
}
static uint64_t width_of(int kind) { return kind == 2 ? 150 : 60; }

/* bytes of record i */
uint64_t wg_record_size(int kind, uint64_t seed, uint64_t i, uint64_t scale_nt) {
    char h[128];
    const uint64_t n = rec_nt(kind, seed, i, scale_nt), w = width_of(kind);
    return (uint64_t)rec_header(kind, seed, i, n, h) + n + (n + w - 1) / w;
}
uint64_t wg_record_nt(int kind, uint64_t seed, uint64_t i, uint64_t scale_nt) { return rec_nt(kind, seed, i, scale_nt); }

uint64_t wg_total_nt(int kind, uint64_t seed, uint64_t first, uint64_t count, uint64_t scale_nt) {
    uint64_t t = 0;
    for (uint64_t k = 0; k < count; k++) t += rec_nt(kind, seed, first + k, scale_nt);
    return t;
}

/* sizes[k] = bytes of record first + k */
void wg_record_sizes(int kind, uint64_t seed, uint64_t first, uint64_t count, uint64_t scale_nt, uint64_t *sizes) {
    for (uint64_t k = 0; k < count; k++) sizes[k] = wg_record_size(kind, seed, first + k, scale_nt);
}

/* Nucleotides [lo, hi) of record i, with the line breaks that fall behind them; returns bytes written. */
static uint64_t fill_bases(int kind, uint64_t seed, uint64_t i, uint64_t n, uint64_t lo, uint64_t hi, uint8_t *dst) {
    static const char ACGT[4] = {'A', 'C', 'G', 'T'};
    const uint64_t w = width_of(kind);
    const uint64_t n_rate = kind == 0 ? 1000 : kind == 1 || kind == 3 ? 10000 : 0;
    uint8_t *p = dst;
    uint64_t pos = lo;
    while (pos < hi) {
        /* one counter block = 24 nucleotides (48 bits of the block's key); a second draw decides
         * whether the block holds an N (probability 24 / n_rate) and where */
        const uint64_t blk = pos / 24;
        const uint64_t r = key(seed, i, blk);
        const uint64_t b0 = blk * 24, b1 = b0 + 24 < hi ? b0 + 24 : hi;
        uint64_t bits = r >> (2 * (pos - b0));
        const uint64_t rn = splitmix64(r);
        const int has_n = n_rate && (rn % n_rate) < 24;
        const uint64_t npos = b0 + (rn >> 32) % 24;
        for (; pos < b1; pos++, bits >>= 2) {
            *p++ = (uint8_t)((has_n && pos == npos) ? 'N' : ACGT[bits & 3]);
            if ((pos + 1) % w == 0 || pos + 1 == n) *p++ = '\n';
        }
    }
    return (uint64_t)(p - dst);
}

/* Records [first, first + count) into dst (capacity cap); returns the bytes written, or 0 if cap is too small. */
uint64_t wg_generate(int kind, uint64_t seed, uint64_t first, uint64_t count, uint64_t scale_nt, uint8_t *dst, uint64_t cap) {
    uint64_t off = 0;
    for (uint64_t k = 0; k < count; k++) {
        const uint64_t i = first + k, n = rec_nt(kind, seed, i, scale_nt);
        const uint64_t need = wg_record_size(kind, seed, i, scale_nt);
        if (off + need > cap) return 0;
        off += (uint64_t)rec_header(kind, seed, i, n, (char *)dst + off);
        off += fill_bases(kind, seed, i, n, 0, n, dst + off);
    }
    return off;
}

/* Part of ONE record: nucleotides [lo, hi) with their line breaks, the header line first when lo == 0
 * (for multi-threaded generation of a giant record; lo must be a multiple of 120). */
uint64_t wg_generate_part(int kind, uint64_t seed, uint64_t i, uint64_t scale_nt, uint64_t lo, uint64_t hi, uint8_t *dst) {
    const uint64_t n = rec_nt(kind, seed, i, scale_nt);
    uint64_t off = 0;
    if (hi > n) hi = n;
    if (lo == 0) off += (uint64_t)rec_header(kind, seed, i, n, (char *)dst);
    if (lo < hi) off += fill_bases(kind, seed, i, n, lo, hi, dst + off);
    return off;
}
