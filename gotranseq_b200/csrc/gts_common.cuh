// gts_common.cuh -- device helpers shared by the kernels (memory-model loads/stores, byte SIMD,
// chained scan, frame arithmetic).
#pragma once
#include <cuda_runtime.h>

#include "gts_device.h"

namespace gts {

// ---- flagged status words of the chained scans ----
__device__ __forceinline__ u64 ld_acquire(const u64 *p) {
    u64 v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u64 ld_relaxed(const u64 *p) {
    u64 v;
    asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_relaxed(u64 *p, u64 v) {
    asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_release(u64 *p, u64 v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// PTX prmt in its default mode: selector nibble bit 3 replicates the chosen byte's sign bit
// (__byte_perm masks that bit away, and costs an extra AND for run-time selectors)
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel) {
    u32 r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

// Programmatic dependent launch: a kernel launched with the programmatic-stream-serialization
// attribute may start while its predecessor in the stream is still running; it must not touch the
// predecessor's results before pdl_wait() (which returns once that grid has completed and its
// writes are visible).  pdl_launch() lets the successor's CTAs be scheduled early.  Both are
// no-ops for ordinary launches.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

#ifdef GTS_PHASE_TIMING
__device__ __forceinline__ u64 gtime() {
    u64 t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define PHASE_MARK(job, idx, t_prev)                                                     \
    do {                                                                                 \
        if ((job).dbg && threadIdx.x == 0) {                                             \
            const u64 now_ = gtime();                                                    \
            atomicAdd((job).dbg + (idx), now_ - (t_prev));                               \
            (t_prev) = now_;                                                             \
        }                                                                                \
    } while (0)
#else
#define PHASE_MARK(job, idx, t_prev) do { } while (0)
#endif

// 0x80 in every byte of x that is '\n'
__device__ __forceinline__ u32 newline_flags(u32 x) {
    return __vabsdiffu4(__vabsdiffu4(x, 0x0A0A0A0Au), 0x80808080u) & 0x80808080u;
}
// Bit i of the result = byte i of the 16 bytes is '\n'.  The per-byte flags (0x80) of a word are
// gathered with one integer dot product: flags . (1,2,4,8) = 128 * nibble, accumulated over the
// four words with the weights shifted by the nibble position.
__device__ __forceinline__ u32 gather_flags16(u32 f0, u32 f1, u32 f2, u32 f3) {  // flags: 0x80 or 0 per byte
    u32 acc = __dp4a(f0, 0x08040201u, 0u);                // bits 7..10
    acc = __dp4a(f1, 0x80402010u, acc);                   // bits 11..14
    u32 hi = __dp4a(f2, 0x08040201u, 0u);
    hi = __dp4a(f3, 0x80402010u, hi);
    return (acc >> 7) | (hi << 1);
}
__device__ __forceinline__ u32 newline_mask16(uint4 v) {
    return gather_flags16(newline_flags(v.x), newline_flags(v.y), newline_flags(v.z), newline_flags(v.w));
}
// 0x80 in every byte of x that equals c (c replicated into the four bytes of c4)
__device__ __forceinline__ u32 byte_eq_flags(u32 x, u32 c4) {
    return __vabsdiffu4(__vabsdiffu4(x, c4), 0x80808080u) & 0x80808080u;
}
// 0x80 in every byte of x that is 'N' or 'n'
__device__ __forceinline__ u32 n_flags(u32 x) {
    return __vabsdiffu4(__vabsdiffu4(x & 0xDFDFDFDFu, 0x4E4E4E4Eu), 0x80808080u) & 0x80808080u;
}
// 0x80 in every byte of c (nucleotide codes 0..4) that is 0
__device__ __forceinline__ u32 zero_code_flags(u32 c) { return (0x80808080u - c) & 0x80808080u; }

// the piece description of a pass: in device memory when the pass was planned on the device
__device__ __forceinline__ const PieceParams &piece_of(const Job &job) { return job.pp ? *job.pp : job.pv; }

struct Pair {
    u64 a, b;
};

// Decoupled look-back over the tiles of a chained scan (sum of two counters).  Called by all 32
// lanes of one warp of tile t with that tile's aggregate; publishes it, walks back over the
// predecessors' published aggregates / inclusive prefixes, publishes this tile's inclusive
// prefix and returns the exclusive one.  Tiles are numbered by an atomic ticket, so every
// predecessor is running or finished and the wait terminates.
__device__ __forceinline__ Pair scan_lookback(u64 *status, u32 t, Pair agg) {
    const int lane = threadIdx.x & 31;
    u64 *mine = status + 4ull * t;
    if (lane == 0) {
        st_relaxed(mine + 1, agg.b);
        st_release(mine + 0, agg.a | kValid);
    }
    Pair excl = {0, 0};
    i64 pos = (i64)t - 1;
    while (pos >= 0) {
        const i64 idx = pos - lane;
        int kind = 2;  // 0 = not published, 1 = aggregate, 2 = inclusive (or before tile 0)
        u64 a = 0, b = 0;
        if (idx >= 0) {
            const u64 *st = status + 4ull * (u64)idx;
            u64 w = ld_acquire(st + 2);
            if (w & kValid) {
                a = w & ~kValid;
                b = ld_relaxed(st + 3);
            } else {
                w = ld_acquire(st + 0);
                if (w & kValid) {
                    kind = 1;
                    a = w & ~kValid;
                    b = ld_relaxed(st + 1);
                } else {
                    kind = 0;
                }
            }
        }
        const unsigned inc_m = __ballot_sync(0xFFFFFFFFu, kind == 2);
        const unsigned none_m = __ballot_sync(0xFFFFFFFFu, kind == 0);
        const int first_inc = inc_m ? __ffs(inc_m) - 1 : 32;
        const int first_none = none_m ? __ffs(none_m) - 1 : 32;
        if (first_none < first_inc) continue;  // a needed predecessor has not published yet
        const int last = first_inc < 31 ? first_inc : 31;
        if (lane > last) a = b = 0;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a += __shfl_xor_sync(0xFFFFFFFFu, a, o);
            b += __shfl_xor_sync(0xFFFFFFFFu, b, o);
        }
        excl.a += a;
        excl.b += b;
        if (first_inc < 32) break;
        pos -= 32;
    }
    if (lane == 0) {
        st_relaxed(mine + 3, excl.b + agg.b);
        st_release(mine + 2, (excl.a + agg.a) | kValid);
    }
    return excl;
}

// exact unsigned division by constants without the 64-bit division routine
__device__ __forceinline__ u64 udiv3(u64 a) { return __umul64hi(a, 0xAAAAAAAAAAAAAAABull) >> 1; }
__device__ __forceinline__ u64 udiv60(u64 a) { return __umul64hi(a >> 2, 0x8888888888888889ull) >> 3; }
__device__ __forceinline__ i64 floordiv3(i64 a) { return a >= 0 ? (i64)udiv3((u64)a) : -(i64)udiv3((u64)(-a) + 2); }
__device__ __forceinline__ i64 ceildiv3(i64 a) { return -floordiv3(-a); }

// rc start position of reverse frame i (0..2): writer.go:64-79
__device__ __forceinline__ int reverse_start(u64 n, int i, bool alternative) {
    if (alternative) return i;
    int r = (int)(n - 3 * udiv3(n)) - i;
    return r < 0 ? r + 3 : r;
}

// residues of each frame before trimming: writer.go:97-112 (Go's % truncates, so a start
// position beyond the sequence emits nothing)
__device__ __forceinline__ void frame_lengths(u64 n, bool alternative, u64 L[6]) {
#pragma unroll
    for (int k = 0; k < 3; k++) L[k] = n >= (u64)k ? udiv3(n - k + 2) : 0;
#pragma unroll
    for (int i = 0; i < 3; i++) {
        const u64 p = (u64)reverse_start(n, i, alternative);
        L[3 + i] = n >= p ? udiv3(n - p + 2) : 0;
    }
}

__device__ __forceinline__ u64 run_size(u64 H, u64 Lp) { return H + 3 + Lp + udiv60(Lp + 59); }

// transeq.go:198-213: a record is sent when a header line follows it (if it has any bytes) and,
// always, at the end of the stream
__device__ __forceinline__ bool record_emitted(u64 s, u64 n, u64 R, bool more_follows) {
    return s >= 1 || n > 0 || (R == 0 && !more_follows);
}

// symbol i of a tile's slot
__device__ __forceinline__ u32 slot_symbol(const u32 *sym, u32 tile, u32 i) {
    const u32 w = sym[(u64)tile * kSlotWords + (i >> 3)];
    const u32 b = i & 7u;
    return (w >> (8 * (b & 3) + 4 * (b >> 2))) & 0xFu;
}

// ---- shared-memory output image -> global memory ----
__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, u32 bytes) {
    const u32 saddr = (u32)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
}

// Copy the bytes [lo, hi) of the image (hi - lo < 16, both inside one aligned 16-byte slot) to
// global memory with naturally aligned stores; g is the global address of image byte 0.
__device__ __forceinline__ void store_edge(const uint8_t *image, uint8_t *g, u32 lo, u32 hi) {
    u32 x = lo;
    if ((x & 1u) && x + 1u <= hi) { g[x] = image[x]; x += 1u; }
    if ((x & 2u) && x + 2u <= hi) { *reinterpret_cast<uint16_t *>(g + x) = *reinterpret_cast<const uint16_t *>(image + x); x += 2u; }
    if ((x & 4u) && x + 4u <= hi) { *reinterpret_cast<u32 *>(g + x) = *reinterpret_cast<const u32 *>(image + x); x += 4u; }
    if (x + 8u <= hi) { *reinterpret_cast<uint2 *>(g + x) = *reinterpret_cast<const uint2 *>(image + x); x += 8u; }
    if (x + 4u <= hi) { *reinterpret_cast<u32 *>(g + x) = *reinterpret_cast<const u32 *>(image + x); x += 4u; }
    if (x + 2u <= hi) { *reinterpret_cast<uint16_t *>(g + x) = *reinterpret_cast<const uint16_t *>(image + x); x += 2u; }
    if (x + 1u <= hi) { g[x] = image[x]; }
}
// The 1..15 bytes in front of the 16-byte aligned image offset `end`: one store per set bit of the length.
__device__ __forceinline__ void store_head(const uint8_t *image, uint8_t *g, u32 lo, u32 end) {
    const u32 h = end - lo;
    u32 x = end;
    if (h & 8u) { x -= 8u; *reinterpret_cast<uint2 *>(g + x) = *reinterpret_cast<const uint2 *>(image + x); }
    if (h & 4u) { x -= 4u; *reinterpret_cast<u32 *>(g + x) = *reinterpret_cast<const u32 *>(image + x); }
    if (h & 2u) { x -= 2u; *reinterpret_cast<uint16_t *>(g + x) = *reinterpret_cast<const uint16_t *>(image + x); }
    if (h & 1u) { x -= 1u; g[x] = image[x]; }
}
// The 1..15 bytes behind the 16-byte aligned image offset `lo`.
__device__ __forceinline__ void store_tail(const uint8_t *image, uint8_t *g, u32 lo, u32 end) {
    const u32 t = end - lo;
    u32 x = lo;
    if (t & 8u) { *reinterpret_cast<uint2 *>(g + x) = *reinterpret_cast<const uint2 *>(image + x); x += 8u; }
    if (t & 4u) { *reinterpret_cast<u32 *>(g + x) = *reinterpret_cast<const u32 *>(image + x); x += 4u; }
    if (t & 2u) { *reinterpret_cast<uint16_t *>(g + x) = *reinterpret_cast<const uint16_t *>(image + x); x += 2u; }
    if (t & 1u) { g[x] = image[x]; }
}

}  // namespace gts
