// gts_kernels.cu -- the sm_100a kernels of the translation hot path.
//
// Reference behaviour reproduced (file:line in feliixx/gotranseq):
//   k_parse      readSequenceFromFasta  transeq/transeq.go:183-216 (bufio.ScanLines semantics),
//                newEncodedSequence     transeq/encodedSequence.go:17-43
//   k_size       frame lengths / tail rules / trim   transeq/writer.go:85-116,141-169
//   k_translate  translate / translate3Frames / writeAA / newLine   transeq/writer.go:57-157,
//                reverseComplement      transeq/encodedSequence.go:71-97 (never materialised:
//                the fused table holds the reverse-strand residue of every window)
//   k_headers    writeHeader            transeq/writer.go:120-131
//
// tests/gpu_model.py is the executable specification of the index arithmetic used here.
#include <cuda_runtime.h>

#include "gts_device.h"
#include "gts_kernels.h"

namespace gts {

// =====================================================================================
// small helpers
// =====================================================================================
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
__device__ __forceinline__ u32 ld_acquire32(const u32 *p) {
    u32 v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release32(u32 *p, u32 v) {
    asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

// 0x80 in every byte of w that equals the byte replicated in pat
__device__ __forceinline__ u32 eq_bytes(u32 w, u32 pat) {
    u32 z = w ^ pat;
    u32 t = (z & 0x7F7F7F7Fu) + 0x7F7F7F7Fu;
    return ~(t | z | 0x7F7F7F7Fu);
}
// PTX prmt in its default mode: selector nibble bit 3 replicates the chosen byte's sign bit
// (__byte_perm masks that bit away, and costs an extra AND for run-time selectors)
__device__ __forceinline__ u32 prmt(u32 a, u32 b, u32 sel) {
    u32 r;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
    return r;
}

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

// the four 0x80 flags of m as bits 0..3
__device__ __forceinline__ u32 gather4(u32 m) { return (((m >> 7) * 0x00204081u) >> 21) & 0xFu; }


struct Pair {
    u64 a, b;
};

// Decoupled look-back over the tiles of a chained scan (sum of two counters).  Called by all 32
// lanes of one warp of tile t with that tile's aggregate; publishes it, walks back over the
// predecessors' published aggregates / inclusive prefixes, publishes this tile's inclusive
// prefix and returns the exclusive one.  Tiles are numbered by an atomic ticket, so every
// predecessor is running or finished and the wait terminates.
__device__ Pair scan_lookback(u64 *status, u32 t, Pair agg) {
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
__device__ __forceinline__ u32 ld_relaxed32(const u32 *p) {
    u32 v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Two-level variant for the parse tiles.  A one-level look-back advances by one 32-tile window per
// global-memory round trip, which caps the kernel at 32 tiles per microsecond; here a tile waits
// only for the aggregates of the tiles of its own group of 32 (published without any dependency),
// and the chained look-back runs over whole groups, i.e. 1024 tiles per round trip.  Every status
// entry carries its flag in the same (at most 16-byte, naturally aligned) word as its payload, so
// plain volatile accesses are enough -- no fences on the critical path.
// Called by the 32 lanes of one warp; S < 2^14, R < 2^13 per tile.
__device__ __forceinline__ void st_volatile_v2(u64 *p, u64 a, u64 b) {
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
__device__ __forceinline__ void ld_volatile_v2(const u64 *p, u64 &a, u64 &b) {
    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(p) : "memory");
}
__device__ __forceinline__ u64 ld_volatile64(const u64 *p) {
    u64 v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ u32 ld_volatile32(const u32 *p) {
    u32 v;
    asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_volatile32(u32 *p, u32 v) {
    asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_volatile64(u64 *p, u64 v) {
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

__device__ Pair scan_lookback2(const Job &job, u32 t, u32 S, u32 R) {
    const int lane = threadIdx.x & 31;
    const u32 g = t >> 5, r = t & 31u;
    if (lane == 0) st_volatile32(job.p_agg + t, 0x80000000u | (R << 14) | S);
    // issue the first round of both levels at once: the aggregates of the tiles before this one in
    // its group, and the status of the 32 groups before this group
    const bool in_group = (u32)lane < r;
    const i64 gidx0 = (i64)g - 1 - lane;
    u32 v = in_group ? ld_volatile32(job.p_agg + (g << 5) + lane) : 0x80000000u;
    u64 ia = 0, ib = 0, ga = 0;
    if (gidx0 >= 0) {
        ld_volatile_v2(job.p_ginc + 2 * gidx0, ia, ib);
        ga = ld_volatile64(job.p_gagg + gidx0);
    }
    // step 1: sum of the aggregates of the tiles before this one in its group
    while (!(v & 0x80000000u)) v = ld_volatile32(job.p_agg + (g << 5) + lane);
    u32 ps = in_group ? (v & 0x3FFFu) : 0u, pr = in_group ? ((v >> 14) & 0x1FFFu) : 0u;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        ps += __shfl_xor_sync(0xFFFFFFFFu, ps, o);
        pr += __shfl_xor_sync(0xFFFFFFFFu, pr, o);
    }
    const bool anchor = r == 31u;  // the group's last tile publishes the group's aggregate / prefix
    const u64 gs = (u64)ps + S, gr = (u64)pr + R;
    if (anchor && lane == 0) st_volatile64(job.p_gagg + g, kValid | (gr << 32) | gs);
    // step 2: chained look-back over the groups before this one
    Pair excl = {0, 0};
    i64 pos = (i64)g - 1;
    bool first_round = true;
    while (pos >= 0) {
        const i64 idx = pos - lane;
        int kind = 2;  // 0 = not published, 1 = aggregate, 2 = inclusive (or before group 0)
        u64 a = 0, b = 0;
        if (idx >= 0) {
            if (!first_round) {
                ld_volatile_v2(job.p_ginc + 2 * idx, ia, ib);
                ga = ld_volatile64(job.p_gagg + idx);
            }
            if (ia & kValid) {
                a = ia & ~kValid;
                b = ib;
            } else if (ga & kValid) {
                kind = 1;
                a = ga & 0xFFFFFFFFull;
                b = (ga & ~kValid) >> 32;
            } else {
                kind = 0;
            }
        }
        first_round = false;
        const unsigned inc_m = __ballot_sync(0xFFFFFFFFu, kind == 2);
        const unsigned none_m = __ballot_sync(0xFFFFFFFFu, kind == 0);
        const int first_inc = inc_m ? __ffs(inc_m) - 1 : 32;
        const int first_none = none_m ? __ffs(none_m) - 1 : 32;
        if (first_none < first_inc) continue;  // a needed group has not published yet
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
    if (anchor && lane == 0) st_volatile_v2(job.p_ginc + 2 * (u64)g, (excl.a + gs) | kValid, excl.b + gr);
    excl.a += ps;
    excl.b += pr;
    return excl;
}

__device__ __forceinline__ i64 floordiv3(i64 a) { return a >= 0 ? (i64)udiv3((u64)a) : -(i64)udiv3((u64)(-a) + 2); }
__device__ __forceinline__ i64 ceildiv3(i64 a) { return -floordiv3(-a); }
__device__ __forceinline__ int mod3(i64 a) { return (int)(a - 3 * floordiv3(a)); }

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

__device__ __forceinline__ u32 stream_symbol(const u32 *sym, u64 x) {
    const u32 w = sym[x >> 3];
    const u32 i = (u32)x & 7u;
    return (w >> (8 * (i & 3) + 4 * (i >> 2))) & 0xFu;
}

__device__ __forceinline__ bool record_emitted(u64 s, u64 n, u64 R) { return s >= 1 || n > 0 || R == 0; }

// =====================================================================================
// k_parse
// =====================================================================================
enum { LS = 0, SEQ = 1, HDR = 2 };

struct ParseShared {
    alignas(16) uint8_t raw[kParseTile];
    alignas(16) uint8_t sym[kParseTile + 96];  // the tile's symbols, one per byte, tile relative (+ zero tail)
    uint8_t enc[256];
    u32 warp_sum[kParseThreads / 32];
    u64 T0, R0;
    u32 tile;
};

// Symbols and header lines of one 32-byte chunk (scalar walk over its lines; used for the chunks
// the fast path does not take).  Same line rules as tests/gpu_model.py parse().
__device__ __forceinline__ void count_chunk(const uint8_t *raw, int len, u32 nlmask, int state,
                                            bool cr_dropped_at_end, u32 &nsym, u32 &nhdr) {
    int a = 0;
    u32 ns = 0, nh = 0;
    int st = state;
    while (true) {
        const u32 m = a < 32 ? (nlmask >> a) : 0u;
        const int e = m ? a + __ffs(m) - 1 : len;  // index of the next '\n', or the chunk end
        if (a < e) {
            if (st == LS) {
                if (raw[a] == '>') {  // transeq.go:198
                    st = HDR;
                    nh++;
                    ns += 2;
                } else {
                    st = SEQ;
                }
            }
            if (st == SEQ) {
                int b = e;
                // bufio.ScanLines dropCR: a CR right before the '\n' (or as the last byte of the
                // input) is not part of the line
                if (raw[e - 1] == '\r' && (e < len || cr_dropped_at_end)) b--;
                ns += b - a;
            }
        }
        if (e < len) {
            st = LS;
            a = e + 1;
        } else {
            break;
        }
    }
    nsym = ns;
    nhdr = nh;
}

// 0x80 in every byte of x that is '\n'
__device__ __forceinline__ u32 newline_flags(u32 x) {
    return __vabsdiffu4(__vabsdiffu4(x, 0x0A0A0A0Au), 0x80808080u) & 0x80808080u;
}
__device__ __forceinline__ u32 newline_mask16(uint4 v) {
    return gather4(newline_flags(v.x)) | gather4(newline_flags(v.y)) << 4 | gather4(newline_flags(v.z)) << 8 |
           gather4(newline_flags(v.w)) << 12;
}

// Nucleotide codes of four bytes at once (encodedSequence.go:25-41): A/a 1, C/c 2, T/t/U/u 3,
// G/g 4, anything else 0.  The low three bits of (byte ^ 'A') tell the five letters apart; two
// byte-permute table lookups give the letter that index stands for and its code, and the code is
// kept only where the byte (case folded) equals that letter.
__device__ __forceinline__ u32 encode4(u32 x) {
    const u32 w = x ^ 0x41414141u;
    const u32 idx = w & 0x07070707u;
    const u32 sel = prmt(idx | (idx >> 4), 0u, 0x4420);
    const u32 canon = prmt(0xFF02FF00u, 0xFF061514u, sel);  // A0 C2 U4 T5 G6
    const u32 code = prmt(0x00020001u, 0x00040303u, sel);
    const u32 diff = (w & 0xDFDFDFDFu) ^ canon;
    const u32 ok = prmt(__vabsdiffu4(diff, 0x80808080u), 0u, 0xBA98);  // 0xFF where diff == 0
    return code & ok;
}

// remove byte p (0..31) from the 32 bytes in C[0..7]; C[8] supplies the byte shifted in
__device__ __forceinline__ void remove_byte(u32 (&C)[9], int p) {
    const int pw = p >> 2;
    const u32 mid = 0x3210u + ((0x1111u << (4 * (p & 3))) & 0xFFFFu);
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const u32 sel = i < pw ? 0x3210u : (i == pw ? mid : 0x4321u);
        C[i] = prmt(C[i], C[i + 1], sel);
    }
}

__global__ void __launch_bounds__(kParseThreads) k_parse(const __grid_constant__ Job job) {
    __shared__ ParseShared sh;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;

#ifdef GTS_PHASE_TIMING
    u64 tprev = gtime();
#endif
    if (tid == 0) sh.tile = atomicAdd(&job.counters[0], 1u);
    {
        uint8_t c = 0;
        switch (tid) {
            case 'A': case 'a': c = 1; break;
            case 'C': case 'c': c = 2; break;
            case 'T': case 't': case 'U': case 'u': c = 3; break;
            case 'G': case 'g': c = 4; break;
        }
        sh.enc[tid] = c;
    }
    __syncthreads();
    PHASE_MARK(job, 0, tprev);  // ticket
    const u32 t = sh.tile;
    const u64 tile_base = (u64)t * kParseTile;
    const u64 warp_base = tile_base + (u64)warp * 1024u;
    const u64 pos0 = warp_base + (u64)lane * kParseBytesPerThread;
    const int len = pos0 >= job.n ? 0 : (job.n - pos0 < 32 ? (int)(job.n - pos0) : 32);

    // ---- load the chunk, keep a copy of the tile in shared memory ----
    uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
    if (len == 32) {
        const uint4 *p = reinterpret_cast<const uint4 *>(job.in + pos0);
        v0 = __ldg(p);
        v1 = __ldg(p + 1);
    } else if (len > 0) {
        u32 w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int i = 0; i < len; i++) w[i >> 2] |= (u32)job.in[pos0 + i] << (8 * (i & 3));
        v0 = make_uint4(w[0], w[1], w[2], w[3]);
        v1 = make_uint4(w[4], w[5], w[6], w[7]);
    }
    uint8_t *raw = sh.raw + tid * kParseBytesPerThread;
    reinterpret_cast<uint4 *>(raw)[0] = v0;
    reinterpret_cast<uint4 *>(raw)[1] = v1;
    u32 nlmask = newline_mask16(v0) | newline_mask16(v1) << 16;
    if (len < 32) nlmask &= len ? ((1u << len) - 1u) : 0u;

    // ---- start of the line each chunk begins in ----
    // inside the warp: max-scan of "position after the last newline"; before the warp's first
    // byte: backward search for the previous '\n'
    u64 warp_ls = 0;
    {
        u64 hi = warp_base < job.n ? warp_base : 0;  // warps past the end do not search
        while (hi > 0) {
            const i64 start = (i64)hi - 16 * (lane + 1);
            u32 m16 = 0;
            if (start >= 0) m16 = newline_mask16(__ldg(reinterpret_cast<const uint4 *>(job.in + start)));
            const unsigned b = __ballot_sync(0xFFFFFFFFu, m16 != 0);
            if (b) {
                const int l = __ffs(b) - 1;
                const u32 mm = __shfl_sync(0xFFFFFFFFu, m16, l);
                warp_ls = hi - 16ull * (l + 1) + (31 - __clz(mm)) + 1;
                break;
            }
            if (hi <= 512) break;
            hi -= 512;
        }
    }
    const u32 my_last = nlmask ? (u32)(lane * 32 + (31 - __clz(nlmask)) + 1) : 0u;
    u32 inc_max = my_last;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(0xFFFFFFFFu, inc_max, o);
        if (lane >= o) inc_max = max(inc_max, y);
    }
    u32 excl_max = __shfl_up_sync(0xFFFFFFFFu, inc_max, 1);
    if (lane == 0) excl_max = 0;
    const u64 line_start = excl_max ? warp_base + excl_max : warp_ls;

    int state = LS;
    if (len > 0 && line_start != pos0) state = job.in[line_start] == '>' ? HDR : SEQ;
    const bool last_cr = len > 0 && raw[len - 1] == '\r';
    const bool cr_end = last_cr && ((pos0 + len >= job.n) || job.in[pos0 + len] == '\n');
    const bool eof_chunk = len > 0 && pos0 + len == job.n;
    const bool first_thread = (t == 0 && tid == 0);

    // ---- classify the chunk ----
    //   fast:  32 bytes of sequence lines only, at most two bytes to drop (newline, CR)
    //   slow:  anything with a header line start / end in it, or short (end of input), ...
    //   none:  no symbols (inside a header line, or past the end of the input)
    u32 removed = 0;
    bool slow = false, fast = false;
    if (len > 0) {
        if (state == HDR) {
            slow = nlmask != 0 || eof_chunk;  // the header line ends here
        } else {
            slow = (state == LS && raw[0] == '>') || len < 32 || first_thread;
            u32 m = nlmask;
            u32 crd = cr_end ? (1u << (len - 1)) : 0u;
            while (m) {
                const int e = __ffs(m) - 1;
                m &= m - 1;
                if (e + 1 < len && raw[e + 1] == '>') slow = true;
                if (e > 0 && raw[e - 1] == '\r') crd |= 1u << (e - 1);
            }
            removed = nlmask | crd;
            if (__popc(removed) > 2) slow = true;
            fast = !slow;
        }
    }
    u32 nsym = 0, nhdr = 0;
    if (fast) {
        nsym = 32u - (u32)__popc(removed);
    } else if (slow) {
        count_chunk(raw, len, nlmask, state, cr_end, nsym, nhdr);
        if (first_thread) nsym += 2;  // pads of slot 0
    } else if (first_thread) {
        nsym = 2;  // empty input
        slow = true;
    }

    PHASE_MARK(job, 1, tprev);  // load, masks, line start, classify, count
    // ---- scan over the tile ----
    const u32 packed = nsym | nhdr << 16;
    u32 inc = packed;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const u32 y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) sh.warp_sum[warp] = inc;
    __syncthreads();
    u32 excl = inc - packed;
    u32 tile_total = 0;
#pragma unroll
    for (int w = 0; w < kParseThreads / 32; w++) {
        const u32 s = sh.warp_sum[w];
        if (w < warp) excl += s;
        tile_total += s;
    }
    const u32 S_tile = tile_total & 0xFFFFu, R_tile = tile_total >> 16;
    const u32 d = excl & 0xFFFFu;  // tile-relative index of the chunk's first symbol

    // ---- chained scan across tiles (warp 0), overlapped with the fast chunks of the others ----
    PHASE_MARK(job, 2, tprev);  // barrier + tile scan
    if (warp == 0) {
        Pair ex = scan_lookback2(job, t, S_tile, R_tile);
        PHASE_MARK(job, 3, tprev);  // look-back
        if (lane == 0) {
            sh.T0 = ex.a;
            sh.R0 = ex.b;
        }
    }

    // ---- fast chunks: encode, drop the newline / CR bytes, store whole words ----
    {
        // "big" chunks (fast, hence >= 30 symbols) link up: a chunk's last, partial word is
        // completed with the first symbols of the next lane's chunk and stored whole
        const unsigned bigm = __ballot_sync(0xFFFFFFFFu, fast);
        u32 C[9];
        C[8] = 0;
        if (fast) {
            C[0] = encode4(v0.x); C[1] = encode4(v0.y); C[2] = encode4(v0.z); C[3] = encode4(v0.w);
            C[4] = encode4(v1.x); C[5] = encode4(v1.y); C[6] = encode4(v1.z); C[7] = encode4(v1.w);
            u32 rm = removed;
            while (rm) {
                const int p = 31 - __clz(rm);
                rm &= ~(1u << p);
                remove_byte(C, p);
            }
        } else {
#pragma unroll
            for (int i = 0; i < 8; i++) C[i] = 0;
        }
        const u32 next3 = __shfl_down_sync(0xFFFFFFFFu, C[0], 1);
        if (fast) {
            const bool link_next = lane < 31 && ((bigm >> (lane + 1)) & 1u);
            const bool link_prev = lane > 0 && ((bigm >> (lane - 1)) & 1u);
            const u32 k = nsym;  // 30..32
            if (link_next) {      // append the next chunk's first three symbols after symbol k-1
                const u32 shb = 8u * (k & 3u);
                if (k == 32u) {
                    C[8] = next3;
                } else {
                    C[7] |= next3 << shb;
                    C[8] = next3 >> (32u - shb);  // shb is 16 or 24 here
                }
            }
            const u32 a = d & 3u;
            const u32 selA = 0x3210u + 0x1111u * (4u - a);
            u32 D[9];
            D[0] = prmt(0u, C[0], selA);
#pragma unroll
            for (int j = 1; j < 9; j++) D[j] = prmt(C[j - 1], C[j], selA);
            u32 *dstw = reinterpret_cast<u32 *>(sh.sym + (d - a));
            uint8_t *dstb = sh.sym + (d - a);
            // word 0: whole if aligned; otherwise its bytes a..3 are ours unless the previous lane stores them
            if (a == 0) {
                dstw[0] = D[0];
            } else if (!link_prev) {
                if (a <= 1) dstb[1] = (uint8_t)(D[0] >> 8);
                if (a <= 2) dstb[2] = (uint8_t)(D[0] >> 16);
                dstb[3] = (uint8_t)(D[0] >> 24);
            }
#pragma unroll
            for (int j = 1; j < 7; j++) dstw[j] = D[j];  // always whole (k >= 28)
            const u32 jt = (a + k) >> 2, nb = (a + k) & 3u;  // last, partial word and its bytes
            if (jt > 7) dstw[7] = D[7];
            if (nb) {
                const u32 Dt = jt == 7 ? D[7] : D[8];
                if (link_next) {
                    dstw[jt] = Dt;
                } else {
                    dstb[4 * jt] = (uint8_t)Dt;
                    if (nb >= 2) dstb[4 * jt + 1] = (uint8_t)(Dt >> 8);
                    if (nb >= 3) dstb[4 * jt + 2] = (uint8_t)(Dt >> 16);
                }
            }
        }
    }
    PHASE_MARK(job, 4, tprev);  // warp 0's fast chunks
    __syncthreads();  // T0 / R0 known
    PHASE_MARK(job, 5, tprev);  // barrier
    const u64 T0 = sh.T0, R0 = sh.R0;

    // ---- slow chunks: the 32 lanes of the warp take one byte each ----
    {
        unsigned slowm = __ballot_sync(0xFFFFFFFFu, slow);
        while (slowm) {
            const int c = __ffs(slowm) - 1;
            slowm &= slowm - 1;
            const int c_len = __shfl_sync(0xFFFFFFFFu, len, c);
            const int c_state = __shfl_sync(0xFFFFFFFFu, state, c);
            const u32 c_d = __shfl_sync(0xFFFFFFFFu, d, c);
            const u32 c_rec = __shfl_sync(0xFFFFFFFFu, excl >> 16, c);
            const bool c_cr_end = __shfl_sync(0xFFFFFFFFu, (int)cr_end, c) != 0;
            const bool c_eof = __shfl_sync(0xFFFFFFFFu, (int)eof_chunk, c) != 0;
            const bool c_first = (t == 0 && warp == 0 && c == 0);
            const u64 c_pos0 = warp_base + (u64)c * 32u;
            const bool valid = lane < c_len;
            const uint8_t b = valid ? sh.raw[(warp * 32 + c) * 32 + lane] : 0;
            const unsigned lt = (1u << lane) - 1u;
            const unsigned nl = __ballot_sync(0xFFFFFFFFu, valid && b == '\n');
            const unsigned before = nl & lt;
            const int ls = before ? 32 - __clz(before) : -1;  // start of this byte's line, -1: before the chunk
            const int first_idx = ls >= 0 ? ls : 0;
            const uint8_t first = (uint8_t)__shfl_sync(0xFFFFFFFFu, (int)b, first_idx);
            const bool fresh = ls >= 0 || c_state == LS;  // the line starts inside the chunk
            const bool is_hdr = fresh ? first == '>' : c_state == HDR;
            const bool is_nl = valid && b == '\n';
            const bool hstart = valid && !is_nl && fresh && lane == first_idx && b == '>';
            const unsigned hs = __ballot_sync(0xFFFFFFFFu, hstart);
            const uint8_t nextb = (uint8_t)__shfl_down_sync(0xFFFFFFFFu, (int)b, 1);
            const bool next_nl = lane + 1 < c_len ? nextb == '\n' : c_cr_end;
            const bool keep = valid && !is_nl && !is_hdr && !(b == '\r' && next_nl);
            const unsigned keepm = __ballot_sync(0xFFFFFFFFu, keep);
            const u32 lead = c_first ? 2u : 0u;
            const u32 off = lead + (u32)__popc(keepm & lt) + 2u * (u32)__popc(hs & lt);
            if (c_first && lane == 0) {  // slot 0: the bytes before the first header line
                sh.sym[0] = 0;
                sh.sym[1] = 0;
                job.hdr_off[0] = 0;
                job.hdr_end[0] = 0;
                job.dbase[0] = 2;
            }
            if (keep) sh.sym[c_d + off] = sh.enc[b];
            const u64 rec_before = R0 + c_rec + (u64)__popc(hs & lt);  // header lines before this byte
            if (hstart) {
                sh.sym[c_d + off] = 0;
                sh.sym[c_d + off + 1] = 0;
                const u64 rec = rec_before + 1;
                if (rec < job.rec_cap) {
                    job.hdr_off[rec] = c_pos0 + lane;
                    job.dbase[rec] = T0 + c_d + off + 2;
                }
            }
            const uint8_t prevb = (uint8_t)__shfl_up_sync(0xFFFFFFFFu, (int)b, 1);
            if (is_nl && is_hdr) {  // the header line ends here
                const uint8_t pb = lane > 0 ? prevb : job.in[c_pos0 - 1];
                const u64 end = c_pos0 + lane - (pb == '\r' ? 1 : 0);
                if (rec_before < job.rec_cap) job.hdr_end[rec_before] = end;
            }
            if (c_eof && lane == c_len - 1 && !is_nl && is_hdr) {  // header line without a final newline
                const u64 rec = rec_before + (hstart ? 1 : 0);
                const u64 end = job.n - (b == '\r' ? 1 : 0);
                if (rec < job.rec_cap) job.hdr_end[rec] = end;
            }
            // the record slot owning the first symbol of a translate sub-tile that starts in this chunk
            const u32 c_nsym = __shfl_sync(0xFFFFFFFFu, nsym, c);
            const u64 g0 = T0 + c_d;
            const u64 x = (g0 + kTransSub - 1) / kTransSub;
            if (c_nsym > 0 && x * kTransSub < g0 + c_nsym) {
                const u64 xs = x * kTransSub;  // stream index of the sub-tile's first symbol
                const unsigned le = __ballot_sync(0xFFFFFFFFu, hstart && T0 + c_d + off + 2 <= xs);
                if (lane == 0) job.t_hint[x] = (u32)(R0 + c_rec + (u64)__popc(le));
            }
        }
    }
    // sub-tile starts inside fast chunks (no header line there: the slot is the one open at the chunk)
    if (fast) {
        const u64 g0 = T0 + d;
        const u64 x = (g0 + kTransSub - 1) / kTransSub;
        if (x * kTransSub < g0 + nsym) job.t_hint[x] = (u32)(R0 + (excl >> 16));
    }
    if (tid < 64) sh.sym[S_tile + tid] = 0;  // zero tail so that the last partial word packs zeros
    const bool last_tile = (t + 1 == job.ntiles_parse);
    if (last_tile && tid == 0) {
        const u64 S_total = T0 + S_tile, R_total = R0 + R_tile;
        job.totals->total_sym = S_total;
        job.totals->n_headers = R_total;
        if (R_total + 2 <= job.rec_cap) {
            job.dbase[R_total + 1] = S_total + 2;
        } else {
            atomicOr(&job.totals->flags, kFlagRecOverflow);
        }
        // the two end pads may open a translate sub-tile of their own
        const u64 x = (S_total + kTransSub - 1) / kTransSub;
        if (x * kTransSub < S_total + 2) job.t_hint[x] = (u32)R_total;
    }
    PHASE_MARK(job, 6, tprev);  // slow chunks, hints
    __syncthreads();  // all symbols staged
    PHASE_MARK(job, 7, tprev);  // barrier

    // ---- pack two symbols per byte and write the stream words ----
    // stream word w of the tile holds the tile's symbols 8w-a .. 8w-a+7 (a = symbols of the stream's
    // last incomplete word that the previous tiles left pending)
    const u32 a_pend = (u32)T0 & 7u;
    const u32 tot = a_pend + S_tile;
    const u32 nw = tot >> 3, rem = tot & 7u;
    u32 *gw = job.sym + ((T0 - a_pend) >> 3);
    {
        const u32 sh4 = (8u - a_pend) & 3u;                // byte misalignment of word w's first symbol
        const u32 sel = 0x3210u + 0x1111u * sh4;
        for (u32 w = tid + (a_pend ? 1u : 0u); w < nw; w += kParseThreads) {
            const u32 s0 = 8u * w - a_pend;                // tile-relative index of the word's first symbol
            const u32 base = s0 & ~3u;
            const u32 x0 = *reinterpret_cast<const u32 *>(sh.sym + base);
            const u32 x1 = *reinterpret_cast<const u32 *>(sh.sym + base + 4);
            const u32 x2 = *reinterpret_cast<const u32 *>(sh.sym + base + 8);
            const u32 lo = prmt(x0, x1, sel), hi = prmt(x1, x2, sel);
            gw[w] = lo | (hi << 4);
        }
    }
    PHASE_MARK(job, 8, tprev);  // pack
    if (tid == 0) {
        // hand-over of the symbols that do not fill a stream word
        const bool own = rem <= S_tile;  // the pending symbols are all ours: publish before waiting
        u32 pv = 0;
        if (own && !last_tile) {
            u32 v = 0x80000000u;
            for (u32 k = 0; k < rem; k++) v |= (u32)sh.sym[S_tile - 1 - k] << (4 * k);
            st_volatile32(job.p_pend + t, v);
        }
        if (a_pend) {
            do {
                pv = ld_volatile32(job.p_pend + (t - 1));
            } while (!(pv & 0x80000000u));
        }
        // symbol i (0..7) of the tile's first stream word: pending ones first, then ours
        auto first_word_sym = [&](u32 i) -> u32 {
            if (i < a_pend) return (pv >> (4 * (a_pend - 1 - i))) & 0xFu;
            const u32 j = i - a_pend;
            return j < S_tile ? sh.sym[j] : 0u;
        };
        if (a_pend && (nw >= 1 || last_tile)) {
            u32 word = 0;
            for (u32 i = 0; i < 8; i++) word |= first_word_sym(i) << (8 * (i & 3) + 4 * (i >> 2));
            gw[0] = word;
        }
        if (!own && !last_tile) {  // fewer than 8 symbols in all: pass the predecessors' on as well
            u32 v = 0x80000000u;
            for (u32 k = 0; k < rem; k++) v |= first_word_sym(rem - 1 - k) << (4 * k);
            st_volatile32(job.p_pend + t, v);
        }
        if (last_tile) {
            // the last partial word (zero padded) and the end pads
            for (u32 w = (nw == 0 && a_pend) ? 1u : nw; w < nw + 4; w++) {
                u32 word = 0;
                for (u32 i = 0; i < 8; i++) {
                    const i64 j = (i64)(8u * w + i) - (i64)a_pend;
                    const u32 s = (j >= 0 && j < (i64)S_tile) ? sh.sym[j] : 0u;
                    word |= s << (8 * (i & 3) + 4 * (i >> 2));
                }
                gw[w] = word;
            }
        }
    }
    PHASE_MARK(job, 9, tprev);  // pending hand-over
}

// =====================================================================================
// k_size
// =====================================================================================
// Residues kept by --trim (writer.go:141-147,159-163): walk each requested frame backwards from
// its last residue until one is neither 'X' nor '*'.
__device__ void trimmed_lengths(const Job &job, const uint16_t *lut, u64 D, u64 n, u64 L[6]) {
    for (int f = 0; f < 6; f++) {
        if (!((job.frame_mask >> f) & 1u)) continue;
        u64 j = L[f];
        const u64 p = f >= 3 ? (u64)reverse_start(n, f - 3, job.alternative != 0) : 0;
        while (j > 0) {
            // window start q of residue j-1; its symbols sit at stream index D+q .. D+q+2
            const i64 q = f < 3 ? (i64)(f + 3 * (j - 1)) : (i64)n - 3 - (i64)p - 3 * (i64)(j - 1);
            const u64 e = (u64)((i64)D + q);
            const u32 idx = stream_symbol(job.sym, e) + 5 * stream_symbol(job.sym, e + 1) +
                            25 * stream_symbol(job.sym, e + 2);
            const u32 r = lut[idx < 125 ? idx : 0];
            const u32 aa = f < 3 ? (r & 0xFFu) : (r >> 8);
            if (aa != 'X' && aa != '*') break;
            j--;
        }
        L[f] = j;
    }
}

__global__ void __launch_bounds__(kSizeThreads) k_size(const __grid_constant__ Job job) {
    __shared__ uint16_t lut[128];
    __shared__ u64 warp_sum[kSizeThreads / 32];
    __shared__ u64 warp_nt[kSizeThreads / 32];
    __shared__ u32 warp_rec[kSizeThreads / 32];
    __shared__ u64 s_excl;
    __shared__ u32 s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 128) lut[tid] = job.lut.v[tid];
    const u64 R = job.totals->n_headers;
    if (R + 2 > job.rec_cap) return;  // record table overflow: the host retries with a larger one
    const u64 nslots = R + 1;
    const u64 ntiles = (nslots + kSizeTile - 1) / kSizeTile;
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(&job.counters[1], 1u);
        __syncthreads();
        const u64 t = s_tile;
        if (t >= ntiles) return;
        u64 size[kSizePerThread];
        RecInfo ri[kSizePerThread];
        u64 sum = 0, nt = 0;
        u32 nrec = 0;
#pragma unroll
        for (int i = 0; i < kSizePerThread; i++) {
            const u64 s = t * kSizeTile + (u64)tid * kSizePerThread + i;
            size[i] = 0;
            if (s < nslots) {
                const u64 D = job.dbase[s];
                const u64 n = job.dbase[s + 1] - 2 - D;
                u64 L[6];
                frame_lengths(n, job.alternative != 0, L);
                if (job.trim) {
                    trimmed_lengths(job, lut, D, n, L);
#pragma unroll
                    for (int f = 0; f < 6; f++) job.trim_len[s * 6 + f] = (u32)L[f];
                }
                const u64 H = job.hdr_end[s] - job.hdr_off[s];
                const bool em = record_emitted(s, n, R);
                if (em) {
#pragma unroll
                    for (int f = 0; f < 6; f++)
                        if ((job.frame_mask >> f) & 1u) size[i] += run_size(H, L[f]);
                    nt += n;
                    nrec++;
                }
                ri[i].dbase = D;
                ri[i].n = n;
                ri[i].H = (u32)H;
                ri[i].emitted = em ? 1u : 0u;
            }
            sum += size[i];
        }
        // block scan of the per-thread sums
        u64 inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nt += __shfl_xor_sync(0xFFFFFFFFu, nt, o);
            nrec += __shfl_xor_sync(0xFFFFFFFFu, nrec, o);
        }
        if (lane == 31) warp_sum[warp] = inc;
        if (lane == 0) {
            warp_nt[warp] = nt;
            warp_rec[warp] = nrec;
        }
        __syncthreads();
        u64 excl = inc - sum, tile_total = 0, tile_nt = 0;
        u32 tile_rec = 0;
        for (int w = 0; w < kSizeThreads / 32; w++) {
            if (w < warp) excl += warp_sum[w];
            tile_total += warp_sum[w];
            tile_nt += warp_nt[w];
            tile_rec += warp_rec[w];
        }
        if (warp == 0) {
            Pair agg = {tile_total, tile_nt};
            Pair ex = scan_lookback(job.s_status, (u32)t, agg);
            if (lane == 0) {
                s_excl = ex.a;
                atomicAdd(&job.totals->records, (u64)tile_rec);
                if (t + 1 == ntiles) {
                    const u64 total = ex.a + tile_total;
                    job.totals->out_total = total;
                    job.totals->nucleotides = ex.b + tile_nt;
                    RecInfo end;  // sentinel
                    end.out_off = total;
                    end.dbase = job.dbase[nslots];
                    end.n = 0; end.H = 0; end.emitted = 0;
                    job.rec[nslots] = end;
                    if (total > job.out_cap) atomicOr(&job.totals->flags, kFlagOutOverflow);
                }
            }
        }
        __syncthreads();
        u64 off = s_excl + excl;
#pragma unroll
        for (int i = 0; i < kSizePerThread; i++) {
            const u64 s = t * kSizeTile + (u64)tid * kSizePerThread + i;
            if (s < nslots) {
                ri[i].out_off = off;
                job.rec[s] = ri[i];
            }
            off += size[i];
        }
    }
}

// =====================================================================================
// k_translate
// =====================================================================================
struct Run {        // the part of one (record, frame) body that a sub-tile writes
    u64 g0;         // output offset of its first byte
    int src;        // byte offset of its first residue in the warp's phase arrays
    u32 img;        // byte offset of its first byte in the warp's output image
    uint16_t naa;   // residues
    uint16_t nbytes;  // residues + newlines
    uint8_t col0;   // column (0..59) of the first residue
    uint8_t flags;  // bit 0: ends the frame (a final newline follows), bit 1: reverse strand
    uint16_t pad;
};

struct RunTable {    // the run table of one batch of record pieces of one sub-tile
    alignas(16) Run runs[kRunsPerBatch];
    u32 line_pre[kRunsPerBatch + 1];
    u32 more;        // record pieces of the sub-tile left after this batch
    u32 next_slot;   // first record slot of the next batch
};

struct WarpShared {  // private to one translating warp
    alignas(16) uint8_t phase[6 * kPhaseStride];  // [forward 0..2 | reverse 0..2][guard|residues|guard]
    alignas(16) uint8_t image[kImageBytes];       // the batch's output bytes, 16-byte congruent to global
    RunTable rt[2];  // the builder warp fills one for the next tile while the other is in use
};

struct TransShared {
    WarpShared w[kTransWarps];
    alignas(16) uint16_t lut[128];
};

__device__ __forceinline__ u32 lds32(const uint8_t *base, u32 off) {
    return *reinterpret_cast<const u32 *>(base + off);
}

// Copy one 60-column line of a run (or the part of it inside the sub-tile) from the phase arrays
// into the output image, newline included.  Whole aligned image words are written: a lane owns
// the words whose first byte lies in its line (for a run's first line also the word holding the
// line's first byte), so words that straddle two lines carry the start of the next line, and the
// bytes written outside the run are image padding that is never copied out.
__device__ __forceinline__ void copy_line(WarpShared &ws, const Run &r, u32 ln) {
    const uint8_t *ph = ws.phase;
    const u32 col0 = r.col0, naa = r.naa;
    const u32 start = ln == 0 ? 0 : (60 - col0) + 60 * (ln - 1);
    const u32 cnt = min(60u - (ln == 0 ? col0 : 0u), naa - start);
    // the newline after column 60; a frame's final newline elsewhere is patched in by the run's
    // copy-out lane (it may fall into a word owned by the previous line)
    const bool nl = ln == 0 ? col0 + cnt == 60 : cnt == 60;
    const u32 d0 = r.img + start + ln;           // image offset of the line's first byte
    const u32 dend = d0 + cnt + (nl ? 1u : 0u);  // one past its last byte
    const u32 w_lo = ln == 0 ? (d0 & ~3u) : ((d0 + 3u) & ~3u);
    const u32 w_hi = ((dend - 1u) & ~3u) + 4u;
    const int nw = (int)(w_hi - w_lo) >> 2;      // <= 16, may be <= 0 for a short last line
    const int rel = (int)w_lo - (int)d0;         // line-relative offset of the first owned byte (-3..3)
    u32 *dst = reinterpret_cast<u32 *>(ws.image + w_lo);
    const bool rev = (r.flags & 2) != 0;
    if (!rev) {
        const u32 A = (u32)(r.src + (int)start + rel);  // source byte of the first owned byte
        const u32 base = A & ~3u;
        const u32 sel = 0x3210u + 0x1111u * (A & 3u);
        u32 lo = lds32(ph, base);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const u32 hi = lds32(ph, base + 4u * (k + 1));
            if (k < nw) dst[k] = prmt(lo, hi, sel);
            lo = hi;
        }
    } else {
        const u32 A = (u32)(r.src - (int)start - rel);  // residues are read downwards
        const u32 base = (A - 3u) & ~3u;
        const u32 sel = 0x0123u + 0x1111u * ((A - 3u) & 3u);
        u32 hi = lds32(ph, base + 4u);
#pragma unroll
        for (int k = 0; k < 16; k++) {
            const u32 lo = lds32(ph, base - 4u * k);
            if (k < nw) dst[k] = prmt(lo, hi, sel);
            hi = lo;
        }
    }
    if (nl) {
        // the word holding line offset cnt: residues, '\n' at byte t, then the next line's start
        const int kn = ((int)cnt - rel) >> 2;
        const int o = rel + 4 * kn;
        const u32 t = (u32)((int)cnt - o);
        u32 X;
        if (!rev) {
            const u32 A2 = (u32)(r.src + (int)start + o);
            X = prmt(lds32(ph, A2 & ~3u), lds32(ph, (A2 & ~3u) + 4u), 0x3210u + 0x1111u * (A2 & 3u));
        } else {
            const u32 A2 = (u32)(r.src - (int)start - o) - 3u;
            X = prmt(lds32(ph, A2 & ~3u), lds32(ph, (A2 & ~3u) + 4u), 0x0123u + 0x1111u * (A2 & 3u));
        }
        // insert the newline at byte t, shifting the later bytes up (selector nibble 4 = '\n')
        const u32 selnl = (u32)(0x4210241021402104ull >> (16u * t));
        dst[kn] = prmt(X, 0x0Au, selnl);
    }
}

__device__ __forceinline__ void bulk_store(void *gdst, const void *ssrc, u32 bytes) {
    const u32 saddr = (u32)__cvta_generic_to_shared(ssrc);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(saddr), "r"(bytes)
                 : "memory");
}

// Run table of one batch of record pieces of one sub-tile.  Called by a group of `width` (4 or 32)
// consecutive lanes, lane `gl` of the group handling record slot s0 + gl (gl < kRecBatch); lane
// kRecBatch of the group, when present, only looks whether the sub-tile has more record pieces.
__device__ __forceinline__ void build_runs(const Job &job, RunTable &ws, u64 R, u64 T0, u64 Tend, bool sub_valid,
                                           u64 s0, int gl, int width, uintptr_t out_mis) {
    const u64 s = s0 + gl;
    Run runs[6];
#pragma unroll
    for (int f = 0; f < 6; f++) {
        runs[f].g0 = 0; runs[f].src = 0; runs[f].img = 0; runs[f].naa = 0; runs[f].nbytes = 0;
        runs[f].col0 = 0; runs[f].flags = 0; runs[f].pad = 0;
    }
    RecInfo ri;
    ri.dbase = ~0ull; ri.n = 0; ri.H = 0; ri.emitted = 0; ri.out_off = 0;
    u64 next_dbase = ~0ull;  // stream base of the record after this lane's
    if (sub_valid && gl < kRecBatch && s <= R) {
        const uint4 *rp = reinterpret_cast<const uint4 *>(job.rec + s);
        const uint4 a = rp[0], b = rp[1];
        ri.out_off = (u64)a.x | (u64)a.y << 32;
        ri.dbase = (u64)a.z | (u64)a.w << 32;
        ri.n = (u64)b.x | (u64)b.y << 32;
        ri.H = b.z;
        ri.emitted = b.w;
        if (gl == kRecBatch - 1 && s + 1 <= R) next_dbase = job.rec[s + 1].dbase;
    }
    const bool valid = ri.dbase < Tend;  // slots past R keep dbase = ~0
    u32 nlines = 0, img_need = 0;
    if (valid && ri.emitted) {
        const u64 D = ri.dbase, n = ri.n, H = ri.H;
        u64 L[6];
        frame_lengths(n, job.alternative != 0, L);
        if (job.trim) {
#pragma unroll
            for (int f = 0; f < 6; f++) L[f] = job.trim_len[s * 6 + f];
        }
        // windows (by start q) of this record whose last symbol lies in the sub-tile
        const i64 dq = (i64)D - (i64)T0;  // local index of the record's first symbol
        const i64 q_lo = -dq - 2 > -2 ? -dq - 2 : -2;
        const i64 q_hi = ((i64)n - 1 < (i64)(Tend - T0) - 1 - dq - 2) ? (i64)n - 1 : (i64)(Tend - T0) - 1 - dq - 2;
        u64 run_base = ri.out_off;
#pragma unroll
        for (int f = 0; f < 6; f++) {
            if (!((job.frame_mask >> f) & 1u)) continue;
            const u64 Lf = L[f];
            const u64 body = run_base + H + 3;
            run_base += run_size(H, Lf);
            i64 j_lo, j_hi, el0;
            if (f < 3) {
                const i64 lo = q_lo > 0 ? q_lo : 0;
                j_lo = ceildiv3(lo - f);
                if (j_lo < 0) j_lo = 0;
                j_hi = q_hi >= f ? floordiv3(q_hi - f) + 1 : 0;
                if (j_hi > (i64)Lf) j_hi = (i64)Lf;
                el0 = dq + f + 2 + 3 * j_lo;
            } else {
                const i64 Q = (i64)n - 3 - reverse_start(n, f - 3, job.alternative != 0);
                j_lo = ceildiv3(Q - q_hi);
                if (j_lo < 0) j_lo = 0;
                j_hi = Q >= q_lo ? floordiv3(Q - q_lo) + 1 : 0;
                if (j_hi > (i64)Lf) j_hi = (i64)Lf;
                el0 = dq + Q - 3 * j_lo + 2;
            }
            if (j_hi <= j_lo) continue;
            const u32 el = (u32)el0;  // 0 <= el0 < kTransSub
            const u32 m0 = el / 3u, phi = el - 3u * m0;
            const u64 line0 = udiv60((u64)j_lo);
            const u32 col0 = (u32)((u64)j_lo - 60 * line0);
            const u32 naa = (u32)(j_hi - j_lo);
            const bool fin = (u64)j_hi == Lf;
            const u32 q60 = (col0 + naa) / 60u;
            const u32 nb = naa + q60 + ((fin && (col0 + naa) != 60u * q60) ? 1u : 0u);
            runs[f].g0 = body + (u64)j_lo + line0;
            runs[f].src = (int)(((f < 3 ? 0u : 3u) + phi) * kPhaseStride + kPhaseGuard + m0);
            runs[f].naa = (uint16_t)naa;
            runs[f].nbytes = (uint16_t)nb;
            runs[f].col0 = (uint8_t)col0;
            runs[f].flags = (uint8_t)((fin ? 1 : 0) | (f < 3 ? 0 : 2));
            nlines += (col0 + naa + 59u) / 60u;
            // image space: 16-byte congruent start, then room for the last word's overhang
            const u32 mis = (u32)((out_mis + runs[f].g0) & 15u);
            img_need += (mis + nb + 3u + 15u) & ~15u;
        }
    }
    // exclusive scans over the group's records: line jobs and image space
    u32 lp = nlines, ip = img_need;
#pragma unroll
    for (int o = 1; o < kRecBatch; o <<= 1) {
        const u32 y = __shfl_up_sync(0xFFFFFFFFu, lp, o, width);
        const u32 z = __shfl_up_sync(0xFFFFFFFFu, ip, o, width);
        if (gl >= o) { lp += y; ip += z; }
    }
    const u32 total_lines = __shfl_sync(0xFFFFFFFFu, lp, kRecBatch - 1, width);
    lp -= nlines;
    ip -= img_need;
    if (gl < kRecBatch) {
#pragma unroll
        for (int f = 0; f < 6; f++) {
            ws.line_pre[gl * 6 + f] = lp;
            if (runs[f].naa) {
                const u32 mis = (u32)((out_mis + runs[f].g0) & 15u);
                runs[f].img = ip + mis;
                ip += (mis + runs[f].nbytes + 3u + 15u) & ~15u;
                lp += (runs[f].col0 + runs[f].naa + 59u) / 60u;
            }
            ws.runs[gl * 6 + f] = runs[f];
        }
        if (gl == 0) ws.line_pre[kRunsPerBatch] = total_lines;
        if (gl == kRecBatch - 1) {
            ws.more = (valid && next_dbase < Tend) ? 1u : 0u;
            ws.next_slot = (u32)(s + 1);
        }
    }
}

// Copy the bytes [lo, hi) of the image (hi - lo < 16, both inside one aligned 16-byte slot) to
// global memory with at most four naturally aligned stores; g is the global address of image byte 0.
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

// Persistent CTAs.  Warps 0..kTransWarps-1 translate one sub-tile each per CTA tile (48 symbols per
// lane); the extra warp builds, one tile ahead, the run table of every sub-tile's first record
// batch.  One CTA barrier per tile hands the tables over; every warp then formats and stores its
// own sub-tile.
__global__ void __launch_bounds__(kTransBlock, 3) k_translate(const __grid_constant__ Job job) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TransShared &sh = *reinterpret_cast<TransShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool builder = warp == kTransWarps;

    const u64 R = job.totals->n_headers;
    const u64 S2 = job.totals->total_sym + 2;
    if (job.totals->flags != 0) return;
    const u64 ntiles = (S2 + kTransTile - 1) / kTransTile;
    const uintptr_t out_mis = (uintptr_t)job.out & 15u;

    if (builder) {
        // 4 lanes per sub-tile: the first kRecBatch record pieces of each
        const int w = lane >> 2, gl = lane & 3;
        u32 it = 0;
        for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            const u64 sub = tile * kTransWarps + w;
            const u64 T0 = sub * kTransSub;
            const bool sub_valid = T0 < S2;
            const u64 Tend = T0 + kTransSub < S2 ? T0 + kTransSub : S2;
            const u64 s0 = sub_valid ? job.t_hint[sub] : 0;
            build_runs(job, sh.w[w].rt[it & 1], R, T0, Tend, sub_valid, s0, gl, 4, out_mis);
            __syncthreads();  // hands tile `it` over
        }
        return;
    }

    if (tid < 128) sh.lut[tid] = job.lut.v[tid];
    asm volatile("bar.sync 1, %0;" ::"r"(kTransWarps * 32) : "memory");  // table staged (translating warps only)
    const uint8_t *lutb = reinterpret_cast<const uint8_t *>(sh.lut);
    WarpShared &ws = sh.w[warp];

    // symbols of this lane in the CTA's first tile: 48 = 6 stream words, plus the word before for
    // (e-2, e-1)
    u32 w6[6] = {0, 0, 0, 0, 0, 0}, prev = 0;
    auto load_symbols = [&](u64 tile) {
        const u64 sub = tile * kTransWarps + warp;
        const u64 T0 = sub * kTransSub;
        if (tile < ntiles && T0 < S2) {
            const u32 *wp = job.sym + (T0 >> 3) + 6 * lane;
            const uint2 x0 = *reinterpret_cast<const uint2 *>(wp);
            const uint2 x1 = *reinterpret_cast<const uint2 *>(wp + 2);
            const uint2 x2 = *reinterpret_cast<const uint2 *>(wp + 4);
            prev = (sub == 0 && lane == 0) ? 0u : wp[-1];
            w6[0] = x0.x; w6[1] = x0.y; w6[2] = x1.x; w6[3] = x1.y; w6[4] = x2.x; w6[5] = x2.y;
        }
    };
    load_symbols(blockIdx.x);

    u32 it = 0;
    for (u64 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        {
            u32 W[13];  // W[k+1] = symbols 4k..4k+3 of the chunk, one per byte; W[0] = the 4 before
            W[0] = (prev >> 4) & 0x0F0F0F0Fu;
#pragma unroll
            for (int g = 0; g < 6; g++) {
                W[1 + 2 * g] = w6[g] & 0x0F0F0F0Fu;
                W[2 + 2 * g] = (w6[g] >> 4) & 0x0F0F0F0Fu;
            }
            // byte i of IDX[k] = 2 * (s[e-2] + 5 s[e-1] + 25 s[e]) for e = 4k+i: the byte offset of
            // the window's entry in the 16-bit table
            u32 IDX[12];
#pragma unroll
            for (int k = 0; k < 12; k++) {
                const u32 A = prmt(W[k], W[k + 1], 0x5432);
                const u32 B = prmt(W[k], W[k + 1], 0x6543);
                IDX[k] = A * 2u + B * 10u + W[k + 1] * 50u;
            }
            load_symbols(tile + gridDim.x);  // prefetch the next tile's symbols
#pragma unroll
            for (int phi = 0; phi < 3; phi++) {
                u32 Fw[4], Rw[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    u32 r[4];
#pragma unroll
                    for (int x = 0; x < 4; x++) {
                        const int i = phi + 3 * (4 * g + x);  // window of the chunk
                        const u32 off = prmt(IDX[i >> 2], 0u, 0x4440u + (i & 3));
                        r[x] = *reinterpret_cast<const uint16_t *>(lutb + off);
                    }
                    const u32 x01 = prmt(r[0], r[1], 0x5140);
                    const u32 x23 = prmt(r[2], r[3], 0x5140);
                    Fw[g] = prmt(x01, x23, 0x5410);
                    Rw[g] = prmt(x01, x23, 0x7632);
                }
                *reinterpret_cast<uint4 *>(ws.phase + phi * kPhaseStride + kPhaseGuard + 16 * lane) =
                    make_uint4(Fw[0], Fw[1], Fw[2], Fw[3]);
                *reinterpret_cast<uint4 *>(ws.phase + (3 + phi) * kPhaseStride + kPhaseGuard + 16 * lane) =
                    make_uint4(Rw[0], Rw[1], Rw[2], Rw[3]);
            }
        }
        __syncthreads();  // this tile's run tables are complete (and the warp's phase arrays)

        // ---- format: every (record piece, frame) of the sub-tile becomes a run of 60-column lines ----
        const u64 sub = tile * kTransWarps + warp;
        const u64 T0 = sub * kTransSub;
        if (T0 >= S2) continue;
        const u64 Tend = T0 + kTransSub < S2 ? T0 + kTransSub : S2;
        RunTable &rt = ws.rt[it & 1];
        while (true) {
            const u32 total = rt.line_pre[kRunsPerBatch];
            const u32 more = rt.more;
            const u64 next_slot = rt.next_slot;
            for (u32 job_i = lane; job_i < total; job_i += 32) {
                // run r with line_pre[r] <= job_i < line_pre[r+1] (kRunsPerBatch = 24 entries)
                u32 lo = rt.line_pre[16] <= job_i ? 16u : 0u;
                if (lo + 8u < kRunsPerBatch && rt.line_pre[lo + 8u] <= job_i) lo += 8u;
                if (rt.line_pre[lo + 4u] <= job_i) lo += 4u;
                if (rt.line_pre[lo + 2u] <= job_i) lo += 2u;
                if (rt.line_pre[lo + 1u] <= job_i) lo += 1u;
                copy_line(ws, rt.runs[lo], job_i - rt.line_pre[lo]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // image writes -> visible to the bulk copy engine
            __syncwarp();                                                  // image complete
            if (lane < kRunsPerBatch) {
                const Run r = rt.runs[lane];
                if (r.naa) {
                    const u32 lo16 = (r.img + 15u) & ~15u, end = r.img + r.nbytes, hi16 = end & ~15u;
                    uint8_t *g = job.out + r.g0 - r.img;  // global address of image byte 0 (16-byte congruent)
                    if (r.flags & 1) {                    // the frame's final newline (writer.go:164-166)
                        ws.image[end - 1u] = '\n';
                        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    }
                    if (hi16 > lo16) {
                        bulk_store(g + lo16, ws.image + lo16, hi16 - lo16);
                        store_edge(ws.image, g, r.img, lo16);
                        store_edge(ws.image, g, hi16, end);
                    } else if (end <= lo16) {
                        store_edge(ws.image, g, r.img, end);
                    } else {  // straddles one slot boundary without filling a slot
                        store_edge(ws.image, g, r.img, lo16);
                        store_edge(ws.image, g, lo16, end);
                    }
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // image may be reused / warp may exit
            }
            if (!more) break;
            __syncwarp();
            build_runs(job, rt, R, T0, Tend, true, next_slot, lane, 32, out_mis);
            __syncwarp();
        }
    }
}

// =====================================================================================
// k_headers
// =====================================================================================
__global__ void __launch_bounds__(256) k_headers(const __grid_constant__ Job job) {
    const int lane = threadIdx.x & 31;
    const u64 R = job.totals->n_headers;
    if (job.totals->flags != 0) return;
    const u64 nwarps = (u64)gridDim.x * (blockDim.x >> 5);
    for (u64 s = (u64)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); s <= R; s += nwarps) {
        const RecInfo ri = job.rec[s];
        if (!ri.emitted) continue;
        const u64 n = ri.n;
        const u64 H = ri.H;
        const uint8_t *hdr = job.in + job.hdr_off[s];
        u64 L[6];
        frame_lengths(n, job.alternative != 0, L);
        if (job.trim) {
#pragma unroll
            for (int f = 0; f < 6; f++) L[f] = job.trim_len[s * 6 + f];
        }
        // index of the first space (writer.go:121), H if there is none
        u64 sp = H;
        for (u64 i0 = 0; i0 < H; i0 += 32) {
            const u64 i = i0 + lane;
            const unsigned b = __ballot_sync(0xFFFFFFFFu, i < H && hdr[i] == ' ');
            if (b) {
                sp = i0 + __ffs(b) - 1;
                break;
            }
        }
        u64 run_base = ri.out_off;
        for (int f = 0; f < 6; f++) {
            if (!((job.frame_mask >> f) & 1u)) continue;
            uint8_t *dst = job.out + run_base;
            for (u64 i = lane; i < H + 3; i += 32) {
                uint8_t c;
                if (i < sp) c = hdr[i];
                else if (i == sp) c = '_';
                else if (i == sp + 1) c = (uint8_t)('1' + f);
                else if (i < H + 2) c = hdr[i - 2];
                else c = '\n';
                dst[i] = c;
            }
            run_base += run_size(H, L[f]);
        }
    }
}

// =====================================================================================
// launch
// =====================================================================================
cudaError_t init_kernels() {
    return cudaFuncSetAttribute(k_translate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(TransShared));
}

cudaError_t launch_pass(const Job &job, u64 max_symbols, int sm_count, cudaStream_t stream, bool sizes_only,
                        unsigned long long *launches, cudaEvent_t *ev) {
    k_parse<<<job.ntiles_parse, kParseThreads, 0, stream>>>(job);
    if (ev) cudaEventRecord(ev[1], stream);
    u64 size_tiles = (job.rec_cap + kSizeTile - 1) / kSizeTile;
    u32 size_grid = (u32)(size_tiles < (u64)sm_count * 4 ? size_tiles : (u64)sm_count * 4);
    if (size_grid == 0) size_grid = 1;
    k_size<<<size_grid, kSizeThreads, 0, stream>>>(job);
    if (ev) cudaEventRecord(ev[2], stream);
    *launches += 2;
    if (!sizes_only) {
        const u64 trans_tiles = (max_symbols + kTransTile - 1) / kTransTile;
        const u32 trans_grid = (u32)(trans_tiles < (u64)sm_count * 3 ? trans_tiles : (u64)sm_count * 3);
        k_translate<<<trans_grid, kTransBlock, sizeof(TransShared), stream>>>(job);
        if (ev) cudaEventRecord(ev[3], stream);
        k_headers<<<sm_count * 4, 256, 0, stream>>>(job);
        if (ev) cudaEventRecord(ev[4], stream);
        *launches += 2;
    }
    return cudaGetLastError();
}

cudaError_t launch_emit(const Job &job, u64 max_symbols, int sm_count, cudaStream_t stream,
                        unsigned long long *launches) {
    const u64 trans_tiles = (max_symbols + kTransTile - 1) / kTransTile;
    const u32 trans_grid = (u32)(trans_tiles < (u64)sm_count * 3 ? trans_tiles : (u64)sm_count * 3);
    k_translate<<<trans_grid, kTransBlock, sizeof(TransShared), stream>>>(job);
    k_headers<<<sm_count * 4, 256, 0, stream>>>(job);
    *launches += 2;
    return cudaGetLastError();
}

}  // namespace gts
