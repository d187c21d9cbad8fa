// gts_parse.cu -- k_parse: raw FASTA bytes -> per-tile symbol slots, header events, tile counts.
//
// Reference behaviour reproduced (feliixx/gotranseq):
//   readSequenceFromFasta  transeq/transeq.go:183-216 with bufio.ScanLines semantics (lines end at
//                          '\n', one trailing '\r' is dropped, empty lines are skipped, a line whose
//                          first byte is '>' starts a record)
//   newEncodedSequence     transeq/encodedSequence.go:17-43 (A/a 1, C/c 2, T/t/U/u 3, G/g 4, else 0)
// tests/gpu_model.py parse() is the executable specification of the line rules used here.
//
// Every tile is independent: its symbols are compacted inside its own slot, the header lines it
// finds are appended to an unordered event list, and the stream offsets are assigned afterwards by
// k_tile_scan / k_records.
#include "gts_common.cuh"
#include "gts_kernels.h"

#ifndef GTS_PARSE_CTAS
#define GTS_PARSE_CTAS 8
#endif

namespace gts {

enum { LS = 0, SEQ = 1, HDR = 2 };

struct ParseShared {
    alignas(16) uint8_t raw[kParseTile];
    alignas(16) uint8_t sym[kSlotSyms + 32];  // the tile's symbols, one per byte (+ zero tail)
    uint8_t enc[256];
    u32 warp_sum[kParseThreads / 32];
    u32 warp_nl[kParseThreads / 32];  // first newline << 16 | last newline + 1 of each warp (tile relative)
    volatile u32 ev_base;  // first slot of the tile's header events, ~0u until allocated
};

// Symbols and header lines of one 32-byte chunk (scalar walk over its lines; used for the chunks
// the fast path does not take).
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

__global__ void __launch_bounds__(kParseThreads, GTS_PARSE_CTAS) k_parse(const __grid_constant__ Job job) {
    __shared__ ParseShared sh;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
#ifdef GTS_PHASE_TIMING
    u64 tprev = gtime();
#endif
    {
        uint8_t c = 0;
        switch (tid) {
            case 'A': case 'a': c = 1; break;
            case 'C': case 'c': c = 2; break;
            case 'T': case 't': case 'U': case 'u': c = 3; break;
            case 'G': case 'g': c = 4; break;
        }
        sh.enc[tid] = c;
        if (tid == 0) sh.ev_base = ~0u;
    }
    const u32 t = blockIdx.x;
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
    // the 512 bytes before the warp's first byte, for the line the warp starts in
    uint4 vb = make_uint4(0, 0, 0, 0);
    const i64 back_start = (i64)warp_base - 16 * (lane + 1);
    const bool back_ok = warp_base < job.n && back_start >= 0;
    if (back_ok) vb = __ldg(reinterpret_cast<const uint4 *>(job.in + back_start));

    uint8_t *raw = sh.raw + tid * kParseBytesPerThread;
    reinterpret_cast<uint4 *>(raw)[0] = v0;
    reinterpret_cast<uint4 *>(raw)[1] = v1;
    u32 nlmask = newline_mask16(v0) | newline_mask16(v1) << 16;
    if (len < 32) nlmask &= len ? ((1u << len) - 1u) : 0u;

    // ---- start of the line each chunk begins in, and the class of that line ----
    // before the warp's first byte: backward search for the previous '\n' (first in the 512 bytes
    // already loaded, then further back); `warp_first` is the first byte of that line
    u64 warp_ls = 0;
    uint8_t warp_first = 0;
    if (warp_base < job.n && warp_base > 0) {
        u32 m16 = back_ok ? newline_mask16(vb) : 0u;
        u64 hi = warp_base;
        uint4 vv = vb;
        while (true) {
            const unsigned b = __ballot_sync(0xFFFFFFFFu, m16 != 0);
            if (b) {
                const int l = __ffs(b) - 1;
                const u32 mm = __shfl_sync(0xFFFFFFFFu, m16, l);
                const int o = 31 - __clz(mm);  // offset of the newline in lane l's 16 bytes
                warp_ls = hi - 16ull * (l + 1) + o + 1;
                // the byte after it: in lane l's vector, or the first byte of lane l-1's
                const int ol = o + 1;
                const int src = ol < 16 ? l : l - 1;
                const int oo = ol & 15;
                const u32 wsel = (oo >> 2) == 0 ? vv.x : (oo >> 2) == 1 ? vv.y : (oo >> 2) == 2 ? vv.z : vv.w;
                const u32 byte = (wsel >> (8 * (oo & 3))) & 0xFFu;
                if (src >= 0) warp_first = (uint8_t)__shfl_sync(0xFFFFFFFFu, byte, src);
                break;  // src < 0: the line starts at hi (== warp_base in the first round): not used
            }
            if (hi <= 512) {  // the line starts at the beginning of the input
                warp_ls = 0;
                warp_first = job.in[0];
                break;
            }
            hi -= 512;
            const i64 st2 = (i64)hi - 16 * (lane + 1);
            vv = make_uint4(0, 0, 0, 0);
            m16 = 0;
            if (st2 >= 0) {
                vv = __ldg(reinterpret_cast<const uint4 *>(job.in + st2));
                m16 = newline_mask16(vv);
            }
        }
        // a newline found in a later round sits below hi: its successor may be the first byte of the
        // previous round's window, which the shuffle above cannot see; reload it in that rare case
        if (warp_ls != 0 && warp_ls != warp_base && (warp_ls & 511u) == 0) warp_first = job.in[warp_ls];
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
    __syncwarp();  // the warp's raw bytes are in shared memory
    const u64 line_start = excl_max ? warp_base + excl_max : warp_ls;
    int state = LS;
    if (len > 0 && line_start != pos0) {
        // the line's first byte: inside the warp's own kilobyte, or found by the backward search
        const uint8_t first = line_start >= warp_base ? sh.raw[warp * 1024 + (u32)(line_start - warp_base)] : warp_first;
        state = first == '>' ? HDR : SEQ;
    }
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
            slow = nlmask != 0;  // the header line ends here
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
    const bool track_nl = job.n >= job.max_line;  // only then can a line reach the reference's limit
    if (track_nl) {
        u32 fn = nlmask ? (u32)(tid * 32 + __ffs(nlmask) - 1) : 0xFFFFu;
        u32 ln = nlmask ? (u32)(tid * 32 + (31 - __clz(nlmask)) + 1) : 0u;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            fn = min(fn, __shfl_xor_sync(0xFFFFFFFFu, fn, o));
            ln = max(ln, __shfl_xor_sync(0xFFFFFFFFu, ln, o));
        }
        if (lane == 0) sh.warp_nl[warp] = fn << 16 | ln;
    }
    __syncthreads();
    u32 excl = inc - packed;
    u32 tile_total = 0;
#pragma unroll
    for (int w = 0; w < kParseThreads / 32; w++) {
        const u32 s = sh.warp_sum[w];
        if (w < warp) excl += s;
        tile_total += s;
    }
    const bool last_tile = (t + 1 == job.ntiles);
    const u32 S_own = tile_total & 0xFFFFu, R_tile = tile_total >> 16;
    const u32 S_tile = S_own + (last_tile && !job.no_end_pads ? 2u : 0u);  // the last tile owns the two end pads
    const u32 d = excl & 0xFFFFu;                      // slot index of the chunk's first symbol
    if (tid == 0 && R_tile) {
        // one slot range of the event list for the tile's header lines
        const u32 base = atomicAdd(&job.counters[2], R_tile);
        sh.ev_base = base;
    }
    PHASE_MARK(job, 2, tprev);  // barrier + tile scan

    // ---- fast chunks: encode, drop the newline / CR bytes, store whole words ----
    {
        // fast chunks (>= 30 symbols) link up: a chunk's last, partial word is completed with the
        // first symbols of the next lane's chunk and stored whole
        const unsigned bigm = __ballot_sync(0xFFFFFFFFu, fast);
        u32 C[9];
        C[8] = 0;
        if (fast) {
            C[0] = encode4(v0.x); C[1] = encode4(v0.y); C[2] = encode4(v0.z); C[3] = encode4(v0.w);
            C[4] = encode4(v1.x); C[5] = encode4(v1.y); C[6] = encode4(v1.z); C[7] = encode4(v1.w);
            if (job.warn) {
                // invalid nucleotides (encodedSequence.go:36-40): code 0, not N/n, not a dropped byte
                const u32 zc = gather_flags16(zero_code_flags(C[0]), zero_code_flags(C[1]), zero_code_flags(C[2]), zero_code_flags(C[3])) |
                               gather_flags16(zero_code_flags(C[4]), zero_code_flags(C[5]), zero_code_flags(C[6]), zero_code_flags(C[7])) << 16;
                const u32 nm = gather_flags16(n_flags(v0.x), n_flags(v0.y), n_flags(v0.z), n_flags(v0.w)) |
                               gather_flags16(n_flags(v1.x), n_flags(v1.y), n_flags(v1.z), n_flags(v1.w)) << 16;
                u32 inv = zc & ~nm & ~removed;
                while (inv) {
                    const int i = __ffs(inv) - 1;
                    inv &= inv - 1;
                    const u32 spos = d + (u32)i - (u32)__popc(removed & ((1u << i) - 1u));
                    const u32 slot = atomicAdd(&job.counters[4], 1u);
                    if (slot < job.warn_cap) job.warn[slot] = (u64)t << 32 | (u64)spos << 8 | raw[i];
                }
            }
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
            if (link_next) {      // append the next chunk's first symbols after symbol k-1
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
    PHASE_MARK(job, 3, tprev);  // fast chunks

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
            if (c_first && lane == 0) {  // pads of slot 0 (the bytes before the first header line)
                sh.sym[0] = job.has_carry ? (uint8_t)((job.carry >> 4) & 0xFu) : 0;
                sh.sym[1] = job.has_carry ? (uint8_t)(job.carry & 0xFu) : 0;
            }
            if (keep) {
                const uint8_t code = sh.enc[b];
                sh.sym[c_d + off] = code;
                if (job.warn && code == 0 && (b & 0xDF) != 'N') {
                    const u32 slot = atomicAdd(&job.counters[4], 1u);
                    if (slot < job.warn_cap) job.warn[slot] = (u64)t << 32 | (u64)(c_d + off) << 8 | b;
                }
            }
            if (hstart) {
                sh.sym[c_d + off] = 0;
                sh.sym[c_d + off + 1] = 0;
                const u32 k = c_rec + (u32)__popc(hs & lt);  // index among the tile's header lines
                u32 base;
                do {
                    base = sh.ev_base;
                } while (base == ~0u);
                if ((u64)base + k < job.rec_cap) {
                    HeaderEvent ev;
                    ev.hdr_off = c_pos0 + lane;
                    ev.tile = t;
                    ev.k = (uint16_t)k;
                    ev.pos = (uint16_t)(c_d + off + 2);
                    job.events[base + k] = ev;
                }
            }
        }
    }
    if (tid < 32) sh.sym[S_own + tid] = 0;  // end pads (last tile) and zero fill of the last word
    PHASE_MARK(job, 4, tprev);  // slow chunks, hints
    __syncthreads();  // all symbols staged
    PHASE_MARK(job, 5, tprev);  // barrier

    // ---- pack two symbols per byte and write the slot ----
    const u32 nw = (S_tile + 7u) >> 3;
    u32 *gw = job.sym + (u64)t * kSlotWords;
#pragma unroll
    for (int i = 0; i < (kSlotWords + kParseThreads - 1) / kParseThreads; i++) {
        const u32 w = tid + i * kParseThreads;
        if (w < nw) {
            const uint2 s2 = *reinterpret_cast<const uint2 *>(sh.sym + 8 * w);
            gw[w] = s2.x | (s2.y << 4);
        }
    }
    if (tid == 0) {
        TileInfo ti;
        ti.nsym = (uint16_t)S_tile;
        ti.nhdr = (uint16_t)R_tile;
        const u32 l1 = S_tile >= 1 ? sh.sym[S_tile - 1] : 0u, l2 = S_tile >= 2 ? sh.sym[S_tile - 2] : 0u;
        ti.tail = (uint8_t)(l1 | l2 << 4);
        ti.pad[0] = ti.pad[1] = ti.pad[2] = 0;
        ti.has_nl = 0; ti.first_nl = 0; ti.last_nl = 0;
        if (track_nl) {
            u32 fn = 0xFFFFu, ln = 0;
            for (int w = 0; w < kParseThreads / 32; w++) {
                fn = min(fn, sh.warp_nl[w] >> 16);
                ln = max(ln, sh.warp_nl[w] & 0xFFFFu);
            }
            if (ln) {
                ti.has_nl = 1;
                ti.first_nl = (uint16_t)fn;
                ti.last_nl = (uint16_t)(ln - 1);
            }
        }
        job.tile_info[t] = ti;
    }
    PHASE_MARK(job, 6, tprev);  // pack
}

cudaError_t launch_parse(const Job &job, cudaStream_t stream) {
    k_parse<<<job.ntiles, kParseThreads, 0, stream>>>(job);
    return cudaGetLastError();
}

}  // namespace gts
