// gts_parse.cu -- k_parse: raw FASTA bytes -> per-tile symbol slots, header events, tile counts.
//
// Reference behaviour reproduced (feliixx/gotranseq):
//   readSequenceFromFasta  transeq/transeq.go:183-216 with bufio.ScanLines semantics (lines end at
//                          '\n', one trailing '\r' is dropped, empty lines are skipped, a line whose
//                          first byte is '>' starts a record)
//   newEncodedSequence     transeq/encodedSequence.go:17-43 (A/a 1, C/c 2, T/t/U/u 3, G/g 4, else 0;
//                          anything that is not ACGTUN is also reported: the WARNING line of :39)
// tests/gpu_model.py parse() is the executable specification of the line rules used here.
//
// One WARP per tile of 8 KB: it walks the tile's eight rows of 1 KB (32 lanes x 32 bytes) in order
// and carries the line state (at a line start / inside a header line / inside a sequence line) and
// the running symbol count from row to row in registers.  Nothing is exchanged between warps: no
// CTA barrier, no scan across warps, and the backward search for the line a tile starts in runs
// once per tile.  Every tile is independent of every other tile: its symbols are compacted inside
// its own slot, the header lines it finds are appended to an unordered event list, and the stream
// offsets are assigned afterwards by k_tile_scan / k_records.
#include "gts_common.cuh"
#include "gts_kernels.h"

#ifndef GTS_PARSE_CTAS
#define GTS_PARSE_CTAS 32
#endif

namespace gts {

enum { LS = 0, SEQ = 1, HDR = 2 };

constexpr int kParseWarps = 1;                          // one warp per CTA: the tile index is uniform for the compiler
constexpr int kRowBytes = 32 * kParseBytesPerThread;    // 1 KB per row
constexpr int kRows = kParseTile / kRowBytes;           // 8 rows per tile
constexpr int kSymBuf = 16 + kRowBytes + 64 + 48;       // carried symbols + a row's symbols + zero tail
static_assert(kSymBuf % 16 == 0, "row buffers stay 16-byte aligned");

struct WarpShared {
    alignas(16) uint8_t raw[kRowBytes];  // the row's raw bytes
    alignas(16) uint8_t sym[kSymBuf];    // symbols not yet packed, one per byte: [0, carry) from the rows before
};
struct ParseShared {
    WarpShared w[kParseWarps];
    alignas(4) uint8_t enc[256];
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

// The line the byte at `base` (0 < base < n, a multiple of 16) lies in: is `base` its first byte,
// and if not, is it a header line?  Backward search for the previous '\n' in windows of 512 bytes,
// all 32 lanes; the results are uniform over the warp.
__device__ __forceinline__ void line_state_at(const Job &job, u64 base, int lane, bool &at_start, bool &in_hdr) {
    at_start = false;
    in_hdr = false;
    u64 hi = base;
    while (true) {
        const i64 st = (i64)hi - 16 * (lane + 1);
        u32 m16 = 0;
        if (st >= 0) m16 = newline_mask16(__ldg(reinterpret_cast<const uint4 *>(job.in + st)));
        const unsigned b = __ballot_sync(0xFFFFFFFFu, m16 != 0);
        if (b) {
            const int l = __ffs(b) - 1;  // the lane holding the closest newline
            const u32 mm = __shfl_sync(0xFFFFFFFFu, m16, l);
            const int o = 31 - __clz(mm);  // its offset in that lane's 16 bytes
            const u64 ls = hi - 16ull * (l + 1) + o + 1;  // first byte of the line
            // (uniform address: one load, broadcast; the votes make the results uniform for the compiler too)
            at_start = __any_sync(0xFFFFFFFFu, ls == base);
            in_hdr = __any_sync(0xFFFFFFFFu, ls != base && job.in[ls] == '>');
            return;
        }
        if (hi <= 512) {  // the line starts at the beginning of the input
            in_hdr = __any_sync(0xFFFFFFFFu, job.in[0] == '>');
            return;
        }
        hi -= 512;
    }
}

__global__ void __launch_bounds__(kParseWarps * 32, GTS_PARSE_CTAS) k_parse(const __grid_constant__ Job job) {
    __shared__ ParseShared sh;
    const int tid = threadIdx.x;
    const int lane = tid & 31, warp = tid >> 5;
    {
        // code table of the slow path (encodedSequence.go:25-41), 64 words: only the words holding
        // "@ABC" / "DEFG" / "TUVW" and their lower-case twins are not zero
        static_assert(kParseWarps == 1, "one warp writes the table");
        const int w = lane & 7;  // words 16..31 hold the letters: 16 + w upper case, 24 + w lower case
        const u32 v = w == 0 ? 0x02000100u : w == 1 ? 0x04000000u : w == 5 ? 0x00000303u : 0u;
        u32 *e = reinterpret_cast<u32 *>(sh.enc);
        e[lane] = lane >= 16 ? v : 0u;
        e[32 + lane] = 0u;
    }
    __syncthreads();  // the only CTA-wide barrier: the table of the slow path
    pdl_launch();     // k_tile_scan's CTAs may be scheduled while the parse drains
    const u32 t = blockIdx.x * kParseWarps + warp;
    if (t >= job.ntiles) return;
    WarpShared &ws = sh.w[warp];
    const u64 n = job.n;
    const u64 tile_base = (u64)t * kParseTile;
    const unsigned lt = (1u << lane) - 1u;

    // ---- the line the tile starts in ----
    bool row_ls = true, row_in_hdr = false;  // uniform: the row's first byte starts a line / continues a header line
    if (tile_base > 0 && tile_base < n) line_state_at(job, tile_base, lane, row_ls, row_in_hdr);

    // ---- running state of the tile ----
    u32 S = 0, R = 0;        // symbols / header lines of the rows done
    u32 carry = 0;           // symbols at the front of ws.sym that are not packed yet
    u32 done_w = 0;          // slot words written
    u32 tail_last = 0, tail_prev = 0;  // (lane 0) the slot's last two symbols so far
    u32 fn = 0xFFFFu, ln = 0;          // first newline / last newline + 1 of the tile (over-long line rule)
    const bool track_nl = n >= job.max_line;  // only then can a line reach the reference's limit
    u32 *gw = job.sym + (u64)t * kSlotWords;
    if (lane < 4) reinterpret_cast<u32 *>(ws.sym)[lane] = 0;
    __syncwarp();
    if (t == 0) {  // pads of slot 0 (the bytes before the first header line), or the two symbols in front of a piece
        if (lane == 0) {
            tail_prev = job.pv.has_carry ? ((job.pv.carry >> 4) & 0xFu) : 0u;
            tail_last = job.pv.has_carry ? (job.pv.carry & 0xFu) : 0u;
            ws.sym[0] = (uint8_t)tail_prev;
            ws.sym[1] = (uint8_t)tail_last;
        }
        carry = 2;
        S = 2;
    }
    __syncwarp();

    // ---- rows: the next row's bytes are requested while the current one is worked on ----
    uint4 nx0 = make_uint4(0, 0, 0, 0), nx1 = make_uint4(0, 0, 0, 0);
    auto load_row = [&](u64 row_base) {
        const u64 p0 = row_base + (u64)lane * kParseBytesPerThread;
        nx0 = make_uint4(0, 0, 0, 0);
        nx1 = make_uint4(0, 0, 0, 0);
        if (p0 + 32 <= n) {
            const uint4 *p = reinterpret_cast<const uint4 *>(job.in + p0);
            nx0 = __ldg(p);
            nx1 = __ldg(p + 1);
        } else if (p0 < n) {
            u32 w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const int l = (int)(n - p0);
            for (int i = 0; i < l; i++) w[i >> 2] |= (u32)job.in[p0 + i] << (8 * (i & 3));
            nx0 = make_uint4(w[0], w[1], w[2], w[3]);
            nx1 = make_uint4(w[4], w[5], w[6], w[7]);
        }
    };
    load_row(tile_base);

    const u32 tile_bytes = n - tile_base < (u64)kParseTile ? (u32)(n - tile_base) : (u32)kParseTile;  // (0 for an empty input)
#pragma unroll 1
    for (int r = 0; r < kRows; r++) {
        if ((u32)r * kRowBytes >= tile_bytes) break;  // (an empty input has no row at all)
        const u64 row_base = tile_base + (u64)r * kRowBytes;
        const uint4 v0 = nx0, v1 = nx1;
        const bool full_row = (u32)(r + 1) * kRowBytes <= tile_bytes;
        const bool has_next_row = r + 1 < kRows && (u32)(r + 1) * kRowBytes < tile_bytes;
        if (has_next_row) load_row(row_base + kRowBytes);

        // ---- newlines ----
        const u32 f0 = newline_flags(v0.x), f1 = newline_flags(v0.y), f2 = newline_flags(v0.z), f3 = newline_flags(v0.w);
        const u32 f4 = newline_flags(v1.x), f5 = newline_flags(v1.y), f6 = newline_flags(v1.z), f7 = newline_flags(v1.w);
        u32 nlmask = gather_flags16(f0, f1, f2, f3) | gather_flags16(f4, f5, f6, f7) << 16;

        // ---- plain rows: 1 KB of sequence lines and nothing else ----
        // Every byte is a letter (bit 6 set: no '>', no CR, no blank, no digit) or a '\n', no chunk
        // holds more than two newlines, and the row does not start inside a header line: then every
        // chunk is a "fast" chunk, there is no header line in the row and no state to work out.
        const u32 letters = (v0.x | (f0 >> 1)) & (v0.y | (f1 >> 1)) & (v0.z | (f2 >> 1)) & (v0.w | (f3 >> 1)) &
                            (v1.x | (f4 >> 1)) & (v1.y | (f5 >> 1)) & (v1.z | (f6 >> 1)) & (v1.w | (f7 >> 1));
        const bool plain_chunk = (letters & 0x40404040u) == 0x40404040u && __popc(nlmask) <= 2;
        const bool plain_row = full_row && (row_ls || !row_in_hdr) && __all_sync(0xFFFFFFFFu, plain_chunk);

        u32 removed = 0, nsym = 0, nhdr = 0, excl_syms = 0, excl_hdrs = 0, row_syms = 0, row_hdrs = 0;
        bool slow = false, fast = false, cr_end = false, next_ls, next_in_hdr;
        int state = SEQ, len = 32;
        unsigned NL;
        uint8_t *raw = ws.raw + lane * kParseBytesPerThread;
        const u64 pos0 = row_base + (u64)lane * kParseBytesPerThread;
        if (plain_row) {
            removed = nlmask;
            fast = true;
            const u32 k = (u32)__popc(nlmask);  // 0, 1 or 2 newlines: the prefix sums come from two ballots
            nsym = 32u - k;
            const unsigned c1 = __ballot_sync(0xFFFFFFFFu, k >= 1u), c2 = __ballot_sync(0xFFFFFFFFu, k == 2u);
            NL = c1;
            excl_syms = 32u * (u32)lane - (u32)__popc(c1 & lt) - (u32)__popc(c2 & lt);
            row_syms = 32u * 32u - (u32)__popc(c1) - (u32)__popc(c2);
            next_ls = __any_sync(0xFFFFFFFFu, lane == 31 && (nlmask >> 31) != 0);
            next_in_hdr = false;
        } else {
        // first byte behind the row: the next row's first byte (already requested), or the next tile's
        u32 after_row = 0;
        if (lane == 0) {
            if (has_next_row) after_row = nx0.x & 0xFFu;
            else if (row_base + kRowBytes < n) after_row = job.in[row_base + kRowBytes];
        }
        len = pos0 >= n ? 0 : (n - pos0 < 32 ? (int)(n - pos0) : 32);
        reinterpret_cast<uint4 *>(raw)[0] = v0;
        reinterpret_cast<uint4 *>(raw)[1] = v1;
        if (len < 32) nlmask &= len ? ((1u << len) - 1u) : 0u;
        const u32 b0 = v0.x & 0xFFu;

        // ---- the state each chunk starts in ----
        NL = __ballot_sync(0xFFFFFFFFu, nlmask != 0);
        const unsigned LNL = __ballot_sync(0xFFFFFFFFu, (nlmask >> 31) != 0);  // the chunk's last byte is '\n'
        const unsigned G = __ballot_sync(0xFFFFFFFFu, len > 0 && b0 == '>');
        u32 next_first = __shfl_down_sync(0xFFFFFFFFu, b0, 1);
        const u32 ar = __shfl_sync(0xFFFFFFFFu, after_row, 0);
        if (lane == 31) next_first = ar;
        // is the line that starts behind the chunk's last newline a header line?
        bool tail_hdr = false;
        if (nlmask) {
            const int e = 31 - __clz(nlmask);
            if (e < 31) tail_hdr = e + 1 < len && raw[e + 1] == '>';
            else tail_hdr = pos0 + 32 < n && next_first == '>';
        }
        const unsigned TH = __ballot_sync(0xFFFFFFFFu, tail_hdr);
        const bool row_line_hdr = row_ls ? (G & 1u) != 0 : row_in_hdr;  // the line the row's first byte lies in
        const unsigned prev_nl = NL & lt;
        const bool chunk_ls = lane > 0 ? ((LNL >> (lane - 1)) & 1u) != 0 : row_ls;
        const bool hdr_line = prev_nl ? ((TH >> (31 - __clz(prev_nl))) & 1u) != 0 : row_line_hdr;
        state = chunk_ls ? LS : (hdr_line ? HDR : SEQ);
        // ... and the next row's
        next_ls = (LNL >> 31) != 0;
        next_in_hdr = NL ? ((TH >> (31 - __clz(NL))) & 1u) != 0 : row_line_hdr;

        const bool last_cr = len == 32 ? (v1.w >> 24) == '\r' : (len > 0 && raw[len - 1] == '\r');
        cr_end = last_cr && ((pos0 + len >= n) || next_first == '\n');

        // ---- classify the chunk ----
        //   fast:  32 bytes of sequence lines only, at most two bytes to drop (newline, CR)
        //   slow:  anything with a header line start / end in it, or short (end of input), ...
        //   none:  no symbols (inside a header line, or past the end of the input)
        if (len > 0) {
            if (state == HDR) {
                slow = nlmask != 0;  // the header line ends here
            } else {
                slow = (state == LS && b0 == '>') || len < 32;
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
        if (fast) nsym = 32u - (u32)__popc(removed);
        else if (slow) count_chunk(raw, len, nlmask, state, cr_end, nsym, nhdr);

        // ---- positions: scan over the row ----
        const u32 packed = nsym | nhdr << 16;
        u32 inc = packed;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u32 y = __shfl_up_sync(0xFFFFFFFFu, inc, o);
            if (lane >= o) inc += y;
        }
        const u32 row_total = __shfl_sync(0xFFFFFFFFu, inc, 31);
        row_syms = row_total & 0xFFFFu;
        row_hdrs = row_total >> 16;
        excl_syms = (inc - packed) & 0xFFFFu;
        excl_hdrs = (inc - packed) >> 16;
        }  // generic row
        const u32 d = carry + excl_syms;  // buffer index of the chunk's first symbol
        const u32 slot0 = S - carry;      // slot index of buffer index 0
        u32 ev_base = 0;
        if (row_hdrs) {  // one range of the event list for the row's header lines
            if (lane == 0) ev_base = atomicAdd(&job.counters[2], row_hdrs);
            ev_base = __shfl_sync(0xFFFFFFFFu, ev_base, 0);
        }
        if (track_nl && NL) {
            const int l1 = 31 - __clz(NL);
            const u32 m1 = __shfl_sync(0xFFFFFFFFu, nlmask, l1);
            ln = (u32)(r * kRowBytes + l1 * 32 + (31 - __clz(m1)) + 1);  // rows are walked upwards
            if (fn == 0xFFFFu) {
                const int l0 = __ffs(NL) - 1;
                const u32 m0 = __shfl_sync(0xFFFFFFFFu, nlmask, l0);
                fn = (u32)(r * kRowBytes + l0 * 32 + __ffs(m0) - 1);
            }
        }

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
                        const u32 spos = slot0 + d + (u32)i - (u32)__popc(removed & ((1u << i) - 1u));
                        const u32 slot = atomicAdd(&job.counters[4], 1u);
                        const u32 wsel = i < 16 ? (i < 8 ? (i < 4 ? v0.x : v0.y) : (i < 12 ? v0.z : v0.w))
                                                : (i < 24 ? (i < 20 ? v1.x : v1.y) : (i < 28 ? v1.z : v1.w));
                        if (slot < job.warn_cap) job.warn[slot] = (u64)t << 32 | (u64)spos << 8 | ((wsel >> (8 * (i & 3))) & 0xFFu);
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
                u32 *dstw = reinterpret_cast<u32 *>(ws.sym + (d - a));
                uint8_t *dstb = ws.sym + (d - a);
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

        // ---- slow chunks: the 32 lanes of the warp take one byte each ----
        {
            unsigned slowm = __ballot_sync(0xFFFFFFFFu, slow);
            if (slowm) __syncwarp();  // the row's raw bytes are read across lanes
            while (slowm) {
                const int c = __ffs(slowm) - 1;
                slowm &= slowm - 1;
                const int c_len = __shfl_sync(0xFFFFFFFFu, len, c);
                const int c_state = __shfl_sync(0xFFFFFFFFu, state, c);
                const u32 c_d = __shfl_sync(0xFFFFFFFFu, d, c);
                const u32 c_rec = R + __shfl_sync(0xFFFFFFFFu, excl_hdrs, c);
                const bool c_cr_end = __shfl_sync(0xFFFFFFFFu, (int)cr_end, c) != 0;
                const u64 c_pos0 = row_base + (u64)c * 32u;
                const bool valid = lane < c_len;
                const uint8_t b = valid ? ws.raw[c * 32 + lane] : 0;
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
                const u32 off = (u32)__popc(keepm & lt) + 2u * (u32)__popc(hs & lt);
                if (keep) {
                    const uint8_t code = sh.enc[b];
                    ws.sym[c_d + off] = code;
                    if (job.warn && code == 0 && (b & 0xDF) != 'N') {
                        const u32 slot = atomicAdd(&job.counters[4], 1u);
                        if (slot < job.warn_cap) job.warn[slot] = (u64)t << 32 | (u64)(slot0 + c_d + off) << 8 | b;
                    }
                }
                if (hstart) {
                    ws.sym[c_d + off] = 0;
                    ws.sym[c_d + off + 1] = 0;
                    const u32 k = c_rec + (u32)__popc(hs & lt);  // index among the tile's header lines
                    const u32 e = ev_base + (k - R);
                    if ((u64)e < job.rec_cap) {
                        HeaderEvent ev;
                        ev.hdr_off = c_pos0 + lane;
                        ev.tile = t;
                        ev.k = (uint16_t)k;
                        ev.pos = (uint16_t)(slot0 + c_d + off + 2);
                        job.events[e] = ev;
                    }
                }
            }
        }
        const u32 end = carry + row_syms;                    // buffer index behind the row's last symbol
        ws.sym[end + lane] = 0;                              // zero tail: the last word's fill, the end pads
        if (lane < 16) ws.sym[end + 32 + lane] = 0;
        __syncwarp();  // the row's symbols are staged

        // ---- pack two symbols per byte and write the slot words that are complete ----
        const u32 nw = end >> 3;  // <= (7 + 32 * 32 + 2) / 8 = 129
        {
            u32 *gp = gw + done_w + lane;
            const uint2 *sp = reinterpret_cast<const uint2 *>(ws.sym) + lane;
#pragma unroll
            for (int i = 0; i < 5; i++) {
                if ((u32)lane + 32u * i < nw) {
                    const uint2 s2 = sp[32 * i];
                    gp[32 * i] = s2.x | (s2.y << 4);
                }
            }
        }
        if (lane == 0) {
            if (row_syms >= 2) {
                tail_last = ws.sym[end - 1];
                tail_prev = ws.sym[end - 2];
            } else if (row_syms == 1) {
                tail_prev = tail_last;
                tail_last = ws.sym[end - 1];
            }
        }
        // the symbols behind the last complete word move to the front (with zeros behind them)
        u32 keepw = 0;
        if (lane < 4) keepw = reinterpret_cast<const u32 *>(ws.sym + 8 * nw)[lane];
        __syncwarp();
        if (lane < 4) reinterpret_cast<u32 *>(ws.sym)[lane] = keepw;
        __syncwarp();
        carry = end & 7u;
        done_w += nw;
        S += row_syms;
        R += row_hdrs;
        row_ls = next_ls;
        row_in_hdr = next_in_hdr;
    }

    // ---- the tile's last word(s), its counts ----
    const bool last_tile = (t + 1 == job.ntiles);
    const u32 S_tile = S + (last_tile && !job.pv.no_end_pads ? 2u : 0u);  // the last tile owns the two end pads
    const u32 nw_total = (S_tile + 7u) >> 3;
    if (done_w + (u32)lane < nw_total) {  // at most two words: up to 7 carried symbols + 2 pads, zero filled
        const uint2 s2 = *reinterpret_cast<const uint2 *>(ws.sym + 8 * lane);
        gw[done_w + lane] = s2.x | (s2.y << 4);
    }
    if (lane == 0) {
        TileInfo ti;
        ti.nsym = (uint16_t)S_tile;
        ti.nhdr = (uint16_t)R;
        u32 l1 = tail_last, l2 = tail_prev;
        if (S_tile != S) { l1 = 0; l2 = 0; }  // the end pads are the slot's last two symbols
        if (S_tile < 2) l2 = 0;
        if (S_tile < 1) l1 = 0;
        ti.tail = (uint8_t)(l1 | l2 << 4);
        ti.pad[0] = ti.pad[1] = ti.pad[2] = 0;
        ti.has_nl = 0; ti.first_nl = 0; ti.last_nl = 0;
        if (track_nl && ln) {
            ti.has_nl = 1;
            ti.first_nl = (uint16_t)fn;
            ti.last_nl = (uint16_t)(ln - 1);
        }
        job.tile_info[t] = ti;
    }
}

cudaError_t launch_parse(const Job &job, cudaStream_t stream) {
    k_parse<<<(job.ntiles + kParseWarps - 1) / kParseWarps, kParseWarps * 32, 0, stream>>>(job);
    return cudaGetLastError();
}

}  // namespace gts
