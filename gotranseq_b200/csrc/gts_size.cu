// gts_size.cu -- k_tile_scan, k_records, k_size: from the per-tile results of k_parse to the record
// table with output offsets.
//
// Reference behaviour reproduced (feliixx/gotranseq):
//   frame lengths / tail rules   transeq/writer.go:85-116 (Go's % truncates)
//   trim                         transeq/writer.go:141-147,159-169
//   header line = bytes up to the line's '\n', one trailing '\r' dropped (bufio.ScanLines)
#include "gts_common.cuh"
#include "gts_kernels.h"

namespace gts {

// Over-long line rule (transeq.go:27,186: bufio.Scanner gives up at a line of >= 100 MiB and the
// reference silently stops reading there).  Newline summaries are folded left to right; a line
// starts after the previous newline (or at 0) and its length excludes its own newline.
__device__ __forceinline__ void nl_fold(NlPart &acc, const NlPart &p, u64 max_line) {
    if (acc.cut < 0) {
        if (acc.last >= 0 && p.first >= 0 && (u64)(p.first - acc.last - 1) >= max_line) acc.cut = acc.last + 1;
        else if (p.cut >= 0) acc.cut = p.cut;
    }
    if (acc.first < 0) acc.first = p.first;
    if (p.last >= 0) acc.last = p.last;
}

// =====================================================================================
// k_tile_scan: exclusive scan of the parse tiles' symbol / header counts
// =====================================================================================
__global__ void __launch_bounds__(kScanThreads) k_tile_scan(const __grid_constant__ Job job) {
    __shared__ u64 warp_s[kScanThreads / 32];
    __shared__ u64 warp_r[kScanThreads / 32];
    __shared__ u64 s_excl_s, s_excl_r;
    __shared__ u32 s_tile;
    __shared__ NlPart s_nl[kScanThreads];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const u32 ntiles = job.ntiles;
    const u32 nscan = (ntiles + kScanTile - 1) / kScanTile;
    pdl_launch();
    pdl_wait();  // k_parse's tile records
    const PieceParams &PP = piece_of(job);
    while (true) {
        __syncthreads();
        if (tid == 0) s_tile = atomicAdd(&job.counters[0], 1u);
        __syncthreads();
        const u32 st = s_tile;
        if (st >= nscan) return;
        const u32 t0 = st * kScanTile + tid * kScanPerThread;
        u32 ns[kScanPerThread], nh[kScanPerThread], tl[kScanPerThread];
        u64 sum_s = 0, sum_r = 0;
        const bool track_nl = job.n >= job.max_line;
        NlPart mine;
        mine.first = mine.last = mine.cut = -1;
        mine.pad = 0;
#pragma unroll
        for (int i = 0; i < kScanPerThread; i++) {
            ns[i] = nh[i] = tl[i] = 0;
            if (t0 + i < ntiles) {
                const TileInfo ti = job.tile_info[t0 + i];
                ns[i] = ti.nsym;
                nh[i] = ti.nhdr;
                tl[i] = ti.tail;
                if (track_nl && ti.has_nl) {
                    // inside one tile no line can reach the limit (it is at least two tiles long)
                    NlPart p;
                    p.first = (i64)(t0 + i) * kParseTile + ti.first_nl;
                    p.last = (i64)(t0 + i) * kParseTile + ti.last_nl;
                    p.cut = -1;
                    p.pad = 0;
                    nl_fold(mine, p, job.max_line);
                }
            }
            sum_s += ns[i];
            sum_r += nh[i];
        }
        if (track_nl) {
            // fold the lanes' summaries in lane order (5 shuffle rounds are not enough for a
            // non-commutative fold of 3 fields, so lane 0 walks its warp: 32 short steps)
            s_nl[tid] = mine;
            __syncwarp();
            if (lane == 0) {
                NlPart acc = mine;
                for (int i = 1; i < 32; i++) nl_fold(acc, s_nl[tid + i], job.max_line);
                s_nl[tid] = acc;
            }
        }
        u64 inc_s = sum_s, inc_r = sum_r;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const u64 ys = __shfl_up_sync(0xFFFFFFFFu, inc_s, o);
            const u64 yr = __shfl_up_sync(0xFFFFFFFFu, inc_r, o);
            if (lane >= o) { inc_s += ys; inc_r += yr; }
        }
        if (lane == 31) { warp_s[warp] = inc_s; warp_r[warp] = inc_r; }
        __syncthreads();
        u64 ex_s = inc_s - sum_s, ex_r = inc_r - sum_r, tot_s = 0, tot_r = 0;
        for (int w = 0; w < kScanThreads / 32; w++) {
            if (w < warp) { ex_s += warp_s[w]; ex_r += warp_r[w]; }
            tot_s += warp_s[w];
            tot_r += warp_r[w];
        }
        if (track_nl && tid == kScanThreads - 1) {  // (after the barrier above) fold the threads' summaries
            NlPart acc;
            acc.first = acc.last = acc.cut = -1;
            acc.pad = 0;
            for (int w = 0; w < kScanThreads / 32; w++) nl_fold(acc, s_nl[32 * w], job.max_line);
            job.scan_nl[st] = acc;
        }
        if (warp == 0) {
            Pair agg = {tot_s, tot_r};
            Pair ex = scan_lookback(job.t_status, st, agg);
            if (lane == 0) {
                s_excl_s = ex.a;
                s_excl_r = ex.b;
                if (st + 1 == nscan) {
                    // totals; the last tile's two end pads are not counted as stream symbols
                    const u64 S2 = ex.a + tot_s, R = ex.b + tot_r;
                    const u64 S = S2 - (PP.no_end_pads ? 0 : 2);
                    job.totals->total_sym = S;
                    u32 tail = 0;
                    {   // the stream's last two symbols (a later piece of the same record needs them)
                        u32 got = 0;
                        for (i64 u = (i64)ntiles - 1; got < 2 && u >= 0; u--) {
                            const TileInfo ti = job.tile_info[u];
                            // the last tile's end pads are not part of the record
                            const u32 ns_u = ti.nsym - ((u32)u == ntiles - 1 && !PP.no_end_pads ? 2u : 0u);
                            for (u32 k = 0; k < ns_u && got < 2; k++) {
                                tail |= slot_symbol(job.sym, (u32)u, ns_u - 1 - k) << (4 * got);
                                got++;
                            }
                        }
                        job.totals->tail = tail;
                    }
                    if (job.info_out) {
                        // piece scan: the counts and boundary codes the other pieces need
                        PieceInfoDev inf;
                        inf.headers = R;
                        inf.nucleotides = S - 2 - 2 * R;
                        inf.tail = tail;
                        inf.pad = 0;
                        // the header line the buffer starts with (first piece of a split record): its
                        // length up to the '\n', one trailing CR dropped (bufio.ScanLines)
                        inf.header_len = ~0ull;
                        if (job.n > 0 && job.in[0] == '>') {
                            u64 p = 0;
                            while (p < job.n && job.in[p] != '\n') p++;
                            inf.header_len = p - (job.in[p - 1] == '\r' ? 1 : 0);
                        }
                        // codes of the buffer's first 192 nucleotides (the piece before it needs them);
                        // slot 0's pads and a header's pads are skipped
                        const u32 skip = 2u + 2u * (u32)R;
                        u32 seen = 0, got = 0;
                        for (u32 u = 0; u < ntiles && got < 192u; u++) {
                            const u32 ns_u = job.tile_info[u].nsym - (u == ntiles - 1 && !PP.no_end_pads ? 2u : 0u);
                            for (u32 k = 0; k < ns_u && got < 192u; k++, seen++)
                                if (seen >= skip) inf.head[got++] = (uint8_t)slot_symbol(job.sym, u, k);
                        }
                        for (; got < 192u; got++) inf.head[got] = 0;
                        for (int i = 0; i < 32; i++) inf.fill[i] = 0;
                        *job.info_out = inf;
                    }
                    job.totals->n_headers = R;
                    if (R + 2 <= job.rec_cap) {
                        // slot 0 (bytes before the first header line) and the sentinel after the last record
                        job.hdr_off[0] = 0;
                        job.dbase[0] = PP.t0_base + 2;
                        job.loc[0] = 2;
                        job.dbase[R + 1] = PP.t0_base + S + 2;
                        job.loc[R + 1] = (u64)(ntiles - 1) << 16 | job.tile_info[ntiles - 1].nsym;
                    } else {
                        atomicOr(&job.totals->flags, kFlagRecOverflow);
                    }
                }
            }
        }
        __syncthreads();
        u64 T = PP.t0_base + s_excl_s + ex_s, Rr = s_excl_r + ex_r;
        // the two symbols before the thread's first tile (windows reach back two symbols): from the
        // tiles before it; for its other tiles they follow from its own tiles' tails
        u32 carry = 0;
        if (t0 < ntiles) {
            u32 got = 0;
            for (i64 u = (i64)t0 - 1; got < 2 && u >= 0; u--) {
                const TileInfo ti = job.tile_info[u];
                if (ti.nsym >= 1) {
                    carry |= (u32)(ti.tail & 0xFu) << (4 * got);
                    got++;
                    if (got < 2 && ti.nsym >= 2) {
                        carry |= (u32)(ti.tail >> 4) << (4 * got);
                        got++;
                    }
                }
            }
        }
#pragma unroll
        for (int i = 0; i < kScanPerThread; i++) {
            const u32 t = t0 + i;
            if (t < ntiles) {
                TilePrefix tp;
                tp.T0 = T;
                tp.R0 = (u32)Rr;
                tp.carry = carry;
                job.tile_pre[t] = tp;
            }
            if (ns[i] >= 2) carry = tl[i];
            else if (ns[i] == 1) carry = (tl[i] & 0xFu) | (carry & 0xFu) << 4;
            T += ns[i];
            Rr += nh[i];
        }
    }
}

// =====================================================================================
// k_records: header events -> record table
// =====================================================================================
__global__ void __launch_bounds__(256) k_records(const __grid_constant__ Job job) {
    pdl_launch();
    pdl_wait();  // k_tile_scan's tile offsets and totals
    if (job.n >= job.max_line && blockIdx.x == 0 && threadIdx.x == 0) {
        // over-long line rule: fold the scan tiles' newline summaries, then the unterminated last line
        const u32 nscan = (job.ntiles + kScanTile - 1) / kScanTile;
        NlPart acc;
        acc.first = acc.last = acc.cut = -1;
        acc.pad = 0;
        i64 cut = -1;
        for (u32 i = 0; i < nscan && cut < 0; i++) {
            const NlPart p = job.scan_nl[i];
            // the first line of the input starts at 0 without a newline in front of it
            if (acc.last < 0 && p.first >= 0 && (u64)p.first >= job.max_line) cut = 0;
            nl_fold(acc, p, job.max_line);
            if (cut < 0) cut = acc.cut;
        }
        if (cut < 0 && job.n - (u64)(acc.last + 1) >= job.max_line) cut = acc.last + 1;
        if (cut >= 0) {
            job.totals->cut = (u64)cut;
            atomicOr(&job.totals->flags, kFlagLongLine);
        }
    }
    const u64 R = job.totals->n_headers;
    if (R + 2 > job.rec_cap) return;  // overflow: the host retries with a larger table
    // A group of kRecLanes lanes per header line: the record's table entry, and the header line measured
    // 16 * kRecLanes bytes per step (one aligned 16-byte load per lane, ballots inside the group for the
    // newline and the first space): its length up to the '\n' or the end of the input, one trailing CR
    // dropped (bufio.ScanLines), and the index of its first space (writer.go:121).  Small groups keep
    // many header lines in flight: the kernel is a chain of dependent loads (event -> tile offset ->
    // header bytes) and nothing else.
    constexpr int G = kRecLanes;
    const int lane = threadIdx.x & 31, sub = lane & (G - 1), gshift = lane & ~(G - 1);
    const unsigned gbits = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    const unsigned gmask = gbits << gshift;
    const u64 ngroups = ((u64)gridDim.x * blockDim.x) / G;
    for (u64 e = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / G; e < R; e += ngroups) {
        const HeaderEvent ev = job.events[e];
        const TilePrefix tp = job.tile_pre[ev.tile];
        const u64 slot = (u64)tp.R0 + ev.k + 1;
        const u64 off = ev.hdr_off;
        u64 p = job.n, sp = ~0ull;  // newline (or end of input), first space
        u32 cr = 0;                  // the byte in front of p is a CR
        u32 last_cr = 0;             // ... the last byte of the step before
        for (u64 base = off & ~15ull; base < job.n; base += 16u * G) {
            const u64 wa = base + 16u * (u32)sub;
            u32 mnl = 0, msp = 0, mcr = 0;  // bit i: byte wa + i is a '\n' / a space / a CR, bytes outside [off, n) masked away
            if (wa < job.n && wa + 16 > off) {
                // (the aligned 16 bytes hold at least one byte of the input: the load stays inside its allocation)
                const uint4 v = *reinterpret_cast<const uint4 *>(job.in + wa);
                u32 vm = job.n - wa < 16 ? (1u << (u32)(job.n - wa)) - 1u : 0xFFFFu;
                if (wa < off) vm &= ~((1u << (u32)(off - wa)) - 1u);
                mnl = newline_mask16(v) & vm;
                msp = gather_flags16(byte_eq_flags(v.x, 0x20202020u), byte_eq_flags(v.y, 0x20202020u),
                                     byte_eq_flags(v.z, 0x20202020u), byte_eq_flags(v.w, 0x20202020u)) & vm;
                mcr = gather_flags16(byte_eq_flags(v.x, 0x0D0D0D0Du), byte_eq_flags(v.y, 0x0D0D0D0Du),
                                     byte_eq_flags(v.z, 0x0D0D0D0Du), byte_eq_flags(v.w, 0x0D0D0D0Du)) & vm;
            }
            const unsigned bnl = (__ballot_sync(gmask, mnl != 0) >> gshift) & gbits;
            const unsigned bsp = (__ballot_sync(gmask, msp != 0) >> gshift) & gbits;
            // bit i: the byte in front of byte wa + i is a CR (the lane before, or the step before, supplies bit 0)
            u32 before = __shfl_up_sync(gmask, mcr, 1, G) >> 15;
            if (sub == 0) before = last_cr;
            const u32 crb = (mcr << 1) | (before & 1u);
            if (bsp && sp == ~0ull) {
                const int l = __ffs(bsp) - 1;
                const u32 f = __shfl_sync(gmask, msp, gshift + l);
                sp = base + 16u * l + (u32)(__ffs(f) - 1);
            }
            if (bnl) {
                const int l = __ffs(bnl) - 1;
                const u32 f = __shfl_sync(gmask, mnl, gshift + l);
                const u32 c = __shfl_sync(gmask, crb, gshift + l);
                const int j = __ffs(f) - 1;
                p = base + 16u * l + (u32)j;
                cr = (c >> j) & 1u;
                break;
            }
            last_cr = __shfl_sync(gmask, mcr, gshift + G - 1) >> 15;
        }
        // (no newline: the line ends with the input, and a CR as its last byte is dropped as well)
        if (p == job.n) cr = job.in[p - 1] == '\r' ? 1u : 0u;  // (p > off: the line holds at least the '>')
        if (sub == 0) {
            u64 H = p - off - cr;
            u64 s0 = sp == ~0ull || sp >= p ? H : sp - off;
            if (s0 > H) s0 = H;
            job.hdr_off[slot] = off;
            job.dbase[slot] = tp.T0 + ev.pos;
            job.loc[slot] = (u64)ev.tile << 16 | ev.pos;
            job.hinfo[slot] = H | s0 << 32;
        }
    }
}

// =====================================================================================
// k_size
// =====================================================================================
struct Cursor {  // a position in the symbol stream: (tile, index in its slot)
    u32 tile;
    int pos;
};
__device__ __forceinline__ void cursor_move(const Job &job, Cursor &c, int delta) {
    c.pos += delta;
    while (c.pos < 0) {
        c.tile--;
        c.pos += job.tile_info[c.tile].nsym;
    }
    while (true) {
        const int n = job.tile_info[c.tile].nsym;
        if (c.pos < n) break;
        c.pos -= n;
        c.tile++;
    }
}
// table index of the window starting at the cursor; the cursor ends on the window's last symbol
__device__ __forceinline__ u32 cursor_window(const Job &job, Cursor &c) {
    const u32 s0 = slot_symbol(job.sym, c.tile, (u32)c.pos);
    cursor_move(job, c, 1);
    const u32 s1 = slot_symbol(job.sym, c.tile, (u32)c.pos);
    cursor_move(job, c, 1);
    const u32 s2 = slot_symbol(job.sym, c.tile, (u32)c.pos);
    return s0 + 5 * s1 + 25 * s2;
}

// Residues kept by --trim (writer.go:141-147,159-163): walk each requested frame backwards from
// its last residue until one is neither 'X' nor '*'.
__device__ void trimmed_lengths(const Job &job, const uint16_t *lut, u64 s, u64 n, u64 L[6]) {
    for (int f = 0; f < 6; f++) {
        if (!((job.frame_mask >> f) & 1u)) continue;
        u64 j = L[f];
        if (j == 0) continue;
        Cursor c;
        if (f < 3) {
            // forward: residue j-1 is the window starting at nucleotide f + 3(j-1); start from the
            // record's end (two symbols before the next record's first nucleotide)
            const u64 l = job.loc[s + 1];
            c.tile = (u32)(l >> 16);
            c.pos = (int)(l & 0xFFFFu);
            const i64 q = (i64)f + 3 * (i64)(j - 1);
            cursor_move(job, c, (int)(q - (i64)n - 2));
            while (true) {
                const u32 idx = cursor_window(job, c);
                const u32 aa = lut[idx < 125 ? idx : 0] & 0xFFu;
                if (aa != 'X' && aa != '*') break;
                if (--j == 0) break;
                cursor_move(job, c, -5);
            }
        } else {
            // reverse: residue j-1 is the window starting at nucleotide n-3-p-3(j-1), i.e. near the
            // record's start for the last residues; walk forwards from there
            const u64 l = job.loc[s];
            c.tile = (u32)(l >> 16);
            c.pos = (int)(l & 0xFFFFu);
            const i64 p = reverse_start(n, f - 3, job.alternative != 0);
            const i64 q = (i64)n - 3 - p - 3 * (i64)(j - 1);
            cursor_move(job, c, (int)q);
            while (true) {
                const u32 idx = cursor_window(job, c);
                const u32 aa = lut[idx < 125 ? idx : 0] >> 8;
                if (aa != 'X' && aa != '*') break;
                if (--j == 0) break;
                cursor_move(job, c, 1);
            }
        }
        L[f] = j;
    }
}

#ifndef GTS_SIZE_CTAS
#define GTS_SIZE_CTAS (1024 / GTS_SIZE_THREADS)
#endif
__global__ void __launch_bounds__(kSizeThreads, GTS_SIZE_CTAS) k_size(const __grid_constant__ Job job) {
    __shared__ uint16_t lut[128];
    __shared__ u64 warp_sum[kSizeThreads / 32];
    __shared__ u64 warp_nt[kSizeThreads / 32];
    __shared__ u32 warp_rec[kSizeThreads / 32];
    __shared__ u64 s_excl;
    __shared__ u32 s_tile;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 128) lut[tid] = job.lut.v[tid];
    pdl_launch();
    pdl_wait();  // k_records' record table
    const u64 R = job.totals->n_headers;
    if (R + 2 > job.rec_cap) return;  // record table overflow: the host retries with a larger one
    const PieceParams &PP = piece_of(job);
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
        u32 flen[kSizePerThread][6];
        u64 sum = 0, nt = 0;
        u32 nrec = 0, nshort = 0;
#pragma unroll
        for (int i = 0; i < kSizePerThread; i++) {
            const u64 s = t * kSizeTile + (u64)tid * kSizePerThread + i;
            size[i] = 0;
            if (s < nslots) {
                u64 D = job.dbase[s];
                u64 n = job.dbase[s + 1] - 2 - D;
                if (PP.piece && s == PP.ov_slot) {  // the whole record's geometry
                    D = PP.ov_dbase;
                    n = PP.ov_n;
                }
                u64 L[6];
                frame_lengths(n, job.alternative != 0, L);
                if (PP.piece && job.ov_trim) {
                    if (s == PP.ov_slot)
                        for (int f = 0; f < 6; f++) L[f] = job.ov_len[f];  // reduced over the pieces (k_piece_trim)
                } else if (job.trim) {
                    trimmed_lengths(job, lut, s, n, L);
#pragma unroll
                    for (int f = 0; f < 6; f++) job.trim_len[s * 6 + f] = (u32)L[f];
                }
                u64 H = s >= 1 ? (job.hinfo[s] & 0xFFFFFFFFull) : 0;
                bool em = record_emitted(s, n, R, job.more_follows != 0);
                if (PP.piece) {
                    em = s == PP.ov_slot;
                    if (em) H = PP.ov_H;
                }
                if (em) {
#pragma unroll
                    for (int f = 0; f < 6; f++)
                        if ((job.frame_mask >> f) & 1u) size[i] += run_size(H, L[f]);
                    nt += n;
                    nrec++;
                }
#pragma unroll
                for (int f = 0; f < 6; f++) flen[i][f] = (em && ((job.frame_mask >> f) & 1u)) ? (u32)L[f] : 0u;
                ri[i].dbase = D;
                ri[i].n = n;
                ri[i].H = (u32)H;
                ri[i].emitted = em ? 1u : 0u;
                if (em && job.short_on && !PP.piece && n <= (u64)kShortMaxN && H <= (u64)kShortMaxH && size[i] <= (u64)kShortMaxOut) {
                    // short enough for k_short -- if its symbols and the two pads behind them lie in one slot,
                    // or go on in the next one and end there
                    const u64 l = job.loc[s];
                    const u32 tl = (u32)(l >> 16);
                    const u64 endp = (l & 0xFFFFu) + n + 2, nsa = job.tile_info[tl].nsym;
                    if (endp <= nsa || (tl + 1 < job.ntiles && endp - nsa <= (u64)job.tile_info[tl + 1].nsym)) {
                        ri[i].emitted = 2u;
                        nshort++;
                    }
                }
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
            nshort += __shfl_xor_sync(0xFFFFFFFFu, nshort, o);
        }
        if (lane == 31) warp_sum[warp] = inc;
        if (lane == 0) {
            warp_nt[warp] = nt;
            warp_rec[warp] = nrec;
            if (nshort) atomicAdd(&job.totals->n_short, (u64)nshort);
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
                if (ri[i].emitted != 2u) {  // (k_short works its records' frame tables out itself: 80 bytes less per read)
                    FrameTab ft;
                    u64 base = off;
#pragma unroll
                    for (int f = 0; f < 6; f++) {
                        ft.body[f] = base + ri[i].H + 3;
                        ft.len[f] = flen[i][f];
                        if (ri[i].emitted && ((job.frame_mask >> f) & 1u)) base += run_size(ri[i].H, flen[i][f]);
                    }
                    ft.pad[0] = ft.pad[1] = 0;
                    job.ftab[s] = ft;
                }
            }
            off += size[i];
        }
    }
}

// =====================================================================================
// Split record (gts_piece_*): the translate pass reuses the slots of the scan pass
// =====================================================================================
// The scan pass parsed the piece without knowing its neighbours.  What changes for the translate
// pass: the two symbols in front of the piece (slot 0's pads become the carried nucleotides), and,
// for the record's last piece, the two end pads behind its last symbol.
__global__ void k_piece_fix(const __grid_constant__ Job job) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    const PieceParams &PP = piece_of(job);
    if (PP.has_carry) {
        const u32 c_prev = (PP.carry >> 4) & 0xFu, c_last = PP.carry & 0xFu;
        u32 w = job.sym[0];  // symbol 0 = low nibble of byte 0, symbol 1 = low nibble of byte 1
        w = (w & ~0x00000F0Fu) | c_prev | c_last << 8;
        job.sym[0] = w;
        TileInfo ti = job.tile_info[0];
        if (ti.nsym == 2) ti.tail = (uint8_t)PP.carry;
        else if (ti.nsym == 3) ti.tail = (uint8_t)((ti.tail & 0xFu) | c_last << 4);
        job.tile_info[0] = ti;
    }
    if (PP.ov_last) {
        const u32 t = job.ntiles - 1;
        TileInfo ti = job.tile_info[t];
        u32 *slot = job.sym + (u64)t * kSlotWords;
        for (u32 i = ti.nsym; i < (u32)ti.nsym + 2u; i++) {  // two zero symbols behind the last one
            const u32 b = i & 7u, sh = 8 * (b & 3) + 4 * (b >> 2);
            slot[i >> 3] &= ~(0xFu << sh);
        }
        ti.nsym = (uint16_t)(ti.nsym + 2);
        ti.tail = 0;
        job.tile_info[t] = ti;
    }
}

// One piece's part of the trim walk (writer.go:141-163): lanes 0..5 = frames.  The residues of a frame
// are checked from its last one backwards while they are 'X' or '*'; a piece can check the windows
// whose first symbol it holds (the first piece also the two windows that start on the record's front
// pads), with the two carried symbols in front and the head of the next piece behind.
__global__ void __launch_bounds__(32) k_piece_trim(const __grid_constant__ Job job) {
    const int f = threadIdx.x;
    if (f >= 6) return;
    u64 len = job.pt_len[f];
    u32 resolved = 0;
    if (!((job.frame_mask >> f) & 1u) || len == 0) {
        resolved = 1;
    } else {
        const PieceParams &PP = job.pv;
        const u64 n = PP.ov_n;
        const bool first = PP.has_carry == 0, last = PP.ov_last != 0;
        const u64 base = first ? 4 : 2;                       // local stream index of the piece's first nucleotide
        const i64 m = (i64)(job.totals->total_sym - base);     // nucleotides of the piece (scan pass: no end pads)
        const i64 own_lo = first ? -2 : 0;
        const i64 step = f < 3 ? -3 : 3;
        i64 q;
        if (f < 3) q = (i64)f + 3 * (i64)(len - 1);
        else q = (i64)n - 3 - reverse_start(n, f - 3, job.alternative != 0) - 3 * (i64)(len - 1);
        i64 i = q - (i64)job.pt_before;
        Cursor c;
        c.tile = 0;
        c.pos = 0;
        i64 c_at = -(i64)base;  // the cursor stands on local stream index 0 = nucleotide index -base
        auto psym = [&](i64 x) -> u32 {
            if (x < 0) return first ? 0u : (x == -1 ? (PP.carry & 0xFu) : ((PP.carry >> 4) & 0xFu));
            if (x >= m) return last ? 0u : (u32)PP.head[x - m];
            if (x != c_at) {
                if (c_at == -(i64)base && x > 4096) {  // first positioning far into the piece: binary search over the tiles
                    u32 lo = 0, hi = job.ntiles - 1;
                    const u64 idx = base + (u64)x;
                    while (lo < hi) {
                        const u32 mid = (lo + hi + 1) >> 1;
                        if (job.tile_pre[mid].T0 <= idx) lo = mid; else hi = mid - 1;
                    }
                    c.tile = lo;
                    c.pos = (int)(idx - job.tile_pre[lo].T0);
                } else {
                    cursor_move(job, c, (int)(x - c_at));
                }
                c_at = x;
            }
            return slot_symbol(job.sym, c.tile, (u32)c.pos);
        };
        while (i >= own_lo && i < m) {
            const u32 idx = psym(i) + 5 * psym(i + 1) + 25 * psym(i + 2);
            const u32 aa = f < 3 ? (job.lut.v[idx] & 0xFFu) : (job.lut.v[idx] >> 8);
            if (aa != 'X' && aa != '*') { resolved = 1; break; }
            if (--len == 0) { resolved = 1; break; }
            i += step;
        }
    }
    job.pt_out[f] = len;
    job.pt_out[6 + f] = resolved;
}

// The piece description of rank `rank` from the all-gathered scan results (what split.plan_pieces
// does on the host): nucleotides before it and in total, the two nucleotides in front of it, the
// 192 behind it, the header length.  Runs between the all-gather and the translate kernels, so the
// step needs no host round trip.
__global__ void k_piece_plan(const PieceInfoDev *all, int rank, int world, PieceParams *out) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    PieceParams p;
    u64 total = 0, before = 0;
    u32 last = 0, second = 0;  // codes of the two nucleotides before the current piece (0 = the pads)
    for (int k = 0; k < world; k++) {
        const u64 m = all[k].nucleotides;
        if (k < rank) {
            before += m;
            if (m >= 2) { last = all[k].tail & 0xFu; second = (all[k].tail >> 4) & 0xFu; }
            else if (m == 1) { second = last; last = all[k].tail & 0xFu; }
        }
        total += m;
    }
    u32 got = 0;
    for (int k = rank + 1; k < world && got < 192u; k++) {
        const u64 m = all[k].nucleotides;
        for (u32 i = 0; i < 192u && i < m && got < 192u; i++) p.head[got++] = all[k].head[i] > 4 ? 0 : all[k].head[i];
    }
    for (; got < 192u; got++) p.head[got] = 0;
    const bool first = rank == 0, lastp = rank == world - 1;
    p.piece = 1;
    p.no_end_pads = lastp ? 0 : 1;
    p.ov_last = lastp ? 1 : 0;
    p.no_headers = first ? 0 : 1;
    p.ov_H = (u32)all[0].header_len;
    p.ov_n = total;
    p.ov_dbase = 4;
    p.pad = 0;
    if (first) {
        p.ov_slot = 1; p.has_carry = 0; p.carry = 0; p.t0_base = 0; p.win_lo = 0;
    } else {
        p.ov_slot = 0; p.has_carry = 1; p.carry = last | second << 4;
        p.t0_base = before + 2;
        p.win_lo = p.t0_base + 2;
    }
    *out = p;
}

cudaError_t launch_piece_plan(const PieceInfoDev *all, int rank, int world, PieceParams *out, cudaStream_t stream) {
    k_piece_plan<<<1, 32, 0, stream>>>(all, rank, world, out);
    return cudaGetLastError();
}

// The totals of a pass, stored straight into (mapped) pinned host memory.  The chunk engine waits for
// an event behind this kernel; a cudaMemcpyAsync in its place would wait in the device-to-host copy
// engine behind the previous chunk's result.
__global__ void k_publish(const Totals *d_totals, Totals *h_totals) {
    static_assert(sizeof(Totals) % 8 == 0 && sizeof(Totals) / 8 <= 32, "one word per lane");
    const u64 *src = reinterpret_cast<const u64 *>(d_totals);
    volatile u64 *dst = reinterpret_cast<volatile u64 *>(h_totals);
    if (threadIdx.x < sizeof(Totals) / 8) dst[threadIdx.x] = src[threadIdx.x];
    __threadfence_system();
}
cudaError_t launch_publish(const Totals *d_totals, Totals *h_totals, cudaStream_t stream) {
    k_publish<<<1, 32, 0, stream>>>(d_totals, h_totals);
    return cudaGetLastError();
}

cudaError_t launch_piece_fix(const Job &job, cudaStream_t stream) {
    k_piece_fix<<<1, 32, 0, stream>>>(job);
    return cudaGetLastError();
}
cudaError_t launch_piece_trim(const Job &job, cudaStream_t stream) {
    k_piece_trim<<<1, 32, 0, stream>>>(job);
    return cudaGetLastError();
}
cudaError_t launch_tile_scan(const Job &job, int sm_count, cudaStream_t stream) {
    const u32 nscan = (job.ntiles + kScanTile - 1) / kScanTile;
    k_tile_scan<<<nscan < (u32)sm_count ? nscan : (u32)sm_count, kScanThreads, 0, stream>>>(job);
    return cudaGetLastError();
}

// =====================================================================================
// k_warn: invalid nucleotides -> (record, position) for the WARNING lines of encodedSequence.go:39
// =====================================================================================
__global__ void __launch_bounds__(256) k_warn(const __grid_constant__ Job job) {
    pdl_wait();  // k_size (the record table is complete after k_records already)
    const u32 n = job.counters[4];
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        job.totals->n_warn = n;
        if (n > job.warn_cap) atomicOr(&job.totals->flags, kFlagWarnOverflow);
    }
    const u64 R = job.totals->n_headers;
    if (R + 2 > job.rec_cap) return;
    const u32 m = n < job.warn_cap ? n : job.warn_cap;
    const u32 stride = gridDim.x * blockDim.x;
    for (u32 i = blockIdx.x * blockDim.x + threadIdx.x; i < m; i += stride) {
        const u64 ev = job.warn[i];
        const u32 tile = (u32)(ev >> 32), spos = (u32)(ev >> 8) & 0xFFFFFFu;
        const u64 S = job.tile_pre[tile].T0 + spos;
        u64 lo = 0, hi = R;  // the record slot holding stream index S: largest s with dbase[s] <= S
        while (lo < hi) {
            const u64 mid = (lo + hi + 1) >> 1;
            if (job.dbase[mid] <= S) lo = mid; else hi = mid - 1;
        }
        WarnRec w;
        w.key = S;
        w.pos = S - job.dbase[lo];
        w.hdr_off = lo ? job.hdr_off[lo] : 0;
        w.H = lo ? (u32)(job.hinfo[lo] & 0xFFFFFFFFull) : 0u;
        w.byte = (u32)(ev & 0xFFu);
        job.wres[i] = w;
    }
}

// Launch with programmatic stream serialization: the kernel's CTAs may be scheduled while the
// kernel before it in the stream drains; it waits in pdl_wait() before touching that kernel's results.
template <typename K>
static cudaError_t launch_pdl(K kernel, dim3 grid, dim3 block, cudaStream_t stream, const Job &job) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = 0;
    cfg.stream = stream;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, job);
}

cudaError_t launch_warn(const Job &job, int sm_count, cudaStream_t stream) {
    return launch_pdl(k_warn, dim3(sm_count), dim3(256), stream, job);
}

cudaError_t launch_sizes(const Job &job, int sm_count, cudaStream_t stream) {
    const u32 nscan = (job.ntiles + kScanTile - 1) / kScanTile;
    cudaError_t e = launch_pdl(k_tile_scan, dim3(nscan < (u32)sm_count ? nscan : (u32)sm_count), dim3(kScanThreads), stream, job);
    if (e != cudaSuccess) return e;
    const u64 rec_blocks = (job.rec_cap * kRecLanes + 255) / 256;  // kRecLanes lanes per header line
    e = launch_pdl(k_records, dim3((u32)(rec_blocks < (u64)sm_count * 8 ? rec_blocks : (u64)sm_count * 8)), dim3(256), stream, job);
    if (e != cudaSuccess) return e;
    const u64 size_tiles = (job.rec_cap + kSizeTile - 1) / kSizeTile;
    u32 size_grid = (u32)(size_tiles < (u64)sm_count * GTS_SIZE_CTAS ? size_tiles : (u64)sm_count * GTS_SIZE_CTAS);
    if (size_grid == 0) size_grid = 1;
    return launch_pdl(k_size, dim3(size_grid), dim3(kSizeThreads), stream, job);
}

}  // namespace gts
