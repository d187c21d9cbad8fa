// gts_translate.cu -- k_translate and k_headers: symbols -> protein FASTA bytes.
//
// k_translate is the residue kernel of the split-record path (gts_piece_*): every residue is produced
// by the tile that holds the LAST symbol of its codon, so a piece needs only the two symbols in front
// of it.  Ordinary passes use k_lines (gts_lines.cu), which needs the 192 symbols behind a tile instead.
//
// Reference behaviour reproduced (feliixx/gotranseq):
//   k_translate  translate / translate3Frames / writeAA / newLine   transeq/writer.go:57-157,
//                reverseComplement      transeq/encodedSequence.go:71-97 (never materialised:
//                the fused table holds the reverse-strand residue of every window)
//   k_headers    writeHeader            transeq/writer.go:120-131
// tests/gpu_model.py is the executable specification of the index arithmetic used here.
#include "gts_common.cuh"
#include "gts_kernels.h"

namespace gts {

// =====================================================================================
// k_translate
// =====================================================================================
struct Run {        // the part of one (record, frame) body that a chunk writes
    u64 g0;         // output offset of its first byte
    int src;        // byte offset of its first residue in the warp's phase arrays
    u32 img;        // byte offset of its first byte in the warp's output image
    uint16_t naa;   // residues
    uint16_t nbytes;  // residues + newlines
    uint8_t col0;   // column (0..59) of the first residue
    uint8_t flags;  // bit 0: ends the frame (a final newline follows), bit 1: reverse strand
    uint16_t pad;
};

struct RunTable {    // the run table of one batch of record pieces of one chunk
    alignas(16) Run runs[kRunsPerBatch];
    u32 line_pre[kRunsPerBatch + 1];
    u32 more;        // record pieces of the chunk left after this batch
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

// Copy one 60-column line of a run (or the part of it inside the chunk) from the phase arrays
// into the output image, newline included.  Whole aligned image words are written: a lane owns
// the words whose first byte lies in its line (for a run's first line also the word holding the
// line's first byte), so words that straddle two lines carry the start of the next line, and the
// bytes written outside the run are image padding that is never copied out.
__device__ __forceinline__ void copy_line(WarpShared &ws, const Run &r, u32 ln) {
    const uint8_t *ph = ws.phase;
    const u32 col0 = r.col0, naa = r.naa;
    if (ln * 60u >= col0 + naa) return;  // a lane in the gap between the forward and the reverse line jobs
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

// Run table of one batch of record pieces of one chunk.  Called by a group of `width` (4 or 32)
// consecutive lanes, lane `gl` of the group handling record slot s0 + gl (gl < kRecBatch); lane
// kRecBatch of the group, when present, only looks whether the chunk has more record pieces.
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
    u32 nlines_f = 0, nlines_r = 0, img_need = 0;
    if (valid && ri.emitted) {
        const u64 D = ri.dbase, n = ri.n, H = ri.H;
        u64 L[6];
        frame_lengths(n, job.alternative != 0, L);
        if (job.trim) {
#pragma unroll
            for (int f = 0; f < 6; f++) L[f] = job.trim_len[s * 6 + f];
        }
        // windows (by start q) of this record whose last symbol lies in the chunk
        const i64 dq = (i64)D - (i64)T0;  // local index of the record's first symbol
        const u64 Tlo = T0 > job.win_lo ? T0 : job.win_lo;  // (a piece leaves its first windows to the piece before)
        const i64 q_first = (i64)Tlo - (i64)D - 2;
        const i64 q_lo = q_first > -2 ? q_first : -2;
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
            (f < 3 ? nlines_f : nlines_r) += (col0 + naa + 59u) / 60u;
            // image space: 16-byte congruent start, then room for the last word's overhang
            const u32 mis = (u32)((out_mis + runs[f].g0) & 15u);
            img_need += (mis + nb + 3u + 15u) & ~15u;
        }
    }
    // Exclusive scans over the group's records: line jobs and image space.  The table lists the
    // forward runs of all records first and the reverse runs after them, and the reverse runs' line
    // jobs start at a multiple of 32: a round of 32 line jobs is then all forward or all reverse, so
    // copy_line's two strands never diverge inside a warp.
    const unsigned gmask = width == 32 ? 0xFFFFFFFFu : (0xFu << ((threadIdx.x & 31) & ~3));
    u32 lf = nlines_f, lr = nlines_r, ip = img_need;
#pragma unroll
    for (int o = 1; o < kRecBatch; o <<= 1) {
        const u32 x = __shfl_up_sync(gmask, lf, o, width);
        const u32 y = __shfl_up_sync(gmask, lr, o, width);
        const u32 z = __shfl_up_sync(gmask, ip, o, width);
        if (gl >= o) { lf += x; lr += y; ip += z; }
    }
    const u32 total_f = __shfl_sync(gmask, lf, kRecBatch - 1, width);
    const u32 total_r = __shfl_sync(gmask, lr, kRecBatch - 1, width);
    const u32 base_r = total_r ? (total_f + 31u) & ~31u : total_f;
    lf -= nlines_f;
    lr += base_r - nlines_r;
    ip -= img_need;
    if (gl < kRecBatch) {
#pragma unroll
        for (int f = 0; f < 6; f++) {
            const int idx = (f < 3 ? 0 : kRecBatch * 3) + gl * 3 + (f < 3 ? f : f - 3);
            u32 &lp = f < 3 ? lf : lr;
            ws.line_pre[idx] = lp;
            if (runs[f].naa) {
                const u32 mis = (u32)((out_mis + runs[f].g0) & 15u);
                runs[f].img = ip + mis;
                ip += (mis + runs[f].nbytes + 3u + 15u) & ~15u;
                lp += (runs[f].col0 + runs[f].naa + 59u) / 60u;
            }
            ws.runs[idx] = runs[f];
        }
        if (gl == 0) ws.line_pre[kRunsPerBatch] = base_r + total_r;
        if (gl == kRecBatch - 1) {
            ws.more = (valid && next_dbase < Tend) ? 1u : 0u;
            ws.next_slot = (u32)(s + 1);
        }
    }
}

// Persistent CTAs, one parse tile (slot) per iteration.  Warps 0..kTransWarps-1 translate one
// chunk of the slot each (48 symbols per lane); the extra warp builds, one tile ahead, the run table
// of every chunk's first record batch.  One CTA barrier per tile hands the tables over; every warp
// then formats and stores its own chunk.
__global__ void __launch_bounds__(kTransBlock, 4) k_translate(const __grid_constant__ Job job) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    TransShared &sh = *reinterpret_cast<TransShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool builder = warp == kTransWarps;

    const u64 R = job.totals->n_headers;
    if (job.totals->flags != 0) return;
    const u32 ntiles = job.ntiles;
    const uintptr_t out_mis = (uintptr_t)job.out & 15u;

    if (builder) {
        // 4 lanes per chunk: the first kRecBatch record pieces of each
        const int w = lane >> 2, gl = lane & 3;
        u32 it = 0;
        for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
            if (w < kTransWarps) {
                const u32 nsym = job.tile_info[tile].nsym;
                const TilePrefix tp = job.tile_pre[tile];
                const bool sub_valid = (u32)w * kTransSub < nsym;
                const u64 T0 = tp.T0 + (u64)w * kTransSub;
                const u32 cend = (u32)(w + 1) * kTransSub < nsym ? (u32)(w + 1) * kTransSub : nsym;
                const u64 Tend = tp.T0 + cend;
                const u64 s0 = sub_valid ? (u64)tp.R0 + job.chunk_hint[(u64)tile * kTransWarps + w] : 0;
                build_runs(job, sh.w[w].rt[it & 1], R, T0, Tend, sub_valid, s0, gl, 4, out_mis);
            }
            __syncthreads();  // hands tile `it` over
        }
        return;
    }

    if (tid < 128) sh.lut[tid] = job.lut.v[tid];
    asm volatile("bar.sync 1, %0;" ::"r"(kTransWarps * 32) : "memory");  // table staged (translating warps only)
    const uint8_t *lutb = reinterpret_cast<const uint8_t *>(sh.lut);
    WarpShared &ws = sh.w[warp];

    // symbols of this lane in the chunk: 48 = 6 slot words, plus the word before for (e-2, e-1)
    u32 w6[6] = {0, 0, 0, 0, 0, 0}, prev = 0;
    auto load_symbols = [&](u32 tile) {
        if (tile < ntiles) {
            const u32 *wp = job.sym + (u64)tile * kSlotWords + (u32)warp * kTransSubWords + 6 * lane;
            const uint2 x0 = *reinterpret_cast<const uint2 *>(wp);
            const uint2 x1 = *reinterpret_cast<const uint2 *>(wp + 2);
            const uint2 x2 = *reinterpret_cast<const uint2 *>(wp + 4);
            if (warp == 0 && lane == 0) {
                // the two symbols before the slot: symbols 6, 7 of the word "before" = its high nibbles 2, 3
                const u32 c = job.tile_pre[tile].carry;
                prev = ((c >> 4) & 0xFu) << 20 | (c & 0xFu) << 28;
            } else {
                prev = wp[-1];
            }
            w6[0] = x0.x; w6[1] = x0.y; w6[2] = x1.x; w6[3] = x1.y; w6[4] = x2.x; w6[5] = x2.y;
        }
    };
    load_symbols(blockIdx.x);

    u32 it = 0;
    for (u32 tile = blockIdx.x; tile < ntiles; tile += gridDim.x, it++) {
        {
            u32 W[13];  // W[k+1] = symbols 4k..4k+3 of the lane's 48, one per byte; W[0] = the 4 before
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
                        const int i = phi + 3 * (4 * g + x);  // window of the lane's 48
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

        // ---- format: every (record piece, frame) of the chunk becomes a run of 60-column lines ----
        const u32 nsym = job.tile_info[tile].nsym;
        if ((u32)warp * kTransSub >= nsym) continue;
        const u64 tileT0 = job.tile_pre[tile].T0;
        const u64 T0 = tileT0 + (u64)warp * kTransSub;
        const u32 cend = (u32)(warp + 1) * kTransSub < nsym ? (u32)(warp + 1) * kTransSub : nsym;
        const u64 Tend = tileT0 + cend;
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
                        store_head(ws.image, g, r.img, lo16);
                        store_tail(ws.image, g, hi16, end);
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
// Header line of one frame (writer.go:120-131): the header's bytes with '_' and the frame digit
// inserted before its first space (or appended), then '\n'.  Half a warp per record, lane i holding
// output bytes i, i+16, ...: the line is assembled once (only the digit differs between frames)
// and stored once per requested frame.
__global__ void __launch_bounds__(256) k_headers(const __grid_constant__ Job job) {
    const int lane = threadIdx.x & 31, sub = lane & 15, half = lane >> 4;
    const u64 R = job.totals->n_headers;
    if (job.totals->flags != 0 || job.no_headers) return;
    const u64 nhalf = (u64)gridDim.x * (blockDim.x >> 4);
    for (u64 s = ((u64)blockIdx.x * blockDim.x + threadIdx.x) >> 4; s <= R; s += nhalf) {
        const RecInfo ri = job.rec[s];
        if (!ri.emitted) continue;
        const u32 H = ri.H;
        const u32 sp = s >= 1 ? (u32)(job.hinfo[s] >> 32) : 0u;
        const uint8_t *hdr = job.in + job.hdr_off[s];
        // run sizes of the six frames
        u64 L[6];
        frame_lengths(ri.n, job.alternative != 0, L);
        if (job.trim) {
#pragma unroll
            for (int f = 0; f < 6; f++) L[f] = job.trim_len[s * 6 + f];
        }
        for (u32 i = sub; i < H + 3; i += 16) {
            // byte i of the line; the digit at sp+1 is patched per frame
            uint8_t c = '\n';
            if (i < sp) c = hdr[i];
            else if (i == sp) c = '_';
            else if (i > sp + 1 && i < H + 2) c = hdr[i - 2];
            const bool digit = i == sp + 1;
            u64 run_base = ri.out_off;
#pragma unroll
            for (int f = 0; f < 6; f++) {
                if (!((job.frame_mask >> f) & 1u)) continue;
                job.out[run_base + i] = digit ? (uint8_t)('1' + f) : c;
                run_base += run_size(H, L[f]);
            }
        }
    }
    (void)half;
}

// =====================================================================================
// launch
// =====================================================================================
cudaError_t init_lines();
cudaError_t launch_lines(const Job &job, int sm_count, cudaStream_t stream);

cudaError_t init_kernels() {
    cudaError_t e = init_lines();
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_translate, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                (int)sizeof(TransShared));
}

static cudaError_t launch_residues(const Job &job, u32 grid, int sm_count, cudaStream_t stream) {
    if (job.piece) {
        k_translate<<<grid, kTransBlock, sizeof(TransShared), stream>>>(job);
        return cudaGetLastError();
    }
    return launch_lines(job, sm_count, stream);
}

// k_headers only needs the record table, so it runs on `side` (when given) next to k_translate: the
// persistent k_translate grid leaves room for its small CTAs on every SM.
cudaError_t launch_emit(const Job &job, int sm_count, cudaStream_t stream, cudaEvent_t ev_mid, cudaStream_t side,
                        cudaEvent_t ev_fork, cudaEvent_t ev_join) {
    static int per_sm = 0;  // resident CTAs per SM: the persistent grid is exactly one wave
    if (!per_sm) {
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_translate, kTransBlock, sizeof(TransShared)) !=
                cudaSuccess || per_sm < 1)
            per_sm = 3;
    }
    const u32 grid = job.ntiles < (u32)(sm_count * per_sm) ? job.ntiles : (u32)(sm_count * per_sm);
    if (side) {
        cudaEventRecord(ev_fork, stream);
        cudaStreamWaitEvent(side, ev_fork, 0);
        k_headers<<<sm_count * 4, 256, 0, side>>>(job);
        cudaEventRecord(ev_join, side);
        launch_residues(job, grid, sm_count, stream);
        if (ev_mid) cudaEventRecord(ev_mid, stream);
        cudaStreamWaitEvent(stream, ev_join, 0);
    } else {
        launch_residues(job, grid, sm_count, stream);
        if (ev_mid) cudaEventRecord(ev_mid, stream);
        k_headers<<<sm_count * 4, 256, 0, stream>>>(job);
    }
    return cudaGetLastError();
}

}  // namespace gts
