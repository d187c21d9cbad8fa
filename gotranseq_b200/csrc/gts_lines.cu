// gts_lines.cu -- k_lines: symbols -> residue lines, one lane per 60-column output line.
//
// Reference behaviour reproduced (feliixx/gotranseq): translate / translate3Frames / writeAA /
// newLine, transeq/writer.go:57-157; reverseComplement, transeq/encodedSequence.go:71-97 (never
// materialised).
//
// One CTA per parse tile.  The tile's symbols are staged in shared memory one per byte, twice: in
// stream order (forward frames) and mirrored (reverse frames read the mirror upwards, which is the
// stream downwards), with the two symbols before the tile and 192 symbols after it, so that every
// output line whose FIRST codon ends in the tile can be produced by one lane from 189 consecutive
// staged bytes: 63 codons (the line's 60 residues and the first 3 of the next line) through one
// integer dot product and one byte-table load each, '\n' at a fixed position, one byte funnel for
// the line's alignment in the output, 16 word stores into the tile's output image.  Nothing in
// the lane's work depends on the strand except two base addresses.  The image is written out
// with TMA bulk stores as in k_translate.
#include <cstddef>

#include "gts_common.cuh"
#include "gts_kernels.h"

namespace gts {

#ifndef GTS_LINES_THREADS
#define GTS_LINES_THREADS 96
#endif
#ifndef GTS_LINES_CTAS
#define GTS_LINES_CTAS 5
#endif
constexpr int kLinesThreads = GTS_LINES_THREADS;  // >= kLineRuns
constexpr int kHalo = 192;                              // symbols staged before / after the tile's own
constexpr int kSymArr = kHalo + kSlotSyms + kHalo;      // bytes per staged array
constexpr int kMirror = kSlotSyms + kHalo - 1;          // mirror index of local symbol x = kMirror - x; = 7 (mod 8)
constexpr int kLineBatch = 16;                          // record pieces per batch
constexpr int kLineRuns = kLineBatch * 6;
constexpr int kLineImage = 20480;                       // output image bytes (one piece needs < 17.8 KB)
constexpr int kLineMaxLines = kLineImage / 61 + 1;      // line jobs per batch
static_assert(kMirror % 8 == 7, "mirror stores are 8-byte aligned");
static_assert(kMirror + kHalo + 1 <= kSymArr, "mirror array too small");

struct LineRun {     // the lines of one (record, frame) whose first codon ends in the tile
    u64 g0;          // output offset of its first byte
    u32 img;         // byte offset of its first byte in the image
    u32 nbytes;      // bytes to copy out
    u32 sym;         // shared-memory offset of the first line's first symbol (+180 per line)
    u32 lut;         // shared-memory offset of the strand's residue table
    u32 flags;       // bit 0: ends the frame with a partial line (the final newline is patched in)
    u32 pad;
};

struct LinesShared {
    alignas(16) uint8_t symf[kSymArr];  // [kHalo - 2 .. ) = the two symbols before the tile, the tile, the halo
    alignas(16) uint8_t symr[kSymArr];  // symr[kMirror - x] = local symbol x
    alignas(16) uint8_t image[kLineImage];
    alignas(16) LineRun runs[kLineRuns];
    u32 line_pre[kLineRuns];
    alignas(16) uint8_t line_run[kLineRuns];  // the runs that have lines, in line job order
    u32 scan_l[6], scan_i[6], scan_c[6];
    u32 nlisted;
    u32 total_lines;
    alignas(16) uint8_t lutf[128];      // forward residue of s0 + 5 s1 + 25 s2
    alignas(16) uint8_t lutr[128];      // reverse-strand residue of the window read downwards: s2 + 5 s1 + 25 s0
    u32 npieces_done;
};

__device__ __forceinline__ i64 floordiv180(i64 a) {
    i64 q = a / 180;
    if (a - q * 180 < 0) q--;
    return q;
}
__device__ __forceinline__ i64 ceildiv180(i64 a) { return -floordiv180(-a); }

// One output line: 189 staged symbols -> 64 image bytes (60 residues, '\n', 3 residues of the next line).
// All offsets are relative to the start of the CTA's shared memory.  The loads of all 16 codon
// groups come before the first store so that they can be in flight together.
__device__ __forceinline__ void emit_line(uint8_t *smem, u32 symoff, u32 lut, u32 dst, bool first) {
    const u32 sbase = symoff & ~3u;
    const u32 sels = 0x3210u + 0x1111u * (symoff & 3u);
    const u32 a = dst & 3u;
    const u32 sela = 0x3210u + 0x1111u * ((4u - a) & 7u);
    const u32 *sw = reinterpret_cast<const u32 *>(smem + sbase);
    u32 z[16];
    u32 wprev = sw[0];
#pragma unroll
    for (int g = 0; g < 16; g++) {
        const u32 w1 = sw[3 * g + 1], w2 = sw[3 * g + 2], w3 = sw[3 * g + 3];
        const u32 c0 = prmt(wprev, w1, sels), c1 = prmt(w1, w2, sels), c2 = prmt(w2, w3, sels);
        wprev = w3;
        // table offsets of the four codons in these 12 symbols
        const u32 a0 = __dp4a(c0, 0x00190501u, lut);
        const u32 a1 = __dp4a(c1, 0x00001905u, __dp4a(c0, 0x01000000u, lut));
        const u32 a2 = __dp4a(c2, 0x00000019u, __dp4a(c1, 0x05010000u, lut));
        const u32 a3 = __dp4a(c2, 0x19050100u, lut);
        const u32 r0 = smem[a0], r1 = smem[a1], r2 = smem[a2], r3 = smem[a3];
        if (g < 15) {
            z[g] = prmt(prmt(r0, r1, 0x0040u), prmt(r2, r3, 0x0040u), 0x5410u);
        } else {  // residues 60..62 belong to the next line: they follow this line's newline
            z[g] = prmt(prmt(0x0Au, r0, 0x0040u), prmt(r1, r2, 0x0040u), 0x5410u);
        }
    }
    // image word g holds line bytes 4g - a .. 4g - a + 3; a lane owns the words that start in its
    // line (word 0 only when the line starts on a word, or when no line precedes it)
    u32 *dw = reinterpret_cast<u32 *>(smem + (dst & ~3u));
    if (a == 0 || first) dw[0] = prmt(0u, z[0], sela);
#pragma unroll
    for (int g = 1; g < 16; g++) dw[g] = prmt(z[g - 1], z[g], sela);
}

__global__ void __launch_bounds__(kLinesThreads, GTS_LINES_CTAS) k_lines(const __grid_constant__ Job job) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    LinesShared &sh = *reinterpret_cast<LinesShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // tables and margins that no tile overwrites (ordered before their first use by the staging barrier)
    if (tid < 32) {
        reinterpret_cast<u32 *>(sh.lutf)[tid] = job.lutb.f[tid];
        reinterpret_cast<u32 *>(sh.lutr)[tid] = job.lutb.r[tid];
    }
    // margins that are read but never hold a tile's data: below the two carried symbols
    for (int i = tid; i < (kHalo - 2) / 2; i += kLinesThreads) {
        reinterpret_cast<uint16_t *>(sh.symf)[i] = 0;
        reinterpret_cast<uint16_t *>(sh.symr + kMirror + 3)[i] = 0;  // mirror of locals -3, -4, ...
    }
    const u64 flags = job.totals->flags;
    const u64 R = job.totals->n_headers;
    if (flags != 0) return;

    // persistent CTAs: the grid is one wave, each CTA takes every gridDim.x-th tile
    for (u32 tile = blockIdx.x; tile < job.ntiles; tile += gridDim.x) {
    // ---- loads that depend on nothing go first: the slot's words (all of them: words past the
    // tile's symbols hold stale but valid symbols), the start of the next slot, the tile's scalars ----
    constexpr int kWordsPerThread = (kSlotWords + kLinesThreads - 1) / kLinesThreads;
    const u32 *gw = job.sym + (u64)tile * kSlotWords;
    u32 xw[kWordsPerThread];
#pragma unroll
    for (int k = 0; k < kWordsPerThread; k++) {
        const u32 w = tid + k * kLinesThreads;
        xw[k] = w < (u32)kSlotWords ? gw[w] : 0u;
    }
    const bool has_next = tile + 1 < job.ntiles;
    u32 hw = 0;
    if (tid < kHalo / 8 && has_next) hw = gw[kSlotWords + tid];
    const u32 next_nsym = has_next ? job.tile_info[tile + 1].nsym : 0u;
    const TileInfo ti = job.tile_info[tile];
    const TilePrefix tp = job.tile_pre[tile];
    const u32 nsym = ti.nsym;
    if (nsym == 0) continue;
    const u64 T0 = tp.T0, Tend = tp.T0 + nsym;
    const uintptr_t out_mis = (uintptr_t)job.out & 15u;

    // Record pieces of the tile: the record open at its first byte, then one per header line.  A
    // builder thread handles one (piece, frame) of a batch of kLineBatch pieces; thread order =
    // forward runs of all pieces, then the reverse runs, so that the line jobs of a warp read one
    // residue table (no bank conflicts between the two).  The record data of the first batch is
    // requested here, before the staging, so that its latency is hidden.
    const u32 npieces = (u32)ti.nhdr + 1u;
    u32 done = 0;
    const u32 strand = (u32)tid / (3u * kLineBatch), rem = (u32)tid - strand * (3u * kLineBatch);
    const u32 bp = rem / 3u, bf = 3u * strand + (rem - 3u * bp);
    RecInfo ri;
    ri.emitted = 0; ri.dbase = 0; ri.n = 0; ri.H = 0; ri.out_off = 0;
    u32 Lf = 0;
    u64 body = 0;
    bool have = false;
    auto fetch = [&](u32 from) {
        have = false;
        Lf = 0;
        if (tid < kLineRuns && from + bp < npieces) {
            const u64 s = (u64)tp.R0 + from + bp;
            if (s <= R && ((job.frame_mask >> bf) & 1u)) {
                ri = job.rec[s];
                Lf = job.ftab[s].len[bf];
                body = job.ftab[s].body[bf];
                have = true;
            }
        }
    };
    fetch(0);

    // ---- stage: the tile's symbols ----
    if (tid == 0) {
        const u32 c = tp.carry;  // last | second-to-last << 4
        sh.symf[kHalo - 1] = (uint8_t)(c & 0xFu);
        sh.symf[kHalo - 2] = (uint8_t)((c >> 4) & 0xFu);
        sh.symr[kMirror + 1] = (uint8_t)(c & 0xFu);
        sh.symr[kMirror + 2] = (uint8_t)((c >> 4) & 0xFu);
    }
#pragma unroll
    for (int k = 0; k < kWordsPerThread; k++) {
        const u32 w = tid + k * kLinesThreads;
        if (w < (u32)kSlotWords) {
            const u32 lo4 = xw[k] & 0x0F0F0F0Fu, hi4 = (xw[k] >> 4) & 0x0F0F0F0Fu;
            *reinterpret_cast<uint2 *>(sh.symf + kHalo + 8u * w) = make_uint2(lo4, hi4);
            *reinterpret_cast<uint2 *>(sh.symr + kMirror - 7 - 8u * w) =
                make_uint2(prmt(hi4, 0u, 0x0123u), prmt(lo4, 0u, 0x0123u));
        }
    }
    __syncthreads();  // the halo overwrites what the slot holds behind the tile's symbols
    {
        // pull the CTA's next tile towards the SM while this one is being translated
        const u32 nt = tile + gridDim.x;
        if (nt < job.ntiles) {
            const char *base = reinterpret_cast<const char *>(job.sym + (u64)nt * kSlotWords);
            const char *pf = nullptr;
            if (tid < (kSlotWords * 4 + 96 + 127) / 128) pf = base + 128 * tid;
            else if (tid == 40) pf = reinterpret_cast<const char *>(job.tile_info + nt);
            else if (tid == 41) pf = reinterpret_cast<const char *>(job.tile_pre + nt);
            if (pf) asm volatile("prefetch.global.L1 [%0];" ::"l"(pf));
        }
    }
    // the 192 symbols after the tile: the stream goes on in the next slot ...
    if (!has_next || next_nsym >= (u32)kHalo) {
        if (tid < kHalo / 8) {
            const u32 lo4 = hw & 0x0F0F0F0Fu, hi4 = (hw >> 4) & 0x0F0F0F0Fu;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint8_t v = (uint8_t)(((j < 4 ? lo4 : hi4) >> (8 * (j & 3))) & 0xFFu);
                const u32 x = nsym + 8u * tid + j;
                sh.symf[kHalo + x] = v;
                sh.symr[kMirror - (int)x] = v;
            }
        }
    } else {
        // ... or, when that one is short, in the slots after it
        for (u32 i = tid; i < (u32)kHalo; i += kLinesThreads) {
            u32 u = tile + 1, off = i, v = 0;
            while (u < job.ntiles) {
                const u32 nu = job.tile_info[u].nsym;
                if (off < nu) {
                    v = slot_symbol(job.sym, u, off);
                    break;
                }
                off -= nu;
                u++;
            }
            sh.symf[kHalo + nsym + i] = (uint8_t)v;
            sh.symr[kMirror - (int)(nsym + i)] = (uint8_t)v;
        }
    }

    // offsets in the CTA's shared memory
    const u32 symf_addr = (u32)offsetof(LinesShared, symf);
    const u32 symr_addr = (u32)offsetof(LinesShared, symr);
    const u32 lutf_addr = (u32)offsetof(LinesShared, lutf);
    const u32 lutr_addr = (u32)offsetof(LinesShared, lutr);
    const u32 image_addr = (u32)offsetof(LinesShared, image);

    while (done < npieces) {
        // ---- build: one thread per (record piece, frame), kLineBatch pieces per batch ----
        if (tid < kLineRuns) {
            LineRun run;
            run.g0 = 0; run.img = 0; run.nbytes = 0; run.sym = 0; run.lut = 0; run.flags = 0; run.pad = 0;
            u32 nl = 0, need = 0;
            if (have && ri.emitted && Lf) {
                const i64 d = (i64)ri.dbase - (i64)T0;  // the record's first nucleotide, relative to the tile
                const u64 n = ri.n;
                const u32 NL = (Lf + 59u) / 60u;
                u64 lo = 0, hi = 0;  // lines [lo, hi)
                i64 symrel;          // index, in its staged array, of the first symbol of line lo
                if (bf < 3) {
                    // line b: codons start at nucleotide f + 180 b; its first codon ends at k0 + 180 b, k0 = D + f + 2
                    const i64 x = -(d + (i64)bf + 2);  // T0 - k0
                    if (x > 0) {
                        const u64 q = (u64)(x - 1) / 180u;
                        const u32 rem = (u32)((u64)(x - 1) - 180u * q);
                        lo = q + 1;
                        hi = q + (rem + nsym + 180u) / 180u;
                        symrel = kHalo + 177 - (i64)rem;
                    } else {
                        const i64 e = x + (i64)nsym;
                        hi = e > 0 ? ((u32)e + 179u) / 180u : 0u;
                        symrel = kHalo - x - 2;
                    }
                } else {
                    // line b: residues 60 b .. are the codons starting at Q - 180 b, Q - 180 b - 3, ...; the
                    // lowest one used (residue 60 b + 62) ends at D + Q - 180 b - 184, unless the record
                    // starts above it: then at D + 2 + qmod (its lowest codon of this phase).  Such
                    // "clipped" lines are the last one or two of the frame: b >= Lc.
                    const u32 n3 = (u32)(n - 3u * udiv3(n));
                    const u32 i3 = bf - 3u;
                    const u32 pstart = job.alternative ? i3 : (n3 + 3u - i3) % 3u;  // reverse_start
                    const i64 qmod = -2 + (i64)((n3 + 5u - pstart) % 3u);          // (Q + 2) mod 3 - 2
                    u32 Lc = Lf >= 63u ? (Lf - 63u) / 60u + 1u : 0u;
                    if (Lc > NL) Lc = NL;
                    const i64 y = d + (i64)n - 3 - (i64)pstart - 184;  // D + Q - 184 - T0
                    if (y >= 0) {
                        const u64 qy = (u64)y / 180u;
                        const u32 ry = (u32)((u64)y - 180u * qy);
                        hi = qy + 1;
                        if (nsym > ry) {
                            const u32 c1 = (nsym - ry + 179u) / 180u;
                            lo = qy + 1 >= c1 ? qy + 1 - c1 : 0;
                        } else {
                            lo = hi;
                        }
                    }
                    if (hi > Lc) hi = Lc;
                    const i64 kcr = d + 2 + qmod;
                    if (Lc < NL && kcr >= 0 && kcr < (i64)nsym) {
                        if (hi <= lo) lo = Lc;
                        hi = NL;
                    }
                    // the line's highest symbol D + Q + 2 - 180 b is read first, then downwards
                    symrel = (i64)kMirror - (y + 186) + 180 * (i64)lo;
                }
                if (hi > NL) hi = NL;
                if (hi > lo) {
                    const u32 c = (u32)(hi - lo);
                    const bool last = hi == NL;
                    const u32 cut = last ? 60u * NL - Lf : 0u;  // residues missing in the frame's last line
                    run.g0 = body + 61ull * lo;
                    run.nbytes = 61u * c - cut;
                    run.sym = (bf < 3 ? symf_addr : symr_addr) + (u32)symrel;
                    run.lut = bf < 3 ? lutf_addr : lutr_addr;
                    run.flags = (last && cut) ? 1u : 0u;
                    nl = c;
                    const u32 mis = (u32)((out_mis + run.g0) & 15u);
                    need = (mis + 61u * c + 3u + 15u) & ~15u;
                }
            }
            // Inclusive scans over the batch's runs: line jobs, image space, runs that have lines.
            // Pieces taken: all of the batch, unless the image is too small for them: then half as
            // many, and so on (a single piece always fits).
            u32 take = npieces - done < (u32)kLineBatch ? npieces - done : (u32)kLineBatch;
            u32 lp, ip, cp;
            for (int pass = 0;; pass++) {
                if (bp >= take) { nl = 0; need = 0; run.nbytes = 0; }
                lp = nl; ip = need; cp = nl ? 1u : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(0xFFFFFFFFu, lp, o);
                    const u32 z = __shfl_up_sync(0xFFFFFFFFu, ip, o);
                    const u32 c = __shfl_up_sync(0xFFFFFFFFu, cp, o);
                    if (lane >= o) { lp += y; ip += z; cp += c; }
                }
                const int sl = 3 * (pass & 1);  // two sets of slots: no barrier between a read and the next pass's write
                if (lane == 31) { sh.scan_l[sl + warp] = lp; sh.scan_i[sl + warp] = ip; sh.scan_c[sl + warp] = cp; }
                asm volatile("bar.sync 1, %0;" ::"r"(kLineRuns) : "memory");
                const u32 total_need = sh.scan_i[sl] + sh.scan_i[sl + 1] + sh.scan_i[sl + 2];
                for (int w = 0; w < warp; w++) { lp += sh.scan_l[sl + w]; ip += sh.scan_i[sl + w]; cp += sh.scan_c[sl + w]; }
                if (total_need <= (u32)kLineImage || take == 1) break;
                take = (take + 1) >> 1;
            }
            if (nl) {
                const u32 mis = (u32)((out_mis + run.g0) & 15u);
                run.img = ip - need + mis;
                sh.line_pre[cp - 1u] = lp - nl;  // the runs that have lines, in line job order
                sh.line_run[cp - 1u] = (uint8_t)tid;
            }
            sh.runs[tid] = run;
            if (tid == kLineRuns - 1) {
                sh.total_lines = lp;
                sh.nlisted = cp;
                sh.npieces_done = done + take;
            }
        }
        __syncthreads();  // symbols staged, run table built
        const u32 nruns = kLineRuns;
        const u32 total = sh.total_lines;
        const u32 nlisted = sh.nlisted;
        for (u32 j0 = (u32)warp * 32u; j0 < total; j0 += kLinesThreads) {
            // The warp takes line jobs j0 .. j0 + 31.  Listed run of a job: the last one that starts
            // at or before it = (runs starting at or before j0) + (runs starting in (j0, job]).
            const u32 job_i = j0 + lane;
            u32 before = 0, starts = 0;
            for (u32 c0 = 0; c0 < nlisted; c0 += 32) {
                const u32 st = c0 + lane < nlisted ? sh.line_pre[c0 + lane] : 0xFFFFFFFFu;
                before += __popc(__ballot_sync(0xFFFFFFFFu, st <= j0));
                starts |= __reduce_or_sync(0xFFFFFFFFu, (st > j0 && st - j0 < 32u) ? 1u << (st - j0) : 0u);
            }
            if (job_i >= total) continue;
            const u32 lo = before - 1u + __popc(starts & (0xFFFFFFFFu >> (31 - lane)));
            const LineRun &r = sh.runs[sh.line_run[lo]];
            const u32 i = job_i - sh.line_pre[lo];
            emit_line(smem_raw, r.sym + 180u * i, r.lut, image_addr + r.img + 61u * i, i == 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // image writes -> visible to the bulk copy engine
        __syncthreads();                                               // image complete
        for (u32 ri = tid; ri < nruns; ri += blockDim.x) {
            const LineRun r = sh.runs[ri];
            if (r.nbytes) {
                const u32 lo16 = (r.img + 15u) & ~15u, end = r.img + r.nbytes, hi16 = end & ~15u;
                uint8_t *g = job.out + r.g0 - r.img;  // global address of image byte 0 (16-byte congruent)
                if (r.flags & 1) {                    // the frame's final newline (writer.go:164-166)
                    sh.image[end - 1u] = '\n';
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                }
                if (hi16 > lo16) {
                    bulk_store(g + lo16, sh.image + lo16, hi16 - lo16);
                    store_head(sh.image, g, r.img, lo16);
                    store_tail(sh.image, g, hi16, end);
                } else if (end <= lo16) {
                    store_edge(sh.image, g, r.img, end);
                } else {
                    store_edge(sh.image, g, r.img, lo16);
                    store_edge(sh.image, g, lo16, end);
                }
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // image may be reused / CTA may exit
        done = sh.npieces_done;
        fetch(done);
        __syncthreads();
    }
    }  // tiles
}

cudaError_t init_lines() {
    return cudaFuncSetAttribute(k_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinesShared));
}

cudaError_t launch_lines(const Job &job, int sm_count, cudaStream_t stream) {
    const u32 wave = (u32)sm_count * GTS_LINES_CTAS;
    k_lines<<<job.ntiles < wave ? job.ntiles : wave, kLinesThreads, sizeof(LinesShared), stream>>>(job);
    return cudaGetLastError();
}

}  // namespace gts
