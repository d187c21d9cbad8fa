// gts_lines.cu -- k_lines: symbols -> residue lines, one lane per 180-nucleotide block of one strand.
//
// Reference behaviour reproduced (feliixx/gotranseq): translate / translate3Frames / writeAA /
// newLine, transeq/writer.go:57-157; reverseComplement, transeq/encodedSequence.go:71-97 (never
// materialised).
//
// Persistent CTAs, one parse tile per iteration.  The tile's symbols are staged in shared memory
// one per byte, with the two symbols before the tile and 192 symbols after it.  A lane then takes
// one BLOCK: 180 nucleotides of one record and strand, i.e. output line b of the strand's three
// frames.  It loads the block's 194 symbols once -- upwards for the forward strand, downwards for
// the reverse strand (word step +-4 and a byte-reversing funnel selector are the only difference);
// frame k is the codons at byte offsets k, k+3, ... of them: 63 codons (the line's 60 residues and
// the first 3 of the next line) through one integer dot product and one byte-table load each,
// '\n' at a fixed position, one byte funnel for the line's alignment in the output, 16 word
// stores into the tile's output image.  A block belongs to the tile that holds the last symbol of
// its lowest codon, so the tile's two carried symbols and the 192 after it are all a lane ever
// needs.  The image is written out with TMA bulk stores.
#include <cstddef>
#include <cstdlib>

#include "gts_common.cuh"
#include "gts_emit.cuh"
#include "gts_kernels.h"

namespace gts {

#ifndef GTS_LINES_THREADS
#define GTS_LINES_THREADS 96
#endif
#ifndef GTS_LINES_CTAS
#define GTS_LINES_CTAS 7
#endif
constexpr int kLinesThreads = GTS_LINES_THREADS;  // >= 3 * kBlockRuns
constexpr int kHalo = 192;                              // symbols staged behind the tile's own
constexpr int kLeft = 208;                              // margin in front of them: zeros and the two carried symbols
constexpr int kSymArr = kLeft + kSlotSyms + kHalo + 16;  // bytes of the staged array
constexpr int kLineBatch = 16;                          // record pieces per batch
constexpr int kBlockRuns = kLineBatch * 2;              // (piece, strand) entries per batch: one builder lane each
constexpr int kLineImage = 18432;                       // output image bytes (one piece needs < 17.8 KB)
static_assert(kLeft % 16 == 0, "staging stores are 8-byte aligned");
static_assert(kBlockRuns == 32, "the run table is built by one warp");
static_assert(kLinesThreads >= 3 * kBlockRuns, "one thread per frame run in the copy-out");

struct FrameRun {    // lines of one frame among a BlockRun's blocks
    u64 g0;          // output offset of its first byte
    u32 img;         // byte offset of its first byte in the image
    u32 nbytes;      // bytes to copy out
    u32 nl;          // lines: the first nl blocks of the BlockRun have a line in this frame
    u32 flags;       // bit 0: ends the frame with a partial line (the final newline is patched in)
};

struct BlockRun {    // the blocks of one (record piece, strand) that belong to the tile
    u32 sym;         // shared-memory offset of the first block's first symbol: forward its lowest one
                     // (+180 per block, read upwards), reverse its highest one (-180 per block, read downwards)
    u32 lut;         // shared-memory offset of the strand's residue table
    u32 nblk;
    int dir;         // +1 forward, -1 reverse
    FrameRun fr[3];  // by codon phase: forward frame k; reverse frame that starts k nucleotides from the end
};

struct LinesShared {
    alignas(16) uint8_t symf[kSymArr];  // [kLeft - 2 .. ) = the two symbols before the tile, the tile, the halo
    alignas(16) uint8_t image[kLineImage];
    alignas(16) BlockRun runs[kBlockRuns];
    u32 job_pre[kBlockRuns];            // first block job of the runs that have blocks, in job order
    uint8_t job_run[kBlockRuns];        // their index in runs[]
    alignas(16) uint8_t lutf[128];      // forward residue of s0 + 5 s1 + 25 s2
    alignas(16) uint8_t lutr[128];      // reverse-strand residue of the codon read downwards: s2 + 5 s1 + 25 s0
    u32 total_jobs, nlisted, npieces_done, next_tile;
};

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
    for (int i = tid; i < (kLeft - 2) / 2; i += kLinesThreads) reinterpret_cast<uint16_t *>(sh.symf)[i] = 0;
    const u64 flags = job.totals->flags;
    const u64 R = job.totals->n_headers;
    const PieceParams &PP = piece_of(job);
    if (flags != 0) return;
    if (job.totals->records != 0 && job.totals->n_short == job.totals->records) return;  // reads only: all k_short's

    // persistent CTAs: the grid is one wave; a CTA starts with tile blockIdx.x and claims its next
    // tile from a counter one iteration ahead (tiles differ in cost: the wave stays balanced)
    for (u32 tile = blockIdx.x, tile_next = 0; tile < job.ntiles; tile = tile_next) {
#ifdef GTS_PHASE_TIMING
    u64 tprev = gtime();
#endif
    if (tid == 0) sh.next_tile = gridDim.x + atomicAdd(&job.counters[3], 1u);
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
    if (nsym == 0) {  // nothing ends in this tile
        __syncthreads();
        tile_next = sh.next_tile;
        __syncthreads();
        continue;
    }
    const u64 T0 = tp.T0, Tend = tp.T0 + nsym;
    const u32 tlo = PP.win_lo > T0 ? (u32)(PP.win_lo - T0) : 0u;  // 2 in the first tile of a later piece of a split record
    const uintptr_t out_mis = (uintptr_t)job.out & 15u;

    // Record pieces of the tile: the record open at its first byte, then one per header line.  A
    // builder lane (warp 0) handles one (piece, strand) of a batch of kLineBatch pieces; lane order =
    // forward runs of all pieces, then the reverse runs, so that the block jobs of a warp mostly
    // read one residue table (no bank conflicts between the two).  The record data of the first
    // batch is requested here, before the staging, so that its latency is hidden.
    const u32 npieces = (u32)ti.nhdr + 1u;
    u32 done = 0;
    const u32 bp = (u32)lane & (kLineBatch - 1), bs = (u32)lane / kLineBatch;
    RecInfo ri;
    ri.emitted = 0; ri.dbase = 0; ri.n = 0; ri.H = 0; ri.out_off = 0;
    u32 flen[3] = {0, 0, 0};   // the strand's three frames, in frame order
    u64 fbody[3] = {0, 0, 0};
    auto fetch = [&](u32 from) {
        ri.emitted = 0;
        if (warp == 0 && from + bp < npieces) {
            const u64 s = (u64)tp.R0 + from + bp;
            if (s <= R) {
                ri = job.rec[s];
                const FrameTab *ft = job.ftab + s;
#pragma unroll
                for (int i = 0; i < 3; i++) {
                    flen[i] = ft->len[3 * bs + i];
                    fbody[i] = ft->body[3 * bs + i];
                }
            }
        }
    };
    fetch(0);

    // ---- stage: the tile's symbols ----
    if (tid == 0) {
        const u32 c = tp.carry;  // last | second-to-last << 4
        sh.symf[kLeft - 1] = (uint8_t)(c & 0xFu);
        sh.symf[kLeft - 2] = (uint8_t)((c >> 4) & 0xFu);
    }
#pragma unroll
    for (int k = 0; k < kWordsPerThread; k++) {
        const u32 w = tid + k * kLinesThreads;
        if (w < (u32)kSlotWords) {
            const u32 lo4 = xw[k] & 0x0F0F0F0Fu, hi4 = (xw[k] >> 4) & 0x0F0F0F0Fu;
            *reinterpret_cast<uint2 *>(sh.symf + kLeft + 8u * w) = make_uint2(lo4, hi4);
        }
    }
    __syncthreads();  // the halo overwrites what the slot holds behind the tile's symbols
    PHASE_MARK(job, 16, tprev);  // loads + staging
    {
        // pull the CTA's next tile towards the SM while this one is being translated
        const u32 nt = tile_next = sh.next_tile;
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
    const bool from_head = PP.piece && PP.no_end_pads;  // a split record goes on in the next piece
    if (has_next ? next_nsym >= (u32)kHalo : !from_head) {
        if (tid < kHalo / 8) {
            const u32 lo4 = hw & 0x0F0F0F0Fu, hi4 = (hw >> 4) & 0x0F0F0F0Fu;
#pragma unroll
            for (int j = 0; j < 8; j++) {
                const uint8_t v = (uint8_t)(((j < 4 ? lo4 : hi4) >> (8 * (j & 3))) & 0xFFu);
                const u32 x = nsym + 8u * tid + j;
                sh.symf[kLeft + x] = v;
            }
        }
    } else {
        // ... or, when that one is short, in the slots after it
        for (u32 i = tid; i < (u32)kHalo; i += kLinesThreads) {
            u32 u = tile + 1, off = i, v = 0;
            bool found = false;
            while (u < job.ntiles) {
                const u32 nu = job.tile_info[u].nsym;
                if (off < nu) {
                    v = slot_symbol(job.sym, u, off);
                    found = true;
                    break;
                }
                off -= nu;
                u++;
            }
            if (!found && from_head) v = PP.head[off];  // off < kHalo symbols past the end of the piece
            sh.symf[kLeft + nsym + i] = (uint8_t)v;
        }
    }

    // offsets in the CTA's shared memory
    const u32 smem_base = (u32)__cvta_generic_to_shared(smem_raw);  // addresses in the shared window
    const u32 symf_addr = (u32)offsetof(LinesShared, symf);         // (relative: the symbol loads go through pointers)
    const u32 lutf_addr = smem_base + (u32)offsetof(LinesShared, lutf);
    const u32 lutr_addr = smem_base + (u32)offsetof(LinesShared, lutr);
    const u32 image_addr = smem_base + (u32)offsetof(LinesShared, image);

    while (done < npieces) {
        // ---- build (warp 0): one lane per (record piece, strand) ----
        if (warp == 0) {
            BlockRun run;
            run.sym = 0; run.lut = 0; run.nblk = 0; run.dir = 1;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                run.fr[k].g0 = 0; run.fr[k].img = 0; run.fr[k].nbytes = 0; run.fr[k].nl = 0; run.fr[k].flags = 0;
            }
            u32 need = 0;
            if (ri.emitted == 1) {  // (2: the whole record is written by k_short)
                const i64 d = (i64)ri.dbase - (i64)T0;  // the record's first nucleotide, relative to the tile
                const u64 n = ri.n;
                // the strand's frames by codon phase k: forward frame k; the reverse frame whose first
                // codon starts k nucleotides from the record's end (reverse_start == k)
                u32 Lk[3];
                u64 Bk[3];
                if (bs == 0 || job.alternative) {
#pragma unroll
                    for (int k = 0; k < 3; k++) { Lk[k] = flen[k]; Bk[k] = fbody[k]; }
                } else {
                    const u32 n3 = (u32)(n - 3u * udiv3(n));  // frame i starts at (n3 - i) mod 3
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        const u32 i = (n3 + 3u - (u32)k) % 3u;
                        Lk[k] = i == 0 ? flen[0] : i == 1 ? flen[1] : flen[2];
                        Bk[k] = i == 0 ? fbody[0] : i == 1 ? fbody[1] : fbody[2];
                    }
                }
                u32 NLk[3], NLmax = 0;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    NLk[k] = (Lk[k] + 59u) / 60u;
                    NLmax = NLk[k] > NLmax ? NLk[k] : NLmax;
                }
                u64 lo = 0, hi = 0;  // blocks [lo, hi)
                i64 symrel = 0;      // index, in the staged array, of the first symbol block lo reads
                if (NLmax) {
                    if (bs == 0) {
                        // block b = nucleotides 180 b ..; its lowest codon (frame 1) ends at k0 + 180 b, k0 = D + 2
                        // (tlo: a piece of a split record leaves the anchors on its first two symbols,
                        // the stand-ins for the nucleotides before it, to the piece before)
                        const i64 x = -(d + 2) + (i64)tlo;  // first anchor position of the tile - k0
                        if (x > 0) lo = (u64)(x - 1) / 180u + 1;
                        const i64 e = (i64)nsym - (d + 2) - 180 * (i64)lo;  // anchors k0 + 180 b < T0 + nsym
                        hi = lo + (e > 0 ? ((u32)e + 179u) / 180u : 0u);
                        symrel = kLeft + d + 180 * (i64)lo;
                    } else {
                        // block b is read downwards from nucleotide n - 1 - 180 b; its lowest codon (phase 2,
                        // residue 60 b + 62) ends at D + n - 189 - 180 b, unless the record starts above it:
                        // then ("clipped", b >= Lc) at D, the end of the codon made of the record's pads
                        u64 Lc = n >= 189 ? (n - 189) / 180u + 1 : 0;
                        if (Lc > NLmax) Lc = NLmax;
                        const i64 y = d + (i64)n - 189;  // anchor of block 0, relative to the tile
                        if (y >= (i64)tlo) {
                            const u64 qy = (u64)(y - tlo) / 180u;
                            const u32 ry = (u32)((u64)(y - tlo) - 180u * qy);
                            hi = qy + 1;
                            if (nsym - tlo > ry) {
                                const u32 c1 = (nsym - tlo - ry + 179u) / 180u;
                                lo = qy + 1 >= c1 ? qy + 1 - c1 : 0;
                            } else {
                                lo = hi;
                            }
                        }
                        if (hi > Lc) hi = Lc;
                        if (Lc < NLmax && d >= (i64)tlo && d < (i64)nsym) {
                            if (hi <= lo) lo = Lc;
                            hi = NLmax;
                        }
                        symrel = kLeft + (y + 188) - 180 * (i64)lo;  // the block's highest symbol
                    }
                    if (hi > NLmax) hi = NLmax;
                }
                if (hi > lo) {
                    const u32 c = (u32)(hi - lo);
                    run.sym = symf_addr + (u32)symrel;
                    run.dir = bs == 0 ? 1 : -1;
                    run.lut = bs == 0 ? lutf_addr : lutr_addr;
                    run.nblk = c;
#pragma unroll
                    for (int k = 0; k < 3; k++) {
                        if ((u64)NLk[k] <= lo) continue;  // (also frames that were not requested: no lines)
                        const u64 left = (u64)NLk[k] - lo;
                        const u32 nl = left < c ? (u32)left : c;
                        const bool last = lo + nl == NLk[k];
                        const u32 cut = last ? 60u * NLk[k] - Lk[k] : 0u;  // residues missing in the frame's last line
                        run.fr[k].g0 = Bk[k] + 61ull * lo;
                        run.fr[k].nbytes = 61u * nl - cut;
                        run.fr[k].nl = nl;
                        run.fr[k].flags = (last && cut) ? 1u : 0u;
                        const u32 mis = (u32)((out_mis + run.fr[k].g0) & 15u);
                        need += (mis + 61u * nl + 3u + 15u) & ~15u;
                    }
                }
            }
            // Inclusive scans over the batch's runs: block jobs, image space, runs that have blocks.
            // Pieces taken: all of the batch, unless the image is too small for them: then half as
            // many, and so on (a single piece always fits).
            u32 take = npieces - done < (u32)kLineBatch ? npieces - done : (u32)kLineBatch;
            u32 jp, ip, cp;
            while (true) {
                if (bp >= take) { run.nblk = 0; need = 0; run.fr[0].nbytes = run.fr[1].nbytes = run.fr[2].nbytes = 0; }
                jp = run.nblk; ip = need; cp = run.nblk ? 1u : 0u;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const u32 y = __shfl_up_sync(0xFFFFFFFFu, jp, o);
                    const u32 z = __shfl_up_sync(0xFFFFFFFFu, ip, o);
                    const u32 c = __shfl_up_sync(0xFFFFFFFFu, cp, o);
                    if (lane >= o) { jp += y; ip += z; cp += c; }
                }
                const u32 total_need = __shfl_sync(0xFFFFFFFFu, ip, 31);
                if (total_need <= (u32)kLineImage || take == 1) break;
                take = (take + 1) >> 1;
            }
            if (run.nblk) {
                u32 img = ip - need;
#pragma unroll
                for (int k = 0; k < 3; k++) {
                    if (run.fr[k].nl) {
                        const u32 mis = (u32)((out_mis + run.fr[k].g0) & 15u);
                        run.fr[k].img = img + mis;
                        img += (mis + 61u * run.fr[k].nl + 3u + 15u) & ~15u;
                    }
                }
                sh.job_pre[cp - 1u] = jp - run.nblk;  // the runs that have blocks, in job order
                sh.job_run[cp - 1u] = (uint8_t)lane;
            }
            sh.runs[lane] = run;
            if (lane == 31) {
                sh.total_jobs = jp;
                sh.nlisted = cp;
                sh.npieces_done = done + take;
            }
        }
        __syncthreads();  // symbols staged, run table built
        PHASE_MARK(job, 17, tprev);  // halo + build
        const u32 total = sh.total_jobs;
        const u32 nlisted = sh.nlisted;
        for (u32 j0 = (u32)warp * 32u; j0 < total; j0 += kLinesThreads) {
            // The warp takes block jobs j0 .. j0 + 31.  Listed run of a job: the last one that starts
            // at or before it = (runs starting at or before j0) + (runs starting in (j0, job]).
            const u32 job_i = j0 + lane;
            const u32 st = (u32)lane < nlisted ? sh.job_pre[lane] : 0xFFFFFFFFu;
            const u32 before = __popc(__ballot_sync(0xFFFFFFFFu, st <= j0));
            const u32 starts = __reduce_or_sync(0xFFFFFFFFu, (st > j0 && st - j0 < 32u) ? 1u << (st - j0) : 0u);
            if (job_i >= total) continue;
            const u32 li = before - 1u + __popc(starts & (0xFFFFFFFFu >> (31 - lane)));
            const BlockRun &r = sh.runs[sh.job_run[li]];
            const u32 i = job_i - sh.job_pre[li];
            // the block's 194 symbols, four per word, in the order the strand reads them: 50 words at
            // x0, x0 + step, ...; symbol 4m..4m+3 = a byte window of words m, m+1
            const int dir = r.dir;
            const u32 symoff = r.sym + (u32)(dir * 180 * (int)i);
            u32 x0, sels;
            if (dir > 0) {
                x0 = symoff & ~3u;
                sels = 0x3210u + 0x1111u * (symoff & 3u);
            } else {
                const u32 a = symoff - 3u;  // lowest byte of the first group of four
                x0 = (a & ~3u) + 4u;
                sels = (0x0123u + 0x1111u * (a & 3u)) ^ 0x4444u;  // bytes reversed; word m is the HIGH word of pair m
            }
            u32 c[49];
            {
                // (constant word offsets in both directions: no address arithmetic between the loads)
                const u32 *sp = reinterpret_cast<const u32 *>(smem_raw + x0);
                u32 wprev = sp[0];
                if (dir > 0) {
#pragma unroll
                    for (int m = 0; m < 49; m++) {
                        const u32 w = sp[m + 1];
                        c[m] = prmt(wprev, w, sels);
                        wprev = w;
                    }
                } else {
#pragma unroll
                    for (int m = 0; m < 49; m++) {
                        const u32 w = sp[-(m + 1)];
                        c[m] = prmt(wprev, w, sels);
                        wprev = w;
                    }
                }
            }
            const u32 lut = r.lut;
            if (i < r.fr[0].nl) emit_frame<0>(c, lut, image_addr + r.fr[0].img + 61u * i, i == 0);
            if (i < r.fr[1].nl) emit_frame<1>(c, lut, image_addr + r.fr[1].img + 61u * i, i == 0);
            if (i < r.fr[2].nl) emit_frame<2>(c, lut, image_addr + r.fr[2].img + 61u * i, i == 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // image writes -> visible to the bulk copy engine
        __syncthreads();                                               // image complete
        PHASE_MARK(job, 18, tprev);  // block jobs
        if (tid < 3 * kBlockRuns) {
            const FrameRun r = sh.runs[tid / 3].fr[tid % 3];
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
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // image may be reused
        done = sh.npieces_done;
        fetch(done);
        __syncthreads();
        PHASE_MARK(job, 19, tprev);  // copy-out
    }
    }  // tiles
}

cudaError_t init_lines() {
    return cudaFuncSetAttribute(k_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinesShared));
}

cudaError_t launch_lines(const Job &job, int sm_count, cudaStream_t stream) {
    // (A/B switch: GTS_LINES_WAVE = CTAs per SM of the persistent grid, 1..GTS_LINES_CTAS; a smaller
    // wave leaves registers for another pass's k_parse on the same SM, profiles/r2_notes.md)
    // (GTS_TUNE=1: read again at every launch, so that one process can sweep it)
    static const bool tune = std::getenv("GTS_TUNE") != nullptr;
    static const int wave_static = std::getenv("GTS_LINES_WAVE") ? std::atoi(std::getenv("GTS_LINES_WAVE")) : 0;
    const int wave_env = tune && std::getenv("GTS_LINES_WAVE") ? std::atoi(std::getenv("GTS_LINES_WAVE")) : wave_static;
    const u32 per_sm = wave_env >= 1 && wave_env <= GTS_LINES_CTAS ? (u32)wave_env : (u32)GTS_LINES_CTAS;
    const u32 wave = (u32)sm_count * per_sm;
    k_lines<<<job.ntiles < wave ? job.ntiles : wave, kLinesThreads, sizeof(LinesShared), stream>>>(job);
    return cudaGetLastError();
}

}  // namespace gts
