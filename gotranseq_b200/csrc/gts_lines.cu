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
#include "gts_common.cuh"
#include "gts_kernels.h"

namespace gts {

constexpr int kLinesThreads = 96;
constexpr int kHalo = 192;                              // symbols staged before / after the tile's own
constexpr int kSymArr = kHalo + kSlotSyms + kHalo;      // bytes per staged array
constexpr int kMirror = kSlotSyms + kHalo - 1;          // mirror index of local symbol x = kMirror - x; = 7 (mod 8)
constexpr int kLineBatch = 16;                          // record pieces per batch
constexpr int kLineRuns = kLineBatch * 6;
constexpr int kLineImage = 20480;                       // output image bytes (one piece needs < 17.8 KB)
static_assert(kMirror % 8 == 7, "mirror stores are 8-byte aligned");
static_assert(kMirror + kHalo + 1 <= kSymArr, "mirror array too small");

struct LineRun {     // the lines of one (record, frame) whose first codon ends in the tile
    u64 g0;          // output offset of its first byte
    u32 img;         // byte offset of its first byte in the image
    u32 nbytes;      // bytes to copy out
    u32 sym;         // shared address of the first line's first symbol (+180 per line)
    u32 lut;         // shared address of the strand's residue table
    u32 flags;       // bit 0: ends the frame with a partial line (the final newline is patched in)
    u32 pad;
};

struct LinesShared {
    alignas(16) uint8_t symf[kSymArr];  // [kHalo - 2 .. ) = the two symbols before the tile, the tile, the halo
    alignas(16) uint8_t symr[kSymArr];  // symr[kMirror - x] = local symbol x
    alignas(16) uint8_t image[kLineImage];
    alignas(16) LineRun runs[kLineRuns];
    u32 line_pre[kLineRuns + 1];
    alignas(16) uint8_t lutf[128];      // forward residue of s0 + 5 s1 + 25 s2
    alignas(16) uint8_t lutr[128];      // reverse-strand residue of the window read downwards: s2 + 5 s1 + 25 s0
    u32 nruns, npieces_done;
};

__device__ __forceinline__ u32 lds_u32(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ u32 lds_u8(u32 addr) {
    u32 v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ void sts_u32(u32 addr, u32 v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ i64 floordiv180(i64 a) {
    i64 q = a / 180;
    if (a - q * 180 < 0) q--;
    return q;
}
__device__ __forceinline__ i64 ceildiv180(i64 a) { return -floordiv180(-a); }

// One output line: 189 staged symbols -> 64 image bytes (60 residues, '\n', 3 residues of the next line).
__device__ __forceinline__ void emit_line(u32 symaddr, u32 lut, u32 dst, bool first) {
    const u32 sbase = symaddr & ~3u;
    const u32 sels = 0x3210u + 0x1111u * (symaddr & 3u);
    const u32 a = dst & 3u;
    const u32 dbase = dst & ~3u;
    const u32 sela = 0x3210u + 0x1111u * ((4u - a) & 7u);
    u32 wprev = lds_u32(sbase);
    u32 zprev = 0;
#pragma unroll
    for (int g = 0; g < 16; g++) {
        const u32 w1 = lds_u32(sbase + 12u * g + 4u);
        const u32 w2 = lds_u32(sbase + 12u * g + 8u);
        const u32 w3 = lds_u32(sbase + 12u * g + 12u);
        const u32 c0 = prmt(wprev, w1, sels), c1 = prmt(w1, w2, sels), c2 = prmt(w2, w3, sels);
        wprev = w3;
        // table addresses of the four codons in these 12 symbols
        const u32 a0 = __dp4a(c0, 0x00190501u, lut);
        const u32 a1 = __dp4a(c1, 0x00001905u, __dp4a(c0, 0x01000000u, lut));
        const u32 a2 = __dp4a(c2, 0x00000019u, __dp4a(c1, 0x05010000u, lut));
        const u32 a3 = __dp4a(c2, 0x19050100u, lut);
        const u32 r0 = lds_u8(a0), r1 = lds_u8(a1), r2 = lds_u8(a2), r3 = lds_u8(a3);
        u32 z;
        if (g < 15) {
            z = prmt(prmt(r0, r1, 0x0040u), prmt(r2, r3, 0x0040u), 0x5410u);
        } else {  // residues 60..62 belong to the next line: they follow this line's newline
            z = prmt(prmt(0x0Au, r0, 0x0040u), prmt(r1, r2, 0x0040u), 0x5410u);
        }
        // image word g holds line bytes 4g - a .. 4g - a + 3; a lane owns the words that start in
        // its line (word 0 only when the line starts on a word, or when no line precedes it)
        const u32 o = prmt(zprev, z, sela);
        if (g > 0 || a == 0 || first) sts_u32(dbase + 4u * g, o);
        zprev = z;
    }
}

__global__ void __launch_bounds__(kLinesThreads) k_lines(const __grid_constant__ Job job) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    LinesShared &sh = *reinterpret_cast<LinesShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (job.totals->flags != 0) return;
    const u64 R = job.totals->n_headers;
    const u32 tile = blockIdx.x;
    const TileInfo ti = job.tile_info[tile];
    const TilePrefix tp = job.tile_pre[tile];
    const u32 nsym = ti.nsym;
    if (nsym == 0) return;
    const u64 T0 = tp.T0, Tend = tp.T0 + nsym;
    const uintptr_t out_mis = (uintptr_t)job.out & 15u;

    // ---- stage: tables, margins, the tile's symbols ----
    for (int i = tid; i < 128; i += kLinesThreads) {
        const int e = i < 125 ? i : 0;
        sh.lutf[i] = (uint8_t)(job.lut.v[e] & 0xFFu);
        const int s0 = e % 5, s1 = (e / 5) % 5, s2 = e / 25;  // read downwards: s0 is the window's LAST symbol
        sh.lutr[i] = (uint8_t)(job.lut.v[s2 + 5 * s1 + 25 * s0] >> 8);
    }
    // margins that are read but never hold the tile's data: below the two carried symbols
    for (int i = tid; i < (kHalo - 2) / 2; i += kLinesThreads) {
        reinterpret_cast<uint16_t *>(sh.symf)[i] = 0;
        reinterpret_cast<uint16_t *>(sh.symr + kMirror + 3)[i] = 0;  // mirror of locals -3, -4, ...
    }
    if (tid == 0) {
        const u32 c = tp.carry;  // last | second-to-last << 4
        sh.symf[kHalo - 1] = (uint8_t)(c & 0xFu);
        sh.symf[kHalo - 2] = (uint8_t)((c >> 4) & 0xFu);
        sh.symr[kMirror + 1] = (uint8_t)(c & 0xFu);
        sh.symr[kMirror + 2] = (uint8_t)((c >> 4) & 0xFu);
    }
    {
        const u32 *gw = job.sym + (u64)tile * kSlotWords;
        const u32 nw = (nsym + 7u) >> 3;
        for (u32 w = tid; w < nw; w += kLinesThreads) {
            const u32 x = gw[w];
            const u32 lo4 = x & 0x0F0F0F0Fu, hi4 = (x >> 4) & 0x0F0F0F0Fu;
            *reinterpret_cast<uint2 *>(sh.symf + kHalo + 8u * w) = make_uint2(lo4, hi4);
            *reinterpret_cast<uint2 *>(sh.symr + kMirror - 7 - 8u * w) =
                make_uint2(prmt(hi4, 0u, 0x0123u), prmt(lo4, 0u, 0x0123u));
        }
    }
    __syncthreads();  // the last word's zero fill is in place before the halo overwrites it
    // the 192 symbols after the tile (the stream goes on in the next slots)
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

    const u32 symf_addr = (u32)__cvta_generic_to_shared(sh.symf);
    const u32 symr_addr = (u32)__cvta_generic_to_shared(sh.symr);
    const u32 lutf_addr = (u32)__cvta_generic_to_shared(sh.lutf);
    const u32 lutr_addr = (u32)__cvta_generic_to_shared(sh.lutr);
    const u32 image_addr = (u32)__cvta_generic_to_shared(sh.image);

    // record pieces of the tile: the record open at its first byte, then one per header line
    const u32 npieces = (u32)ti.nhdr + 1u;
    u32 done = 0;
    while (done < npieces) {
        // ---- build: warp 0, one lane per record piece (16 per batch) ----
        if (warp == 0) {
            const u64 s = (u64)tp.R0 + done + lane;
            LineRun runs[6];
            u32 nl[6];
            u32 lines = 0, need = 0;
#pragma unroll
            for (int f = 0; f < 6; f++) {
                runs[f].g0 = 0; runs[f].img = 0; runs[f].nbytes = 0; runs[f].sym = 0; runs[f].lut = 0;
                runs[f].flags = 0; runs[f].pad = 0;
                nl[f] = 0;
            }
            if (lane < kLineBatch && done + lane < npieces && s <= R) {
                const RecInfo ri = job.rec[s];
                if (ri.emitted) {
                    const i64 D = (i64)ri.dbase;
                    const u64 n = ri.n, H = ri.H;
                    u64 L[6];
                    frame_lengths(n, job.alternative != 0, L);
                    if (job.trim) {
#pragma unroll
                        for (int f = 0; f < 6; f++) L[f] = job.trim_len[s * 6 + f];
                    }
                    u64 run_base = ri.out_off;
#pragma unroll
                    for (int f = 0; f < 6; f++) {
                        if (!((job.frame_mask >> f) & 1u)) continue;
                        const u64 Lf = L[f];
                        const u64 body = run_base + H + 3;
                        run_base += run_size(H, Lf);
                        const i64 NL = (i64)udiv60(Lf + 59);
                        if (NL == 0) continue;
                        i64 lo, hi, sym0;  // lines [lo, hi); sym0 = staged index of line 0's first symbol
                        u32 base_addr;
                        if (f < 3) {
                            // line b: codons start at nucleotide f + 180 b; its first codon ends at D + f + 2 + 180 b
                            const i64 k0 = D + f + 2;
                            lo = ceildiv180((i64)T0 - k0);
                            hi = ceildiv180((i64)Tend - k0);
                            sym0 = kHalo + (k0 - 2 - (i64)T0);
                            base_addr = symf_addr;
                        } else {
                            // line b: residues 60 b .. are the codons starting at Q - 180 b, Q - 180 b - 3, ...;
                            // the lowest one used (residue 60 b + 62) ends at D + Q - 180 b - 184, unless the
                            // record starts above it: then at D + 2 + qmod (its lowest codon of this phase)
                            const i64 Q = (i64)n - 3 - reverse_start(n, f - 3, job.alternative != 0);
                            const i64 qmod = -2 + (Q + 2) % 3;
                            i64 Lc = floordiv180(Q - 186 - qmod) + 1;  // first line whose lowest codon is clipped
                            if (Lc < 0) Lc = 0;
                            if (Lc > NL) Lc = NL;
                            const i64 K0 = D + Q - 184;
                            lo = floordiv180(K0 - (i64)Tend) + 1;
                            hi = floordiv180(K0 - (i64)T0) + 1;
                            if (lo < 0) lo = 0;
                            if (hi > Lc) hi = Lc;
                            const i64 kc = D + 2 + qmod;
                            if (Lc < NL && kc >= (i64)T0 && kc < (i64)Tend) {
                                if (hi > lo) hi = NL;
                                else { lo = Lc; hi = NL; }
                            }
                            // the line's highest symbol D + Q + 2 - 180 b is read first, then downwards
                            sym0 = kMirror - (D + Q + 2 - (i64)T0);
                            base_addr = symr_addr;
                        }
                        if (lo < 0) lo = 0;
                        if (hi > NL) hi = NL;
                        if (hi <= lo) continue;
                        const u32 c = (u32)(hi - lo);
                        const bool last = hi == NL;
                        const u32 cut = last ? (u32)(60 * (u64)NL - Lf) : 0u;  // residues missing in the frame's last line
                        runs[f].g0 = body + 61ull * (u64)lo;
                        runs[f].nbytes = 61u * c - cut;
                        runs[f].sym = base_addr + (u32)(sym0 + 180 * lo);
                        runs[f].lut = f < 3 ? lutf_addr : lutr_addr;
                        runs[f].flags = (last && cut) ? 1u : 0u;
                        nl[f] = c;
                        lines += c;
                        const u32 mis = (u32)((out_mis + runs[f].g0) & 15u);
                        need += (mis + 61u * c + 3u + 15u) & ~15u;
                    }
                }
            }
            // exclusive scans over the pieces: line jobs and image space
            u32 lp = lines, ip = need;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const u32 y = __shfl_up_sync(0xFFFFFFFFu, lp, o);
                const u32 z = __shfl_up_sync(0xFFFFFFFFu, ip, o);
                if (lane >= o) { lp += y; ip += z; }
            }
            // pieces taken: as many as fit the image (at least one: a single piece always fits)
            const unsigned over = __ballot_sync(0xFFFFFFFFu, ip > (u32)kLineImage);
            u32 take = over ? (u32)__ffs(over) - 1u : 32u;
            if (take > (u32)kLineBatch) take = kLineBatch;
            if (take > npieces - done) take = npieces - done;
            const u32 total_lines = __shfl_sync(0xFFFFFFFFu, lp, take - 1);
            lp -= lines;
            ip -= need;
            if ((u32)lane < take) {
#pragma unroll
                for (int f = 0; f < 6; f++) {
                    sh.line_pre[lane * 6 + f] = lp;
                    if (nl[f]) {
                        const u32 mis = (u32)((out_mis + runs[f].g0) & 15u);
                        runs[f].img = ip + mis;
                        ip += (mis + 61u * nl[f] + 3u + 15u) & ~15u;
                        lp += nl[f];
                    }
                    sh.runs[lane * 6 + f] = runs[f];
                }
            }
            if (lane == 0) {
                sh.line_pre[take * 6] = total_lines;
                sh.nruns = take * 6;
                sh.npieces_done = done + take;
            }
        }
        __syncthreads();  // symbols staged, run table built
        const u32 nruns = sh.nruns;
        const u32 total = sh.line_pre[nruns];
        for (u32 job_i = tid; job_i < total; job_i += kLinesThreads) {
            // run r with line_pre[r] <= job_i < line_pre[r + 1]
            u32 lo = 0, hi = nruns;  // invariant: line_pre[lo] <= job_i < line_pre[hi]
            while (hi - lo > 1) {
                const u32 mid = (lo + hi) >> 1;
                if (sh.line_pre[mid] <= job_i) lo = mid; else hi = mid;
            }
            const LineRun &r = sh.runs[lo];
            const u32 i = job_i - sh.line_pre[lo];
            emit_line(r.sym + 180u * i, r.lut, image_addr + r.img + 61u * i, i == 0);
        }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // image writes -> visible to the bulk copy engine
        __syncthreads();                                               // image complete
        for (u32 ri = tid; ri < nruns; ri += kLinesThreads) {
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
        __syncthreads();
    }
}

cudaError_t init_lines() {
    return cudaFuncSetAttribute(k_lines, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(LinesShared));
}

cudaError_t launch_lines(const Job &job, cudaStream_t stream) {
    k_lines<<<job.ntiles, kLinesThreads, sizeof(LinesShared), stream>>>(job);
    return cudaGetLastError();
}

}  // namespace gts
