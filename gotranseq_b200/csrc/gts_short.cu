// gts_short.cu -- k_short: the whole output of SHORT records (reads), one warp per 16 records.
//
// Reference behaviour reproduced (feliixx/gotranseq): (*writer).translate / translate3Frames /
// writeHeader / writeAA / trimAndReturn, transeq/writer.go:57-169, for records whose every frame
// is a single output line (n <= 180) -- the shape of BASELINE config 5 (150-bp reads).
//
// k_lines is built around long records (a tile's symbols staged once, one lane per 180-nucleotide
// block, a run table per tile); with a hundred reads per tile its per-record bookkeeping and
// k_headers' byte stores dominate.  Here a warp takes 16 consecutive record slots.  k_size marks
// the records that qualify (RecInfo.emitted == 2: n <= 180, output <= 768 bytes, symbols inside one
// slot); k_lines and k_headers skip them.  Lane (r, strand) loads record r's symbols straight from
// the packed slot words (26 words), unpacks them in registers to the byte-per-symbol form of
// k_lines and emits the strand's three frames with the same emit_frame<K> into the record's output
// image in shared memory (whole words: up to three bytes spill into the neighbouring header lines);
// the six header lines are then laid over the image, and every record leaves with one TMA bulk store.
#include <cstddef>

#include "gts_common.cuh"
#include "gts_emit.cuh"
#include "gts_kernels.h"

namespace gts {

constexpr int kShortWarps = 4;
constexpr int kShortRecs = 16;         // records per warp and round: lanes = (record, strand)
constexpr int kShortSlot = kShortMaxOut + 16;  // image bytes per record: its output at (address & 15)
constexpr int kShortImage = kShortRecs * kShortSlot + 80;

struct ShortMeta {       // per record of the round, written by lane r < 16
    u32 img[6];          // image offset of each frame's header line
    u32 H, sp;           // header line bytes, index of its first space
    u64 hoff;            // raw offset of the header line
    u64 g0;              // global address of image byte 0 of the record's slot
    u32 lo, hi;          // the record's output in its slot: image bytes [lo, hi)
    u32 ok;              // handled here
    u32 pad;
};

constexpr int kShortLine = 76;  // 4 + (kShortMaxH + 3) + slack, a multiple of 4 that is odd in words (19: no bank conflicts)
struct ShortWarp {
    alignas(16) uint8_t image[kShortImage];
    alignas(16) uint8_t lines[kShortRecs][kShortLine];  // the records' header lines "id_N rest\n" at [4, 4 + H + 3); N patched per frame
    ShortMeta meta[kShortRecs];
};
struct ShortShared {
    ShortWarp w[kShortWarps];
    alignas(16) uint8_t lutf[128];
    alignas(16) uint8_t lutr[128];
};

constexpr u32 kShortBias = 256;  // symbols: windows may start below a slot's first symbol (32 packed words)

// c[m] = the four symbols start + 4m .. start + 4m + 3 of a slot (start biased by kShortBias), one per
// byte, m = 0..48: the slot's packed words (two symbols per byte, gts_device.h) unpacked in registers.
// kMerge: only the symbols from `split` on (counted from c[0]'s first) are replaced -- the part of a
// record that lies in the next slot.
template <bool kMerge>
__device__ __forceinline__ void sym_words_up(const u32 *slot, u32 start, u32 split, u32 (&c)[49]) {
    constexpr u32 M = 0x0F0F0F0Fu;
    const u32 k0 = start >> 3, par = (start >> 2) & 1u;
    const u32 sels = 0x3210u + 0x1111u * (start & 3u);
    const u32 sh_even = 4u * par, sh_odd = 4u * (1u - par);
    u32 P[27];
#pragma unroll
    for (int i = 0; i < 27; i++) {
        const int w = (int)(k0 + (u32)i) - (int)(kShortBias / 8);
        P[i] = (w >= 0 && w < kSlotWords) ? slot[w] : 0u;
    }
    const u32 ms = split >> 2, rs = split & 3u;
    // the split word takes its first rs bytes from what is there already
    const u32 selm = rs == 0 ? 0x7654u : rs == 1 ? 0x7650u : rs == 2 ? 0x7610u : 0x7210u;
    u32 wprev = (P[0] >> sh_even) & M;
#pragma unroll
    for (int m = 0; m < 49; m++) {
        const int t = m + 1;  // byte word t of the unpacked sequence
        u32 w;
        if (t & 1) w = ((par ? P[(t + 1) / 2] : P[(t - 1) / 2]) >> sh_odd) & M;
        else w = (P[t / 2] >> sh_even) & M;
        const u32 v = prmt(wprev, w, sels);
        wprev = w;
        if (!kMerge) c[m] = v;
        else if ((u32)m > ms) c[m] = v;
        else if ((u32)m == ms) c[m] = prmt(c[m], v, selm);
    }
}

// c[m] = the four symbols top - 4m, top - 4m - 1, .. - 3 (downwards from `top`, biased), one per byte.
// kMerge: only the first `split` symbols are replaced (the record's top lies in the next slot).
template <bool kMerge>
__device__ __forceinline__ void sym_words_down(const u32 *slot, u32 top, u32 split, u32 (&c)[49]) {
    constexpr u32 M = 0x0F0F0F0Fu;
    const u32 a = top - 3u;  // lowest symbol of the first group of four
    const u32 jt = (a >> 2) + 1u, kt = jt >> 1, par = jt & 1u;
    const u32 sels = (0x0123u + 0x1111u * (a & 3u)) ^ 0x4444u;  // bytes reversed; word m is the HIGH word of pair m
    const u32 sh_even = 4u * par, sh_odd = 4u * (1u - par);
    u32 P[27];  // packed words downwards
#pragma unroll
    for (int i = 0; i < 27; i++) {
        const int w = (int)kt - i - (int)(kShortBias / 8);
        P[i] = (w >= 0 && w < kSlotWords) ? slot[w] : 0u;
    }
    const u32 ms = split >> 2, rs = split & 3u;
    // the split word takes its first rs bytes from the new words, the others from what is there already
    const u32 selm = rs == 0 ? 0x3210u : rs == 1 ? 0x3214u : rs == 2 ? 0x3254u : 0x3654u;
    u32 wprev = (P[0] >> sh_even) & M;
#pragma unroll
    for (int m = 0; m < 49; m++) {
        const int t = m + 1;  // byte word jt - t
        u32 w;
        if (t & 1) w = ((par ? P[(t - 1) / 2] : P[(t + 1) / 2]) >> sh_odd) & M;
        else w = (P[t / 2] >> sh_even) & M;
        const u32 v = prmt(wprev, w, sels);
        wprev = w;
        if (!kMerge) c[m] = v;
        else if ((u32)m < ms) c[m] = v;
        else if ((u32)m == ms) c[m] = prmt(c[m], v, selm);
    }
}

__global__ void __launch_bounds__(kShortWarps * 32) k_short(const __grid_constant__ Job job) {
    extern __shared__ __align__(16) uint8_t smem_raw[];
    ShortShared &sh = *reinterpret_cast<ShortShared *>(smem_raw);
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 32) {
        reinterpret_cast<u32 *>(sh.lutf)[tid] = job.lutb.f[tid];
        reinterpret_cast<u32 *>(sh.lutr)[tid] = job.lutb.r[tid];
    }
    __syncthreads();
    const u64 R = job.totals->n_headers;
    if (job.totals->flags != 0 || job.totals->n_short == 0 || piece_of(job).piece) return;
    ShortWarp &ws = sh.w[warp];
    const u32 smem_base = (u32)__cvta_generic_to_shared(smem_raw);
    const u32 image_addr = smem_base + (u32)(offsetof(ShortShared, w) + warp * sizeof(ShortWarp) + offsetof(ShortWarp, image));
    const u32 lut_addr = smem_base + (u32)(lane < 16 ? offsetof(ShortShared, lutf) : offsetof(ShortShared, lutr));
    const int r = lane & 15, strand = lane >> 4;
    const u64 ngroups = (R + 1 + kShortRecs - 1) / kShortRecs;
    for (u64 g = (u64)blockIdx.x * kShortWarps + warp; g < ngroups; g += (u64)gridDim.x * kShortWarps) {
        const u64 s = g * kShortRecs + r;
        RecInfo ri;
        ri.emitted = 0; ri.n = 0; ri.H = 0; ri.out_off = 0; ri.dbase = 0;
        if (s <= R) ri = job.rec[s];
        const bool ok = ri.emitted == 2;
        if (!__any_sync(0xFFFFFFFFu, ok)) continue;
        const u32 n = (u32)ri.n;
        // The record's frame table (k_size writes FrameTab only for the records of k_lines / k_headers):
        // residues of every requested frame (frame_lengths of gts_common.cuh for n <= 180; --trim: the
        // lengths k_size found) and where its first residue lies, relative to the record's first output byte.
        u32 flen6[6], body6[6], rec_bytes = 0;
        {
            const u32 n3r = n - 3u * ((n * 43691u) >> 17);  // n % 3
#pragma unroll
            for (int f = 0; f < 6; f++) {
                u32 start = (u32)(f % 3);
                if (f >= 3 && !job.alternative) start = (n3r + 3u - start) % 3u;  // reverse_start (writer.go:64-79)
                u32 L = n >= start ? ((n - start + 2u) * 43691u) >> 17 : 0u;     // ceil((n - start) / 3)
                if (job.trim && ok) L = job.trim_len[s * 6 + f];
                const bool req = ((job.frame_mask >> f) & 1u) != 0;
                flen6[f] = req ? L : 0u;
                body6[f] = rec_bytes + ri.H + 3u;
                if (req) rec_bytes += ri.H + 3u + L + (L ? 1u : 0u);  // run_size for one line at most
            }
        }
        // ---- the record's slot in the image, its frames ----
        u64 gaddr = 0;
        u32 lo = 0;
        u32 flen[3] = {0, 0, 0};   // this strand's frames by codon phase
        u32 fdst[3] = {0, 0, 0};   // image address of the frame's first residue
        u32 loc_tile = 0, loc_pos = 0, nsA = 0;
        bool cross = false;
        // all 16 qualify: their outputs are one contiguous piece of the output and of the image (one bulk
        // store for the group); otherwise every record sits in a slot of its own
        const bool all_ok = __all_sync(0xFFFFFFFFu, ok);
        const u64 first_off = __shfl_sync(0xFFFFFFFFu, ri.out_off, 0);
        if (ok) {
            const uintptr_t a = reinterpret_cast<uintptr_t>(job.out) + ri.out_off;
            u32 slot_img;
            if (all_ok) {
                const uintptr_t a0 = reinterpret_cast<uintptr_t>(job.out) + first_off;
                gaddr = (u64)(a0 & ~(uintptr_t)15);
                slot_img = 0;
                lo = (u32)(a - gaddr);
            } else {
                lo = (u32)(a & 15u);
                gaddr = (u64)(a - lo);
                slot_img = (u32)r * kShortSlot;
            }
            const u32 n3 = n % 3u;
#pragma unroll
            for (int k = 0; k < 3; k++) {
                // forward frame k; the reverse frame whose first codon starts k nucleotides from the record's end
                const u32 i = (strand == 0 || job.alternative) ? (u32)k : (n3 + 3u - (u32)k) % 3u;
                const u32 f = 3u * strand + i;
                flen[k] = f == 0 ? flen6[0] : f == 1 ? flen6[1] : f == 2 ? flen6[2] : f == 3 ? flen6[3] : f == 4 ? flen6[4] : flen6[5];
                const u32 brel = f == 0 ? body6[0] : f == 1 ? body6[1] : f == 2 ? body6[2] : f == 3 ? body6[3] : f == 4 ? body6[4] : body6[5];
                fdst[k] = image_addr + slot_img + lo + brel;
            }
            const u64 l = job.loc[s];
            loc_tile = (u32)(l >> 16);
            loc_pos = (u32)(l & 0xFFFFu);
            nsA = job.tile_info[loc_tile].nsym;
            cross = loc_pos + n + 2u > nsA;  // (k_size made sure that the rest lies in the next slot)
            if (strand == 0) {
                ShortMeta m;
#pragma unroll
                for (int f = 0; f < 6; f++) m.img[f] = slot_img + lo + body6[f] - (ri.H + 3u);
                m.H = ri.H;
                const u64 hinf = s >= 1 ? job.hinfo[s] : 0;
                m.sp = (u32)(hinf >> 32);
                m.hoff = s >= 1 ? job.hdr_off[s] : 0;
                m.g0 = gaddr;
                m.lo = slot_img + lo;
                m.hi = slot_img + lo + rec_bytes;
                m.ok = 1;
                m.pad = 0;
                ws.meta[r] = m;
            }
        } else if (strand == 0) {
            ws.meta[r].ok = 0;
        }

        // ---- the header lines (writer.go:120-131): lane (r, half) fetches 32 bytes of record r's header
        // (aligned words + a byte funnel) and writes them where they go in "id_N rest\n": the bytes up to the
        // first space as they are, the others two further on; all 16 records at once ----
        if (ok) {
            const u32 H = ri.H;
            const u64 hinf = s >= 1 ? job.hinfo[s] : 0;
            const u32 sp = (u32)(hinf >> 32);
            uint8_t *ln = ws.lines[r] + 4;
            if (H > 32u * (u32)strand) {
                const u64 hoff = s >= 1 ? job.hdr_off[s] : 0;
                const u64 src = hoff + 32u * (u32)strand;  // first byte wanted
                const u32 *wp = reinterpret_cast<const u32 *>(job.in + (src & ~3ull));
                const u32 sel = 0x3210u + 0x1111u * (u32)(src & 3u);
                const u32 nb = min(H - 32u * (u32)strand, 32u);  // header bytes of this half
                u32 prev = wp[0];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    if (4u * (u32)j < nb) {
                        // (the word behind is only read while it starts inside the input, rounded up to whole words)
                        const bool more = (src & ~3ull) + 4u * (j + 1) < ((job.n + 3) & ~3ull);
                        const u32 nxt = more ? wp[j + 1] : 0u;
                        const u32 v = prmt(prev, nxt, sel);
                        prev = nxt;
#pragma unroll
                        for (int b = 0; b < 4; b++) {
                            const u32 i = 32u * (u32)strand + 4u * (u32)j + (u32)b;  // index in the header line
                            if (i < H) ln[i < sp ? i : i + 2u] = (uint8_t)(v >> (8 * b));
                        }
                    }
                }
            }
            if (strand == 0) {
                ln[sp] = '_';
                ln[sp + 1] = '0';  // (the frame digit goes in per frame)
                ln[H + 2] = '\n';
            }
        }

        // ---- the strand's symbols, one per byte, in the order the strand reads them ----
        if (ok && (flen[0] | flen[1] | flen[2])) {
            const u32 *slot = job.sym + (u64)loc_tile * kSlotWords;
            u32 c[49];
            if (strand == 0) {
                sym_words_up<false>(slot, loc_pos + kShortBias, 0, c);
                // the record goes on in the next slot: its symbols from index nsA on come from there
                if (cross) sym_words_up<true>(slot + kSlotWords, loc_pos + kShortBias - nsA, nsA - loc_pos, c);
            } else {
                const u32 top = loc_pos + n - 1u;  // the record's last nucleotide
                sym_words_down<false>(slot, top + kShortBias, 0, c);
                // its top symbols (top down to nsA) lie in the next slot
                if (cross && top >= nsA) sym_words_down<true>(slot + kSlotWords, top + kShortBias - nsA, top - nsA + 1u, c);
            }
            // (only the words that hold the frame's residues: a word spills at most three bytes into the
            // header lines on either side, and those are laid over the image afterwards)
            if (flen[0]) emit_frame<0, true>(c, lut_addr, fdst[0], true, ((fdst[0] & 3u) + flen[0] + 3u) >> 2);
            if (flen[1]) emit_frame<1, true>(c, lut_addr, fdst[1], true, ((fdst[1] & 3u) + flen[1] + 3u) >> 2);
            if (flen[2]) emit_frame<2, true>(c, lut_addr, fdst[2], true, ((fdst[2] & 3u) + flen[2] + 3u) >> 2);
        }
        __syncwarp();
        // the final newline of a line shorter than 60 residues (writer.go:164-166)
        if (ok) {
#pragma unroll
            for (int k = 0; k < 3; k++)
                if (flen[k]) smem_raw[fdst[k] - smem_base + flen[k]] = '\n';
        }
        __syncwarp();

        // ---- the header lines into the image: one lane per (record, frame), word by word ----
        for (u32 p = lane; p < (u32)kShortRecs * 6u; p += 32) {
            const u32 q = p / 6u, f = p - 6u * q;
            const ShortMeta &m = ws.meta[q];
            if (!m.ok || !((job.frame_mask >> f) & 1u)) continue;
            const u32 L = m.H + 3u;
            const u32 d = m.img[f];                  // image offset of line byte 0
            const u32 a = d & 3u;
            const u32 nwords = (a + L + 3u) >> 2;    // aligned image words the line touches (<= 18)
            // image word k holds line bytes 4k - a .. 4k - a + 3 = lines[q][4 + 4k - a ..]
            const u32 *lw = reinterpret_cast<const u32 *>(ws.lines[q]) + (a == 0 ? 1 : 0);
            const u32 sel = 0x3210u + 0x1111u * ((4u - a) & 3u);
            uint8_t *dw = ws.image + (d - a);
            u32 prev = lw[0];
#pragma unroll
            for (int k = 0; k < 18; k++) {
                if ((u32)k < nwords) {
                    const u32 nxt = lw[k + 1];
                    const u32 v = prmt(prev, nxt, sel);
                    prev = nxt;
                    const int blo = 4 * k - (int)a;  // line bytes blo .. blo + 3
                    if (blo >= 0 && blo + 4 <= (int)L) {
                        *reinterpret_cast<u32 *>(dw + 4 * k) = v;
                    } else {  // the line's ragged ends: the other bytes of the word are residues
#pragma unroll
                        for (int b = 0; b < 4; b++)
                            if (blo + b >= 0 && blo + b < (int)L) dw[4 * k + b] = (uint8_t)(v >> (8 * b));
                    }
                }
            }
            ws.image[d + m.sp + 1u] = (uint8_t)('1' + f);  // (after the word that holds it: same lane, program order)
        }

        // ---- out: one bulk store for the group (or one per record), plus the ragged ends ----
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (all_ok) {
            if (lane == 0) {
                uint8_t *gp = reinterpret_cast<uint8_t *>(ws.meta[0].g0);  // global address of image byte 0
                const u32 b = ws.meta[0].lo, e = ws.meta[kShortRecs - 1].hi;
                const u32 lo16 = (b + 15u) & ~15u, hi16 = e & ~15u;
                if (hi16 > lo16) {
                    bulk_store(gp + lo16, ws.image + lo16, hi16 - lo16);
                    store_head(ws.image, gp, b, lo16);
                    store_tail(ws.image, gp, hi16, e);
                } else {
                    for (u32 x = b; x < e; x++) gp[x] = ws.image[x];
                }
            }
        } else if (lane < kShortRecs && ws.meta[lane].ok) {
            const ShortMeta &m = ws.meta[lane];
            const u32 slot_img = (u32)lane * kShortSlot;
            uint8_t *gp = reinterpret_cast<uint8_t *>(m.g0) - slot_img;  // global address of image byte 0
            const u32 lo16 = (m.lo + 15u) & ~15u, hi16 = m.hi & ~15u;
            if (hi16 > lo16) {
                bulk_store(gp + lo16, ws.image + lo16, hi16 - lo16);
                store_head(ws.image, gp, m.lo, lo16);
                store_tail(ws.image, gp, hi16, m.hi);
            } else if (m.hi <= lo16) {
                store_edge(ws.image, gp, m.lo, m.hi);
            } else {
                store_edge(ws.image, gp, m.lo, lo16);
                store_edge(ws.image, gp, lo16, m.hi);
            }
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the image may be reused
        __syncwarp();
    }
}

cudaError_t init_short() {
    return cudaFuncSetAttribute(k_short, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ShortShared));
}

cudaError_t launch_short(const Job &job, int sm_count, cudaStream_t stream) {
    k_short<<<sm_count * 4, kShortWarps * 32, sizeof(ShortShared), stream>>>(job);
    return cudaGetLastError();
}

}  // namespace gts
