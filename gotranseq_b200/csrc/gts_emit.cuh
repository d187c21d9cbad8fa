// gts_emit.cuh -- one output line of one frame from a block's symbols: the inner loop shared by
// k_lines (gts_lines.cu) and k_short (gts_short.cu).
//
// Reference behaviour reproduced (feliixx/gotranseq): the codon loop of translate3Frames,
// transeq/writer.go:97-100, with the codon array of transeq/transeq.go:30-86 as a 125-byte table.
#pragma once
#include "gts_common.cuh"

namespace gts {

// dp4a weights (1, 5, 25 for a codon's three symbols) of codon j of phase k inside word `word` of
// a 12-symbol group (word 3 = the first word of the next group); 0 = the codon has no symbol there
__host__ __device__ constexpr u32 codon_weights(int k, int j, int word) {
    u32 m = 0;
    for (int t = 0; t < 3; t++) {
        const int b = k + 3 * j + t;
        if ((b >> 2) == word) m |= (t == 0 ? 1u : t == 1 ? 5u : 25u) << (8 * (b & 3));
    }
    return m;
}

// Line i of one frame of a block run: 63 codons of phase K from the block's symbols c[] -> 64 image
// bytes (60 residues, '\n', 3 residues of the next line).
// (shared-window addresses throughout: `lut` is the table's, so that a lookup is one LDS.U8 [R]
// without an add of the window's base -- 189 of them per block -- and `dst` the line's in the image)
__device__ __forceinline__ u32 lds_u8(u32 saddr) {
    u32 v;
    asm("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void sts_u32(u32 saddr, u32 v) {
    asm volatile("st.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
}

// kPartial (k_short): only the first `nw` image words are stored -- the line is shorter than 60
// residues and what follows it in the image is not this lane's.
template <int K, bool kPartial = false>
__device__ __forceinline__ void emit_frame(const u32 (&c)[49], u32 lut, u32 dst, bool first, u32 nw = 16) {
    const u32 a = dst & 3u;
    const u32 sela = 0x3210u + 0x1111u * ((4u - a) & 7u);
    u32 z[16];
#pragma unroll
    for (int g = 0; g < 16; g++) {
        u32 r[4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            u32 addr = lut;  // table offset of the codon: lut + s0 + 5 s1 + 25 s2
#pragma unroll
            for (int w = 0; w < 4; w++) {
                constexpr int dummy = 0;
                (void)dummy;
                const u32 m = codon_weights(K, j, w);
                if (m != 0 && 3 * g + w < 49) addr = __dp4a(c[3 * g + w], m, addr);
            }
            r[j] = lds_u8(addr);
        }
        if (g < 15) {
            z[g] = prmt(prmt(r[0], r[1], 0x0040u), prmt(r[2], r[3], 0x0040u), 0x5410u);
        } else {  // residues 60..62 belong to the next line: they follow this line's newline
            z[g] = prmt(prmt(0x0Au, r[0], 0x0040u), prmt(r[1], r[2], 0x0040u), 0x5410u);
        }
    }
    // image word g holds line bytes 4g - a .. 4g - a + 3; a lane owns the words that start in its
    // line (word 0 only when the line starts on a word, or when no line precedes it)
    const u32 dw = dst & ~3u;
    if (a == 0 || first) sts_u32(dw, prmt(0u, z[0], sela));
#pragma unroll
    for (int g = 1; g < 16; g++)
        if (!kPartial || (u32)g < nw) sts_u32(dw + 4u * g, prmt(z[g - 1], z[g], sela));
}

}  // namespace gts
