// gts_codes.h -- genetic codes and option decoding shared by the host side of the library.
//
// The 21 tables are the ones the reference ships (ncbicode/code.go:341-363), written as the
// 64 amino acids of the *effective* codon map LoadTableCode returns (code.go:367-388: the
// standard map code.go:16-81 overlaid with the per-table diff, U read as T), in T,C,A,G
// codon order (index = 16*b0 + 4*b1 + b2).  They reproduce the reference as coded, which is
// the parity target (e.g. its table 11 equals the standard code and its table 23 differs
// from it only by TTA -> '*').  tests/test_abi.py checks all 21 against tests/golden/ncbi_tables.json,
// derived from the reference's ncbicode/code.go by tests/golden/make_ncbi_tables.py.
#pragma once
#include <cstdint>
#include <cstring>

namespace gts {

struct CodonTable { int id; const char *aa; };

static const CodonTable kCodonTables[] = {
    { 0, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    { 2, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSS**VVVVAAAADDEEGGGG"},
    { 3, "FFLLSSSSYY**CCWWTTTTPPPPHHQQRRRRIIMMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    { 4, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    { 5, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSSSVVVVAAAADDEEGGGG"},
    { 6, "FFLLSSSSYYQQCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    { 9, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
    {10, "FFLLSSSSYY**CCCWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {11, "FFLLSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {12, "FFLLSSSSYY**CC*WLLLSPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {13, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNKKSSGGVVVVAAAADDEEGGGG"},
    {14, "FFLLSSSSYYY*CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
    {16, "FFLLSSSSYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {21, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIMMTTTTNNNKSSSSVVVVAAAADDEEGGGG"},
    {22, "FFLLSS*SYY*LCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {23, "FF*LSSSSYY**CC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {24, "FFLLSSSSYY**CCWWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSSKVVVVAAAADDEEGGGG"},
    {25, "FFLLSSSSYY**CCGWLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {26, "FFLLSSSSYY**CC*WLLLAPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {29, "FFLLSSSSYYYYCC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
    {30, "FFLLSSSSYYEECC*WLLLLPPPPHHQQRRRRIIIMTTTTNNKKSSRRVVVVAAAADDEEGGGG"},
};
static const int kNumCodonTables = sizeof(kCodonTables) / sizeof(kCodonTables[0]);

inline const char *find_codon_table(int id) {
    for (int i = 0; i < kNumCodonTables; i++)
        if (kCodonTables[i].id == id) return kCodonTables[i].aa;
    return nullptr;
}

// Nucleotide codes of the reference (transeq/transeq.go:14-21): N=0 A=1 C=2 T/U=3 G=4.
// complement (encodedSequence.go:78-93): A<->T, C<->G, N stays.
inline int complement_code(int c) { return c == 0 ? 0 : (((c - 1) ^ 2) + 1); }

// Amino acid the reference's codon array gives for codes (c0,c1,c2) (transeq.go:30-86):
//  - three real bases: the table entry ('X' instead of '*' when clean, transeq.go:54)
//  - (x, y, N): the two-letter entry, set only if the 4 codons xy? agree (transeq.go:65-83)
//  - anything else: 'X' (the array's default, transeq.go:33-35)
inline char codon_aa(const char *aa64, bool clean, int c0, int c1, int c2) {
    static const int tcag_of_code[5] = {-1, 2, 1, 0, 3};  // code -> index in T,C,A,G
    if (c0 == 0 || c1 == 0) return 'X';
    int b0 = tcag_of_code[c0], b1 = tcag_of_code[c1];
    char r;
    if (c2 == 0) {
        const char *q = aa64 + 16 * b0 + 4 * b1;
        if (!(q[0] == q[1] && q[1] == q[2] && q[2] == q[3])) return 'X';
        r = q[0];
    } else {
        r = aa64[16 * b0 + 4 * b1 + tcag_of_code[c2]];
    }
    if (clean && r == '*') return 'X';
    return r;
}

// 125-entry fused table: index c0 + 5*c1 + 25*c2; low byte = forward residue of the window
// (c0,c1,c2), high byte = residue of the same window read on the reverse strand, i.e. of the
// codon (comp c2, comp c1, comp c0).  This is what lets the kernels skip the in-place
// reverseComplement pass of encodedSequence.go:71-97.
inline void build_fused_lut(const char *aa64, bool clean, uint16_t lut[125]) {
    for (int c2 = 0; c2 < 5; c2++)
        for (int c1 = 0; c1 < 5; c1++)
            for (int c0 = 0; c0 < 5; c0++) {
                unsigned f = (unsigned char)codon_aa(aa64, clean, c0, c1, c2);
                unsigned r = (unsigned char)codon_aa(aa64, clean, complement_code(c2),
                                                     complement_code(c1), complement_code(c0));
                lut[c0 + 5 * c1 + 25 * c2] = (uint16_t)(f | (r << 8));
            }
}

// computeFrames (transeq.go:88-110): bit k of mask = frame index k requested
// (0,1,2 = frames 1,2,3; 3,4,5 = frames -1,-2,-3).
inline bool frame_mask(const char *name, uint32_t *mask, bool *reverse) {
    static const struct { const char *n; uint32_t m; } kMap[] = {
        {"1", 0x01}, {"2", 0x02}, {"3", 0x04}, {"F", 0x07}, {"-1", 0x08},
        {"-2", 0x10}, {"-3", 0x20}, {"R", 0x38}, {"6", 0x3f},
    };
    if (name)
        for (auto &e : kMap)
            if (std::strcmp(e.n, name) == 0) {
                *mask = e.m;
                *reverse = (e.m & 0x38) != 0;
                return true;
            }
    return false;
}

}  // namespace gts
