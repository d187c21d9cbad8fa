// gts_device.h -- data layout in HBM shared by the host API (gts_api.cu) and the kernels
// (gts_kernels.cu).  See DESIGN.md "Data layout" for the narrative.
//
// One device pass turns a FASTA buffer into protein FASTA with four kernels:
//
//   k_parse      raw bytes -> dense 4-bit symbol stream + record table      (transeq.go:183-216,
//                                                                            encodedSequence.go:17-43)
//   k_size       record table -> frame lengths (trim), output offsets       (writer.go:85-116,159-169)
//   k_translate  symbol stream -> residues of all requested frames, laid
//                out as 60-column lines at their final output offsets       (writer.go:57-157,
//                                                                            encodedSequence.go:71-97)
//   k_headers    record table + raw header bytes -> ">id_N rest" lines      (writer.go:120-131)
//
// Symbol stream: one 4-bit symbol per nucleotide in input order, codes N=0 A=1 C=2 T/U=3 G=4
// (transeq.go:14-21).  Two 0 symbols ("pads") are inserted in front of every record and two
// more follow the last one, so that EVERY residue of EVERY frame is the table entry of three
// consecutive symbols (e-2, e-1, e): the pads play the role of the reference's tail rules
// (two-letter codon / 'X', writer.go:102-112) on both strands.  Word w of the stream holds
// symbols 8w..8w+7; byte i of the word = symbol 8w+i | symbol 8w+i+4 << 4.
#pragma once
#include <cstdint>

namespace gts {

typedef unsigned long long u64;
typedef long long i64;
typedef unsigned int u32;

// ---- k_parse geometry: one CTA per tile of raw bytes ----
constexpr int kParseThreads = 256;
constexpr int kParseBytesPerThread = 32;
constexpr int kParseTile = kParseThreads * kParseBytesPerThread;  // 8192 raw bytes

// ---- k_translate geometry: one warp per sub-tile of symbols, kTransWarps sub-tiles per CTA ----
constexpr int kTransSymPerThread = 48;
constexpr int kTransSub = 32 * kTransSymPerThread;        // 1536 symbols per warp
constexpr int kTransSubWords = kTransSub / 8;             // 192 stream words
constexpr int kTransWarps = 8;                            // translating warps per CTA (+ 1 builder warp)
constexpr int kTransTile = kTransWarps * kTransSub;       // 12288 symbols per CTA
constexpr int kTransTileWords = kTransTile / 8;
constexpr int kTransBlock = (kTransWarps + 1) * 32;
constexpr int kPhaseLen = kTransSub / 3;                  // residues per phase array (per warp)
constexpr int kPhaseGuard = 16;
constexpr int kPhaseStride = kPhaseLen + 2 * kPhaseGuard; // bytes per phase array in smem
constexpr int kRecBatch = 4;                              // record pieces per sub-tile and format batch
constexpr int kRunsPerBatch = kRecBatch * 6;
// Output image of one sub-tile batch in shared memory: the residue lines a sub-tile writes (at most
// 2 * kTransSub * 61 / 60 bytes + one final newline per run) plus up to 34 bytes of 16-byte
// alignment padding per run.
constexpr int kImageBytes = (2 * kTransSub * 61 / 60 + kRunsPerBatch * 35 + 128 + 15) / 16 * 16;

// ---- k_size geometry ----
constexpr int kSizeThreads = 256;
constexpr int kSizePerThread = 4;
constexpr int kSizeTile = kSizeThreads * kSizePerThread;  // record slots per scan tile

// Chained-scan status: 4 words per tile {agg.a, agg.b, inc.a, inc.b}; bit 63 of the .a words
// is the "published" flag (values stay below 2^63).  agg = this tile only, inc = tiles 0..t.
constexpr u64 kValid = 1ull << 63;

struct Totals {  // device-resident result block, copied to pinned host memory after a pass
    u64 n_headers;    // R: number of header lines; record slots are 0..R (slot 0 = bytes before
                      // the first header line, emitted only if it has nucleotides or R == 0)
    u64 total_sym;    // symbols in the stream, the two end pads not counted
    u64 out_total;    // exact size of the protein FASTA
    u64 flags;        // kFlag*
    u64 records;      // records emitted
    u64 nucleotides;  // nucleotides in emitted records
    u64 pad[2];
};
constexpr u64 kFlagRecOverflow = 1;  // record table too small (n_headers is still exact)
constexpr u64 kFlagOutOverflow = 2;  // output buffer too small (out_total is still exact)

struct RecInfo {  // 32 bytes per record slot, written by k_size, read by k_translate / k_headers
    u64 out_off;  // offset of the record's first output byte
    u64 dbase;    // stream index of the record's first nucleotide
    u64 n;        // nucleotides
    u32 H;        // header line bytes (CR dropped)
    u32 emitted;  // 0: the record produces no output (empty slot 0)
};

struct Lut {  // 125 used entries: forward residue | reverse-strand residue << 8
    uint16_t v[128];
};

struct Job {  // kernel argument block (by value)
    const uint8_t *in;  // raw FASTA, 16-byte aligned
    u64 n;
    uint8_t *out;
    u64 out_cap;

    u32 *sym;  // packed symbol stream

    // two-level chained scan of (symbols, header lines) over the parse tiles, groups of 32 tiles
    u32 *p_agg;     // [ntiles_parse] bit 31 valid | header lines << 14 | symbols      (one tile)
    u64 *p_gagg;    // [groups] bit 63 valid | header lines << 32 | symbols            (one group)
    u64 *p_ginc;    // [groups * 2] {bit 63 valid | symbols, header lines} of groups 0..g
    u32 *p_pend;    // [ntiles_parse] symbols of the last incomplete stream word, bit 31 = valid
    u32 *t_hint;    // [translate sub-tiles] the record slot owning the sub-tile's first symbol

    // record table, structure of arrays, slots 0..R (+1 sentinel in dbase / out_off)
    u64 *hdr_off;   // raw offset of the header line's '>'
    u64 *hdr_end;   // raw offset one past the header line's last byte (CR dropped)
    u64 *dbase;     // stream index of the record's first nucleotide (its pads sit at -2, -1)
    RecInfo *rec;   // per record summary (+1 sentinel: dbase = end of stream, out_off = total)
    u32 *trim_len;  // [slot*6 + frame] residues kept after --trim (only with trim)
    u64 rec_cap;    // entries in each of the arrays above

    u64 *s_status;  // [size tiles * 4] chained scan of output sizes
    u32 *counters;  // [0] parse tile ticket, [1] size tile ticket
    Totals *totals;
    u64 *dbg;       // optional phase timing sums (GTS_PHASE_TIMING builds), else nullptr

    u32 ntiles_parse;
    u32 frame_mask;  // bit k = frame index k requested (0..2 forward, 3..5 reverse)
    u32 alternative;
    u32 trim;
    Lut lut;
};

}  // namespace gts
