// gts_device.h -- data layout in HBM shared by the host API (gts_api.cu) and the kernels
// (gts_kernels.cu).  See DESIGN.md "Data layout" for the narrative.
#pragma once
#include <cstdint>

namespace gts {

typedef unsigned long long u64;
typedef long long i64;

// One tile = kTileBytes consecutive bytes of the raw FASTA buffer, processed by one CTA of
// kThreads threads, kBytesPerThread consecutive bytes per thread.
constexpr int kTileBytes = 8192;
constexpr int kThreads = 256;
constexpr int kBytesPerThread = 32;
constexpr int kWarps = kThreads / 32;
static_assert(kThreads * kBytesPerThread == kTileBytes, "tile geometry");

// Records per CTA pass of the sizing kernel.
constexpr int kSizeTile = 1024;

// Chained-scan status words.  Top two bits: 0 = not published, 1 = aggregate of this tile
// only, 2 = inclusive prefix over tiles 0..t.  Low 62 bits: the value.
constexpr u64 kFlagAgg = 1ull << 62;
constexpr u64 kFlagInc = 2ull << 62;
constexpr u64 kFlagMask = 3ull << 62;
constexpr u64 kValMask = ~kFlagMask;

struct TileStatus {  // 32 B, zeroed before every pass
    u64 w0;          // phase 1: (position + 1) of the last '\n' seen so far (max), 0 = none
    u64 w1;          // phase 2: dense symbols (nucleotides + 2 pads per record start) (sum)
    u64 w2;          // phase 2: header lines (sum)
    u64 pad;
};

// Line state at the first byte of a tile.
enum : uint32_t { kStateLineStart = 0, kStateInSeq = 1, kStateInHeader = 2 };

struct TilePrefix {  // 16 B, written by the parse kernel, read by the emit kernel
    u64 dense_before;     // dense symbols in tiles < t (tile 0 owns two leading pads)
    uint32_t rec_before;  // header lines in tiles < t  (== slot of the record open at tile start)
    uint32_t state;       // kState*
};

struct Totals {  // device-resident result block, copied to pinned host memory after a pass
    u64 n_headers;    // R: number of header lines; record slots are 0..R (slot 0 = bytes before
                      // the first header line)
    u64 total_dense;  // dense symbols in the whole buffer
    u64 out_total;    // exact size of the protein FASTA
    u64 nucleotides;  // nucleotides of emitted records
    u64 records;      // records emitted (slot 0 counted only when present)
    u64 flags;        // bit 0: record table too small, bit 1: output buffer too small
    u64 pad[2];
};
constexpr u64 kFlagRecOverflow = 1;
constexpr u64 kFlagOutOverflow = 2;

struct Lut {  // 125 used entries: forward residue | reverse-strand residue << 8
    uint16_t v[128];
};

struct Job {  // kernel argument block (by value)
    const uint8_t *in;
    u64 n;
    uint8_t *out;
    u64 out_cap;

    TileStatus *tstatus;  // [ntiles]
    TilePrefix *tprefix;  // [ntiles]
    unsigned int *counters;  // [0] parse tile ticket, [1] size tile ticket
    u64 *sstatus;            // [size tiles] chained-scan words of the sizing kernel

    // record table, structure of arrays, slots 0..R (+1 sentinel in dbase)
    u64 *hdr_off;   // raw offset of the header line's '>'
    u64 *hdr_nl;    // raw offset of the '\n' ending the header line (or n at EOF);
                    // bit 63 = a '\r' before it was dropped (bufio.ScanLines dropCR)
    u64 *dbase;     // dense index of the record's first nucleotide
    u64 *out_off;   // offset of the record's first output byte
    u64 *trim_len;  // [slot*6 + frame] residues kept after --trim (only with trim)
    u64 rec_cap;    // usable slots

    Totals *totals;
    uint32_t ntiles;
    uint32_t frame_mask;   // bit k = frame index k requested (0..2 forward, 3..5 reverse)
    uint32_t alternative;
    uint32_t trim;
};

}  // namespace gts
