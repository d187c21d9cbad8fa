// gts_device.h -- data layout in HBM shared by the host API (gts_api.cu) and the kernels.
// See DESIGN.md "Data layout" for the narrative.
//
// One device pass turns a FASTA buffer into protein FASTA with six kernels:
//
//   k_parse      raw bytes -> per-tile 4-bit symbol slots + header events   (transeq.go:183-216,
//                                                                            encodedSequence.go:17-43)
//   k_tile_scan  per-tile symbol / header counts -> stream offsets of the tiles
//   k_records    header events + tile offsets -> record table (header offset, stream base)
//   k_size       record table -> frame lengths (trim), output offsets       (writer.go:85-116,159-169)
//   k_lines      symbols -> residues of all requested frames, laid out as
//                60-column lines at their final output offsets              (writer.go:57-157,
//                                                                            encodedSequence.go:71-97)
//   k_headers    record table + raw header bytes -> ">id_N rest" lines      (writer.go:120-131)
//
// Symbol stream: one 4-bit symbol per nucleotide in input order, codes N=0 A=1 C=2 T/U=3 G=4
// (transeq.go:14-21).  Two 0 symbols ("pads") are inserted in front of every record and two
// more follow the last one, so that EVERY residue of EVERY frame is the table entry of three
// consecutive symbols (e-2, e-1, e): the pads play the role of the reference's tail rules
// (two-letter codon / 'X', writer.go:102-112) on both strands.
//
// The stream is stored as one fixed-size slot per parse tile (kParseTile raw bytes): tile t's
// symbols are compacted inside its slot only, so k_parse has no dependency between tiles; the
// stream index of a symbol is T0[t] + its index in the slot, with T0 the exclusive scan of the
// tiles' symbol counts (k_tile_scan).  Word w of a slot holds its symbols 8w..8w+7; byte i of the
// word = symbol 8w+i | symbol 8w+i+4 << 4.
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
constexpr int kSlotSyms = kParseTile + 64;  // symbols a slot can hold: the tile's bytes + 2 leading + 2 end pads
constexpr int kSlotWords = kSlotSyms / 8;   // 1032 stream words per slot

// ---- k_tile_scan / k_size geometry ----
constexpr int kScanThreads = 256;
#ifndef GTS_SCAN_PER_THREAD
#define GTS_SCAN_PER_THREAD 4
#endif
constexpr int kScanPerThread = GTS_SCAN_PER_THREAD;
constexpr int kScanTile = kScanThreads * kScanPerThread;  // parse tiles per scan tile
// ---- k_short: records small enough for the warp-per-16-records emit (gts_short.cu) ----
constexpr int kShortMaxN = 180;    // nucleotides: every frame is a single output line
constexpr int kShortMaxH = 64;     // header line bytes
constexpr int kShortMaxOut = 672;  // output bytes of the record (all requested frames)

// ---- k_records: lanes per header line (a power of two; each lane reads 16 bytes per step) ----
#ifndef GTS_REC_LANES
#define GTS_REC_LANES 4
#endif
constexpr int kRecLanes = GTS_REC_LANES;

#ifndef GTS_SIZE_THREADS
#define GTS_SIZE_THREADS 512
#endif
constexpr int kSizeThreads = GTS_SIZE_THREADS;
#ifndef GTS_SIZE_PER_THREAD
#define GTS_SIZE_PER_THREAD 1
#endif
constexpr int kSizePerThread = GTS_SIZE_PER_THREAD;
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
    u64 cut;          // with kFlagLongLine: the reference stops reading at this offset
    u64 tail;         // the stream's last two symbols (end pads excluded): last | second-to-last << 4
    u64 n_warn;       // invalid nucleotides seen (WARNING lines of encodedSequence.go:39), exact even on overflow
    u64 n_short;      // records handed to k_short (RecInfo.emitted == 2)
    u64 aux;          // piece scan: bytes of the header line the buffer starts with (CR dropped), ~0 if it starts with none
};
constexpr u64 kFlagRecOverflow = 1;  // record table too small (n_headers is still exact)
constexpr u64 kFlagOutOverflow = 2;  // output buffer too small (out_total is still exact)
constexpr u64 kFlagLongLine = 4;     // a line of >= max_line bytes: the pass must be redone on [0, cut)
constexpr u64 kFlagWarnOverflow = 8; // warning list too small (n_warn is still exact)

// One invalid nucleotide (anything but ACGTUN in either case, encodedSequence.go:25-41), resolved
// to its record by k_warn: what the reference's WARNING line needs.
struct WarnRec {   // 32 bytes
    u64 key;       // stream index of the symbol: input order
    u64 hdr_off;   // raw offset of the record's header line (0 with H == 0: no header line)
    u64 pos;       // index of the nucleotide in its record
    u32 H;         // header line bytes
    u32 byte;      // the offending byte
};

struct TileInfo {  // 16 bytes per parse tile, written by k_parse
    uint16_t nsym;      // symbols in the tile's slot (pads included)
    uint16_t nhdr;      // header lines starting in the tile
    uint8_t tail;       // its last two symbols: last | second-to-last << 4
    uint8_t has_nl;     // the tile holds a '\n' (filled in only when the input can hold an over-long line)
    uint16_t first_nl;  // offset of its first / last '\n' in the tile
    uint16_t last_nl;
    uint16_t pad[3];
};

struct NlPart {  // newline summary of a run of tiles, for the over-long line rule (transeq.go:27,186)
    i64 first;   // position of its first '\n', -1 = none
    i64 last;    // position of its last '\n', -1 = none
    i64 cut;     // start of the first line of >= max_line bytes lying between two of its newlines, -1 = none
    i64 pad;
};

struct TilePrefix {  // 16 bytes per parse tile, written by k_tile_scan
    u64 T0;     // stream index of the slot's first symbol
    u32 R0;     // header lines before the tile == record slot open at its first byte
    u32 carry;  // the two symbols before the tile: last | second-to-last << 4
};

struct HeaderEvent {  // 16 bytes per header line, appended by k_parse in no particular order
    u64 hdr_off;   // raw offset of the '>'
    u32 tile;
    uint16_t k;    // index of the header line among those starting in its tile
    uint16_t pos;  // slot index of the record's first nucleotide (its pads sit at pos-2, pos-1)
};

struct alignas(16) RecInfo {  // 32 bytes per record slot, written by k_size, read by k_lines / k_headers
    u64 out_off;  // offset of the record's first output byte
    u64 dbase;    // stream index of the record's first nucleotide
    u64 n;        // nucleotides
    u32 H;        // header line bytes (CR dropped)
    u32 emitted;  // 0: the record produces no output (empty slot 0); 1: written by k_lines + k_headers; 2: by k_short
};

struct alignas(16) FrameTab {  // 80 bytes per record slot, written by k_size, read by k_lines
    u64 body[6];   // output offset of the first residue of each requested frame
    u32 len[6];    // residues of the frame (after --trim); 0 for frames not requested / records not emitted
    u32 pad[2];
};

struct Lut {  // 125 used entries: forward residue | reverse-strand residue << 8
    uint16_t v[128];
};

struct LutBytes {  // k_lines' tables, one residue per byte (125 used entries each)
    u32 f[32];     // forward residue of the codon s0 + 5 s1 + 25 s2
    u32 r[32];     // reverse-strand residue of the codon read downwards: index s2 + 5 s1 + 25 s0
};

// A piece of one record split over several GPUs (gts_piece_*).  The piece's symbols keep their index
// in the whole record's stream: t0_base is the stream index of the buffer's first symbol, and one
// record slot is given the whole record's geometry.
struct PieceParams {
    u32 piece;         // 0 = an ordinary pass
    u32 no_end_pads;   // the stream goes on after this buffer: no end pads
    u32 has_carry;     // the two symbols in front of the buffer are `carry` (they replace slot 0's pads)
    u32 carry;         // last | second-to-last << 4
    u32 no_headers;    // the record's header lines are written by the first piece only
    u32 ov_H;          // the record's header line bytes
    u32 ov_last;       // this piece ends the record (k_piece_fix adds the end pads to the reused slots)
    u32 pad;
    u64 t0_base;
    u64 win_lo;        // windows ending below this stream index belong to the piece before
    u64 ov_slot;       // the record's slot in this pass; every other slot is silent
    u64 ov_n;          // the whole record's nucleotides
    u64 ov_dbase;      // stream index of its first nucleotide
    uint8_t head[192]; // the 192 symbols behind the buffer (the record goes on in the next piece)
};

// What the scan of a piece reports (256 bytes: the unit of the all-gather between the ranks).
struct PieceInfoDev {
    u64 nucleotides;
    u64 headers;       // header lines seen: 1 for the first piece, 0 for the others
    u64 header_len;    // bytes of the header line the buffer starts with (CR dropped), ~0 = it starts with none
    u32 tail;          // codes of its last | second-to-last << 4 nucleotide
    u32 pad;
    uint8_t head[192]; // codes of its first 192 nucleotides (0 beyond its end)
    uint8_t fill[32];
};
static_assert(sizeof(PieceInfoDev) == 256, "all-gather unit");

struct Job {  // kernel argument block (by value)
    const uint8_t *in;  // raw FASTA, 16-byte aligned
    u64 n;
    uint8_t *out;
    u64 out_cap;

    u32 *sym;              // [ntiles * kSlotWords] symbol slots
    TileInfo *tile_info;   // [ntiles]
    TilePrefix *tile_pre;  // [ntiles]
    HeaderEvent *events;   // [rec_cap]

    // record table, slots 0..R (+1 sentinel)
    u64 *hdr_off;   // raw offset of the header line's '>'
    u64 *dbase;     // stream index of the record's first nucleotide (its pads sit at -2, -1)
    u64 *loc;       // tile << 16 | slot index of the record's first nucleotide
    u64 *hinfo;     // header line bytes (CR dropped) | index of its first space << 32
    RecInfo *rec;   // per record summary (+1 sentinel: dbase = end of stream, out_off = total)
    u32 *trim_len;  // [slot*6 + frame] residues kept after --trim (only with trim)
    FrameTab *ftab; // per record: where each frame's residues go and how many there are
    u64 rec_cap;    // entries in each of the arrays above

    u64 *t_status;  // [tile-scan tiles * 4] chained scan of the parse tiles' counts
    NlPart *scan_nl;// [tile-scan tiles] newline summaries (only used when n >= max_line)
    u64 max_line;   // bufio.Scanner token limit of the reference: 100 MiB (transeq.go:27)
    u64 *s_status;  // [size tiles * 4] chained scan of output sizes
    u32 *counters;  // [0] tile-scan ticket, [1] size ticket, [2] header events appended, [3] k_lines tile ticket,
                    // [4] invalid nucleotides appended to `warn`
    u64 *warn;      // [warn_cap] tile << 32 | slot index << 8 | byte, appended by k_parse (nullptr = off)
    WarnRec *wres;  // [warn_cap] resolved by k_warn
    u32 warn_cap;
    u32 pad0;
    Totals *totals;
    u64 *dbg;       // optional phase timing sums (GTS_PHASE_TIMING builds), else nullptr

    u32 ntiles;      // parse tiles
    u32 frame_mask;  // bit k = frame index k requested (0..2 forward, 3..5 reverse)
    u32 alternative;
    u32 trim;
    u32 more_follows;  // this buffer is not the end of the stream: an empty slot 0 is never emitted
    u32 short_on;      // short records go to k_short (emitted == 2)
    u32 pad2;

    // ---- a piece of ONE record split over several passes / GPUs (gts_piece_*) ----
    // The piece's description: by value (pv, host-driven calls) or in device memory (pp, written by
    // k_piece_plan from the all-gathered scan results: no host round trip between scan and translate).
    PieceParams pv;
    const PieceParams *pp;
    PieceInfoDev *info_out;  // scan pass: receives the piece's counts and boundary codes
    u32 ov_trim;       // --trim on a split record: the record's trimmed frame lengths are given (ov_len)
    u32 pad1;
    u32 ov_len[6];
    // k_piece_trim: one step of the trim walk over the pieces (writer.go:141-163 on a split record)
    u64 pt_len[6];     // residues of each frame not yet known to be trimmed away
    u64 pt_before;     // nucleotides of the record in front of the piece
    u64 *pt_out;       // [12] the lengths after this piece's part of the walk, then "resolved" flags
    Lut lut;
    LutBytes lutb;
};

}  // namespace gts
