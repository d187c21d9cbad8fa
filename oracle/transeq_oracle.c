/*
 * transeq_oracle.c -- TEST INFRASTRUCTURE ONLY (see transeq_oracle.h).
 *
 * Plain-C restatement of the reference's CPU algorithm for the translation hot path:
 * same data structures (263 173-byte codon array, [4B LE size][header][codes] records,
 * in-place reverse complement, per-residue append with lazy 60-column wrapping, 1 MiB
 * worker buffers) and the same threading shape (one reader thread feeding a bounded
 * queue of 100 records to NumWorker translate threads), so that it can double as the
 * timed CPU baseline.  Every function cites the reference file:line it follows.
 *
 * Parity: PINNED by the reference's 29 golden cases (tests/test_oracle_goldens.py).
 */
#define _GNU_SOURCE
#include "transeq_oracle.h"

#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* ---- constants: transeq/transeq.go:14-28, transeq/writer.go:10-21 ---- */
enum { N_CODE = 0, A_CODE = 1, C_CODE = 2, T_CODE = 3, G_CODE = 4 };
#define MB (1u << 20)
#define MAX_BUFFER_SIZE (1u * MB)
#define MAX_LINE_SIZE 60
#define STOP_AA '*'
#define UNKNOWN_AA 'X'
static const char SUFFIXES[] = "123456";
/* bufio.Scanner max token size: transeq/transeq.go:27,186 (overridable for tests) */
static size_t g_max_seq_length = 100u * MB;

void to_set_max_line_length(size_t n) { g_max_seq_length = n; }

/* ---- genetic codes: ncbicode/code.go:16-337 ----
 * "standard" is the 64-codon map (code.go:16-81) written in T,C,A,G order; every other table
 * is the list of overrides the reference applies on top of it (code.go:83-313), spelled as
 * the reference spells them (some with U), 4 chars per entry: codon + amino acid.  Tables 11
 * and 23 are full 64-entry maps in the reference; only entries that differ from standard are
 * listed here (11: none; 23: TTA->*). */
static const char STANDARD[65] =
    "FFLLSSSSYY**CC*W" "LLLLPPPPHHQQRRRR" "IIIMTTTTNNKKSSRR" "VVVVAAAADDEEGGGG";

typedef struct { int id; const char *diff; } table_diff;
static const table_diff DIFFS[] = {
    {2, "AGA*AGG*AUAMUGAW"},
    {3, "AUAMCUUTCUCTCUATCUGTUGAW"},
    {4, "UGAW"},
    {5, "AGASAGGSAUAMUGAW"},
    {6, "UAAQUAGQ"},
    {9, "AAANAGASAGGSUGAW"},
    {10, "UGAC"},
    {11, ""},
    {12, "CUGS"},
    {13, "AGAGAGGGAUAMUGAW"},
    {14, "AAANAGASAGGSUAAYUGAW"},
    {16, "TAGL"},
    {21, "TGAWATAMAGASAGGSAAAN"},
    {22, "TCA*TAGL"},
    {23, "TTA*"},
    {24, "AGASAGGKUGAW"},
    {25, "UGAG"},
    {26, "CUGA"},
    {29, "UAAYUAGY"},
    {30, "UAAEUAGE"},
};

static int base_index(char c) {
    switch (c) {
    case 'T': case 'U': return 0;
    case 'C': return 1;
    case 'A': return 2;
    case 'G': return 3;
    }
    return -1;
}

/* ncbicode/code.go:367-388 LoadTableCode */
int to_load_table_code(int table, char aa[64], char *err, size_t errcap) {
    memcpy(aa, STANDARD, 64);
    if (table == 0) return 0;
    for (size_t i = 0; i < sizeof(DIFFS) / sizeof(DIFFS[0]); i++) {
        if (DIFFS[i].id != table) continue;
        for (const char *p = DIFFS[i].diff; *p; p += 4) {
            /* strings.Replace(codon, "U", "T", -1): base_index folds U onto T */
            int ix = 16 * base_index(p[0]) + 4 * base_index(p[1]) + base_index(p[2]);
            aa[ix] = p[3];
        }
        return 0;
    }
    if (err) snprintf(err, errcap, "invalid table code: %d", table);
    return -1;
}

/* transeq/transeq.go:30-86 createCodeArray */
int to_create_code_array(int table, int clean, uint8_t *codes, char *err, size_t errcap) {
    static const uint8_t letter_code[4] = {T_CODE, C_CODE, A_CODE, G_CODE}; /* T,C,A,G */
    char aa[64];
    memset(codes, UNKNOWN_AA, TO_ARRAY_CODE_SIZE);
    if (to_load_table_code(table, aa, err, errcap) != 0) return -1;

    for (int b0 = 0; b0 < 4; b0++)
        for (int b1 = 0; b1 < 4; b1++) {
            /* twoLetterMap: the set of distinct amino acids among the 4 codons b0 b1 x */
            int all_same = 1;
            for (int b2 = 0; b2 < 4; b2++) {
                char a = aa[16 * b0 + 4 * b1 + b2];
                if (!(clean && a == STOP_AA)) {
                    uint32_t index = (uint32_t)letter_code[b0] | (uint32_t)letter_code[b1] << 8 |
                                     (uint32_t)letter_code[b2] << 16;
                    codes[index] = (uint8_t)a;
                }
                if (a != aa[16 * b0 + 4 * b1]) all_same = 0;
            }
            char a0 = aa[16 * b0 + 4 * b1];
            if (all_same && !(clean && a0 == STOP_AA)) {
                uint32_t index = (uint32_t)letter_code[b0] | (uint32_t)letter_code[b1] << 8;
                codes[index] = (uint8_t)a0;
            }
        }
    return 0;
}

/* transeq/transeq.go:88-110 computeFrames */
int to_compute_frames(const char *name, int frames[6], int *reverse, char *err, size_t errcap) {
    static const struct { const char *name; int f[6]; int rev; } MAP[] = {
        {"1", {1, 0, 0, 0, 0, 0}, 0}, {"2", {0, 1, 0, 0, 0, 0}, 0}, {"3", {0, 0, 1, 0, 0, 0}, 0},
        {"F", {1, 1, 1, 0, 0, 0}, 0}, {"-1", {0, 0, 0, 1, 0, 0}, 1}, {"-2", {0, 0, 0, 0, 1, 0}, 1},
        {"-3", {0, 0, 0, 0, 0, 1}, 1}, {"R", {0, 0, 0, 1, 1, 1}, 1}, {"6", {1, 1, 1, 1, 1, 1}, 1},
    };
    if (!name) name = "";
    for (size_t i = 0; i < sizeof(MAP) / sizeof(MAP[0]); i++)
        if (strcmp(MAP[i].name, name) == 0) {
            memcpy(frames, MAP[i].f, sizeof(MAP[i].f));
            *reverse = MAP[i].rev;
            return 0;
        }
    memset(frames, 0, 6 * sizeof(int));
    *reverse = 0;
    if (err) snprintf(err, errcap, "wrong value for -f | --frame parameter: %s", name);
    return -1;
}

/* ---- growable byte buffer ---- */
typedef struct { uint8_t *p; size_t len, cap; } bytes;

static void bytes_reserve(bytes *b, size_t extra) {
    if (b->len + extra <= b->cap) return;
    size_t nc = b->cap ? b->cap : 4096;
    while (nc < b->len + extra) nc *= 2;
    b->p = (uint8_t *)realloc(b->p, nc);
    if (!b->p) { perror("realloc"); abort(); }
    b->cap = nc;
}
static inline void bytes_push(bytes *b, uint8_t c) {
    if (b->len == b->cap) bytes_reserve(b, 1);
    b->p[b->len++] = c;
}
static inline void bytes_append(bytes *b, const uint8_t *s, size_t n) {
    bytes_reserve(b, n);
    memcpy(b->p + b->len, s, n);
    b->len += n;
}

/* ---- encodedSequence: transeq/encodedSequence.go:15-69 ----
 * s[0:4] = header size + 4 (u32 LE), s[4:hs] = header line, s[hs:] = nucleotide codes. */
typedef struct { uint8_t *s; size_t len; } enc_seq;

static inline size_t seq_header_size(const enc_seq *e) {
    return (size_t)e->s[0] | (size_t)e->s[1] << 8 | (size_t)e->s[2] << 16 | (size_t)e->s[3] << 24;
}
static inline size_t seq_nucl_size(const enc_seq *e) { return e->len - seq_header_size(e); }

/* transeq/encodedSequence.go:17-43 newEncodedSequence */
/* Optional capture of the reference's WARNING lines (encodedSequence.go:39), byte for byte what
 * fmt.Printf writes to stdout: "%s" of s[:headerSize] is the 4 raw little-endian length bytes
 * followed by the header line, and string(n) of a byte is the UTF-8 encoding of the rune n. */
static bytes *g_warn_sink = NULL;

static void warn_line(const enc_seq *e, size_t header_size, uint8_t n, size_t i) {
    static const char a[] = "WARNING: invalid char in sequence ";
    char tail[64];
    bytes_append(g_warn_sink, (const uint8_t *)a, sizeof(a) - 1);
    bytes_append(g_warn_sink, e->s, header_size);
    bytes_append(g_warn_sink, (const uint8_t *)": '", 3);
    if (n < 0x80) {
        bytes_push(g_warn_sink, n);
    } else {
        bytes_push(g_warn_sink, (uint8_t)(0xC0 | (n >> 6)));
        bytes_push(g_warn_sink, (uint8_t)(0x80 | (n & 0x3F)));
    }
    int m = snprintf(tail, sizeof(tail), "' ( pos %zu), replacing with 'N'\n", i);
    bytes_append(g_warn_sink, (const uint8_t *)tail, (size_t)m);
}

static enc_seq new_encoded_sequence(const bytes *buf, size_t header_size, uint64_t *warnings) {
    enc_seq e;
    e.len = 4 + buf->len;
    e.s = (uint8_t *)malloc(e.len ? e.len : 1);
    header_size += 4;
    e.s[0] = (uint8_t)header_size; e.s[1] = (uint8_t)(header_size >> 8);
    e.s[2] = (uint8_t)(header_size >> 16); e.s[3] = (uint8_t)(header_size >> 24);
    if (buf->len) memcpy(e.s + 4, buf->p, buf->len);
    for (size_t i = header_size; i < e.len; i++) {
        switch (e.s[i]) {
        case 'A': case 'a': e.s[i] = A_CODE; break;
        case 'C': case 'c': e.s[i] = C_CODE; break;
        case 'G': case 'g': e.s[i] = G_CODE; break;
        case 'T': case 'U': case 't': case 'u': e.s[i] = T_CODE; break;
        case 'N': case 'n': e.s[i] = N_CODE; break;
        default:
            if (g_warn_sink) warn_line(&e, header_size, e.s[i], i - header_size);
            e.s[i] = N_CODE;
            (*warnings)++; /* the reference prints a WARNING line here (encodedSequence.go:39) */
        }
    }
    return e;
}

/* transeq/encodedSequence.go:71-97 reverseComplement */
static void reverse_complement(enc_seq *e) {
    size_t hs = seq_header_size(e);
    for (size_t i = hs; i < e->len; i++) {
        switch (e->s[i]) {
        case A_CODE: e->s[i] = T_CODE; break;
        case T_CODE: e->s[i] = A_CODE; break;
        case C_CODE: e->s[i] = G_CODE; break;
        case G_CODE: e->s[i] = C_CODE; break;
        default: break;
        }
    }
    if (e->len == 0) return;
    for (size_t i = hs, j = e->len - 1; i < j; i++, j--) {
        uint8_t t = e->s[i]; e->s[i] = e->s[j]; e->s[j] = t;
    }
}

/* ---- writer: transeq/writer.go:23-169 ---- */
typedef struct {
    uint8_t codes[TO_ARRAY_CODE_SIZE]; /* per-worker copy, as in writer.go:24,40 */
    bytes buf;
    int current_line_len;
    int start_pos[3];
    int frame_index;
    int frames_to_generate[6];
    int reverse, alternative, trim;
    long to_trim;
} writer;

static void w_new_line(writer *w) { /* writer.go:150-157 */
    bytes_push(&w->buf, '\n');
    w->current_line_len = 0;
    if (w->trim) w->to_trim++;
}

static void w_write_header(writer *w, const uint8_t *hdr, size_t hlen) { /* writer.go:120-131 */
    const uint8_t *sp = (const uint8_t *)memchr(hdr, ' ', hlen);
    if (sp) {
        size_t end = (size_t)(sp - hdr);
        bytes_append(&w->buf, hdr, end);
        bytes_push(&w->buf, '_'); bytes_push(&w->buf, (uint8_t)SUFFIXES[w->frame_index]);
        bytes_append(&w->buf, hdr + end, hlen - end);
    } else {
        bytes_append(&w->buf, hdr, hlen);
        bytes_push(&w->buf, '_'); bytes_push(&w->buf, (uint8_t)SUFFIXES[w->frame_index]);
    }
    w_new_line(w);
}

static inline void w_write_aa(writer *w, uint8_t aa) { /* writer.go:133-148 */
    if (w->current_line_len == MAX_LINE_SIZE) w_new_line(w);
    bytes_push(&w->buf, aa);
    w->current_line_len++;
    if (w->trim) {
        if (aa == STOP_AA || aa == UNKNOWN_AA) w->to_trim++;
        else w->to_trim = 0;
    }
}

static void w_trim_and_return(writer *w) { /* writer.go:159-169 */
    if (w->to_trim > 0) {
        w->buf.len -= (size_t)w->to_trim;
        w->current_line_len -= (int)w->to_trim;
    }
    if (w->current_line_len != 0) w_new_line(w);
    w->to_trim = 0;
}

static void w_translate_3_frames(writer *w, const enc_seq *e) { /* writer.go:85-116 */
    size_t hs = seq_header_size(e);
    long nucl = (long)seq_nucl_size(e);
    for (int k = 0; k < 3; k++) {
        int start = w->start_pos[k];
        if (w->frames_to_generate[w->frame_index] == 0) { w->frame_index++; continue; }
        w_write_header(w, e->s + 4, hs - 4);
        for (size_t pos = hs + (size_t)start; pos + 2 < e->len; pos += 3) {
            uint32_t index = (uint32_t)e->s[pos] | (uint32_t)e->s[pos + 1] << 8 |
                             (uint32_t)e->s[pos + 2] << 16;
            w_write_aa(w, w->codes[index]);
        }
        /* Go's % truncates toward zero, and so does C's: negative operands fall through */
        switch ((nucl - start) % 3) {
        case 2: {
            uint32_t index = (uint32_t)e->s[e->len - 2] | (uint32_t)e->s[e->len - 1] << 8;
            w_write_aa(w, w->codes[index]);
            break;
        }
        case 1:
            w_write_aa(w, UNKNOWN_AA);
            break;
        }
        w_trim_and_return(w);
        w->frame_index++;
    }
}

static void w_translate(writer *w, enc_seq *e) { /* writer.go:50-83 */
    w->frame_index = 0;
    if (w->reverse && !w->alternative) { w->start_pos[0] = 0; w->start_pos[1] = 1; w->start_pos[2] = 2; }
    w_translate_3_frames(w, e);
    if (w->reverse) {
        if (!w->alternative) {
            switch (seq_nucl_size(e) % 3) {
            case 0: w->start_pos[0] = 0; w->start_pos[1] = 2; w->start_pos[2] = 1; break;
            case 1: w->start_pos[0] = 1; w->start_pos[1] = 0; w->start_pos[2] = 2; break;
            case 2: w->start_pos[0] = 2; w->start_pos[1] = 1; w->start_pos[2] = 0; break;
            }
        }
        reverse_complement(e);
        w_translate_3_frames(w, e);
    }
}

/* ---- channel of capacity 100 (transeq.go:126) ---- */
#define CHAN_CAP 100
typedef struct {
    enc_seq q[CHAN_CAP];
    int head, count, closed;
    pthread_mutex_t mu;
    pthread_cond_t not_empty, not_full;
} chan_t;

static void chan_send(chan_t *c, enc_seq e) {
    pthread_mutex_lock(&c->mu);
    while (c->count == CHAN_CAP) pthread_cond_wait(&c->not_full, &c->mu);
    c->q[(c->head + c->count) % CHAN_CAP] = e;
    c->count++;
    pthread_cond_signal(&c->not_empty);
    pthread_mutex_unlock(&c->mu);
}
static int chan_recv(chan_t *c, enc_seq *e) {
    pthread_mutex_lock(&c->mu);
    while (c->count == 0 && !c->closed) pthread_cond_wait(&c->not_empty, &c->mu);
    if (c->count == 0) { pthread_mutex_unlock(&c->mu); return 0; }
    *e = c->q[c->head];
    c->head = (c->head + 1) % CHAN_CAP;
    c->count--;
    pthread_cond_signal(&c->not_full);
    pthread_mutex_unlock(&c->mu);
    return 1;
}
static void chan_close(chan_t *c) {
    pthread_mutex_lock(&c->mu);
    c->closed = 1;
    pthread_cond_broadcast(&c->not_empty);
    pthread_mutex_unlock(&c->mu);
}

/* ---- shared output (the io.Writer): each flush appends its whole buffer atomically ---- */
typedef struct { bytes out; pthread_mutex_t mu; } sink_t;

typedef struct {
    chan_t *ch;
    sink_t *sink;
    const uint8_t *codes;
    int frames[6];
    int reverse, alternative, trim;
} worker_arg;

static void w_flush(writer *w, sink_t *sink) { /* writer.go:171-181 */
    pthread_mutex_lock(&sink->mu);
    bytes_append(&sink->out, w->buf.p, w->buf.len);
    pthread_mutex_unlock(&sink->mu);
    w->buf.len = 0;
}

static void *worker_main(void *p) { /* transeq.go:137-159 */
    worker_arg *a = (worker_arg *)p;
    writer *w = (writer *)calloc(1, sizeof(writer)); /* newWriter: writer.go:38-48 */
    memcpy(w->codes, a->codes, TO_ARRAY_CODE_SIZE);
    bytes_reserve(&w->buf, MAX_BUFFER_SIZE);
    w->start_pos[0] = 0; w->start_pos[1] = 1; w->start_pos[2] = 2;
    memcpy(w->frames_to_generate, a->frames, sizeof(a->frames));
    w->reverse = a->reverse; w->alternative = a->alternative; w->trim = a->trim;

    enc_seq e;
    while (chan_recv(a->ch, &e)) {
        w_translate(w, &e);
        if (w->buf.len > MAX_BUFFER_SIZE) w_flush(w, a->sink);
        free(e.s);
    }
    w_flush(w, a->sink);
    free(w->buf.p);
    free(w);
    return NULL;
}

/* transeq/transeq.go:183-216 readSequenceFromFasta, with bufio.ScanLines semantics:
 * tokens end at '\n'; one trailing '\r' is dropped; a final unterminated line is returned;
 * a line of >= maxSeqLength bytes makes Scan() fail, which silently ends the loop. */
static void read_sequence_from_fasta(const uint8_t *in, size_t n, chan_t *ch, uint64_t *warnings) {
    bytes buf = {0, 0, 0};
    size_t header_size = 0;
    size_t pos = 0;
    bytes_reserve(&buf, 4096);
    while (pos < n) {
        const uint8_t *nl = (const uint8_t *)memchr(in + pos, '\n', n - pos);
        size_t end = nl ? (size_t)(nl - in) : n;
        size_t next = nl ? end + 1 : n;
        if (end - pos >= g_max_seq_length) break; /* bufio.ErrTooLong, swallowed (transeq.go:192) */
        const uint8_t *line = in + pos;
        size_t len = end - pos;
        pos = next;
        if (len > 0 && line[len - 1] == '\r') len--; /* dropCR */
        if (len == 0) continue;
        if (line[0] == '>') {
            if (buf.len > 0) chan_send(ch, new_encoded_sequence(&buf, header_size, warnings));
            buf.len = 0;
            header_size = len;
        }
        bytes_append(&buf, line, len);
    }
    chan_send(ch, new_encoded_sequence(&buf, header_size, warnings)); /* transeq.go:213 */
    chan_close(ch);
    free(buf.p);
}

/* transeq/transeq.go:114-173 Translate */
int to_translate(const uint8_t *in, size_t n, const to_options *opt, uint8_t **out, size_t *out_n,
                 uint64_t *n_warnings, char *err, size_t errcap) {
    int frames[6], reverse;
    *out = NULL; *out_n = 0;
    if (to_compute_frames(opt->frame, frames, &reverse, err, errcap) != 0) return -1;
    uint8_t *codes = (uint8_t *)malloc(TO_ARRAY_CODE_SIZE);
    if (to_create_code_array(opt->table, opt->clean, codes, err, errcap) != 0) { free(codes); return -1; }
    int nw = opt->num_worker;
    if (nw <= 0) {
        /* the reference would start no worker and block (transeq.go:133-135); not restated */
        if (err) snprintf(err, errcap, "oracle: num_worker must be >= 1");
        free(codes);
        return -1;
    }

    chan_t ch;
    memset(&ch, 0, sizeof(ch));
    pthread_mutex_init(&ch.mu, NULL);
    pthread_cond_init(&ch.not_empty, NULL);
    pthread_cond_init(&ch.not_full, NULL);
    sink_t sink;
    memset(&sink, 0, sizeof(sink));
    pthread_mutex_init(&sink.mu, NULL);

    worker_arg arg;
    arg.ch = &ch; arg.sink = &sink; arg.codes = codes;
    memcpy(arg.frames, frames, sizeof(frames));
    arg.reverse = reverse; arg.alternative = opt->alternative; arg.trim = opt->trim;

    pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * (size_t)nw);
    for (int i = 0; i < nw; i++) pthread_create(&th[i], NULL, worker_main, &arg);
    uint64_t warnings = 0;
    read_sequence_from_fasta(in, n, &ch, &warnings);
    for (int i = 0; i < nw; i++) pthread_join(th[i], NULL);
    free(th);
    free(codes);
    pthread_mutex_destroy(&ch.mu);
    pthread_cond_destroy(&ch.not_empty);
    pthread_cond_destroy(&ch.not_full);
    pthread_mutex_destroy(&sink.mu);

    if (n_warnings) *n_warnings = warnings;
    *out = sink.out.p ? sink.out.p : (uint8_t *)malloc(1);
    *out_n = sink.out.len;
    return 0;
}

/* to_translate plus the text of the WARNING lines (the reader goroutine prints them in input
 * order, encodedSequence.go:39).  Not reentrant: one call at a time. */
int to_translate_warn(const uint8_t *in, size_t n, const to_options *opt, uint8_t **out, size_t *out_n,
                      uint8_t **warn_text, size_t *warn_n, char *err, size_t errcap) {
    bytes sink = {0, 0, 0};
    uint64_t count = 0;
    g_warn_sink = &sink;
    int rc = to_translate(in, n, opt, out, out_n, &count, err, errcap);
    g_warn_sink = NULL;
    *warn_text = sink.p ? sink.p : (uint8_t *)malloc(1);
    *warn_n = sink.len;
    return rc;
}

void to_free(void *p) { free(p); }
