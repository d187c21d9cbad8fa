/*
 * transeq_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C) of the gotranseq translation hot path, used as the
 * parity oracle for the CUDA implementation and as the timed CPU baseline.
 * Nothing in gotranseq_b200/ may include, link or call this; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 *
 * Parity status: PINNED against the reference's own golden vectors
 * (transeq/testdata/test.fna x transeq/testdata/data.json, 29 cases, asserted by
 * transeq/transeq_test.go:55) -- see tests/test_oracle_goldens.py.
 * The Go reference itself cannot be compiled here (no Go toolchain), so there is
 * no oracle/_ref build; see DESIGN.md.
 */
#ifndef TRANSEQ_ORACLE_H
#define TRANSEQ_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Mirrors transeq/options.go:4-11 field for field. */
typedef struct {
    const char *frame;   /* "1","2","3","F","-1","-2","-3","R","6" */
    int table;           /* NCBI table id */
    int clean;
    int alternative;
    int trim;
    int num_worker;      /* 1 => deterministic (input) order, like the reference test */
} to_options;

/* size of the reference's codon array: transeq/transeq.go:25 */
#define TO_ARRAY_CODE_SIZE ((4u | 4u << 8 | 4u << 16) + 1u)

/* transeq/transeq.go:30-86. Returns 0 or -1 (invalid table; message in err). */
int to_create_code_array(int table, int clean, uint8_t *codes /*TO_ARRAY_CODE_SIZE*/,
                         char *err, size_t errcap);

/* transeq/transeq.go:88-110. Returns 0 or -1. */
int to_compute_frames(const char *name, int frames[6], int *reverse, char *err, size_t errcap);

/* ncbicode/code.go:367-388: fills aa[64] in T,C,A,G order (index 16*b0+4*b1+b2). */
int to_load_table_code(int table, char aa[64], char *err, size_t errcap);

/*
 * transeq/transeq.go:114-173 on an in-memory input. *out is malloc'ed (caller frees with
 * to_free). Returns 0, or -1 with a message in err (same text as the reference's errors).
 * n_warnings (optional) receives the number of "WARNING: invalid char" lines the reference
 * would have printed (encodedSequence.go:39); they are not printed.
 */
int to_translate(const uint8_t *in, size_t n, const to_options *opt,
                 uint8_t **out, size_t *out_n, uint64_t *n_warnings,
                 char *err, size_t errcap);

/* As to_translate, and *warn_text (malloc'ed) receives the bytes the reference would have printed
 * to stdout: one "WARNING: invalid char in sequence %s: '%s' ( pos %d), replacing with 'N'\n"
 * line per invalid nucleotide, in input order (encodedSequence.go:39; the first %s is the 4 raw
 * little-endian bytes of header length + 4 followed by the header line). */
int to_translate_warn(const uint8_t *in, size_t n, const to_options *opt, uint8_t **out, size_t *out_n,
                      uint8_t **warn_text, size_t *warn_n, char *err, size_t errcap);

void to_free(void *p);

/* Test hook: override the bufio.Scanner max token size (transeq.go:27,186; default 100 MiB). */
void to_set_max_line_length(size_t n);

#ifdef __cplusplus
}
#endif
#endif
