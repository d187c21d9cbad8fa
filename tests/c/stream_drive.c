/* Drives the exact call sequence of the Go shim (go/transeq/translate_b200.go) from plain C, on the
 * device: gts_create -> { gts_acquire_input -> read -> gts_submit -> drain gts_next_output /
 * gts_release_output } -> gts_destroy.  Built and run by tests/test_stream_c.py (-m gpu).
 *
 *   stream_drive IN OUT FRAME TABLE CLEAN ALT TRIM CHUNK_BYTES NUM_GPUS READ_SIZE [fail_after_writes]
 *
 * NUM_GPUS > 0 uses device 0 that many times (every "GPU" of the context is the same device: the
 * dealing and the in-order hand-back are exercised on a single-GPU box).  With fail_after_writes the
 * output "file" fails on that write: the program must then see GTS_ERR_WRITE with the reference's
 * message (writer.go:175) and the context must be reusable after gts_stream_reset.
 * WARNING lines (encodedSequence.go:39) go to stdout, formatted by gts_format_warning. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "gotranseq_b200.h"

static void warn(void *user, const uint8_t *header, size_t header_len, uint8_t byte, uint64_t pos) {
    char line[4096];
    size_t n = gts_format_warning(header, header_len, byte, pos, line, sizeof(line));
    (void)user;
    if (n > sizeof(line)) n = sizeof(line);
    fwrite(line, 1, n, stdout);
}

static int run(gts_ctx *ctx, const char *in_path, const char *out_path, size_t read_size, long fail_after) {
    FILE *in = fopen(in_path, "rb"), *out = fopen(out_path, "wb");
    int eof = 0, rc = GTS_OK;
    long writes = 0;
    if (!in || !out) return 100;
    while (!eof && rc == GTS_OK) {
        uint8_t *buf = NULL;
        size_t cap = 0, got;
        rc = gts_acquire_input(ctx, &buf, &cap);
        if (rc != GTS_OK) break;
        if (read_size && cap > read_size) cap = read_size;
        got = fread(buf, 1, cap, in);
        eof = got == 0;
        rc = gts_submit(ctx, got, eof);
        if (rc != GTS_OK) break;
        for (;;) {
            const uint8_t *p = NULL;
            size_t m = 0;
            rc = gts_next_output(ctx, &p, &m);
            if (rc != GTS_OK || m == 0) break;
            if (fail_after >= 0 && writes == fail_after) {
                gts_report_write_error(ctx, "disk full");
                continue; /* the next gts_next_output reports GTS_ERR_WRITE */
            }
            fwrite(p, 1, m, out);
            writes++;
            gts_release_output(ctx);
        }
    }
    fclose(in);
    fclose(out);
    return rc;
}

int main(int argc, char **argv) {
    gts_options opt;
    gts_ctx *ctx = NULL;
    int rc, i;
    long fail_after = -1;
    if (argc < 11) return 99;
    memset(&opt, 0, sizeof(opt));
    opt.frame = argv[3];
    opt.table = atoi(argv[4]);
    opt.clean = (uint8_t)atoi(argv[5]);
    opt.alternative = (uint8_t)atoi(argv[6]);
    opt.trim = (uint8_t)atoi(argv[7]);
    opt.chunk_bytes = (uint64_t)atoll(argv[8]);
    opt.num_gpus = atoi(argv[9]);
    for (i = 0; i < 8; i++) opt.device_ids[i] = 0;
    opt.device = 0;
    if (argc > 11) fail_after = atol(argv[11]);
    if (gts_create(&opt, &ctx) != GTS_OK) {
        fprintf(stderr, "create: %s\n", gts_last_error(NULL));
        return 98;
    }
    gts_set_warning_callback(ctx, warn, NULL);
    rc = run(ctx, argv[1], argv[2], (size_t)atoll(argv[10]), fail_after);
    if (fail_after >= 0) {
        if (rc != GTS_ERR_WRITE || strcmp(gts_last_error(ctx), "fail to write to output file: disk full") != 0) {
            fprintf(stderr, "expected the write error, got %d: %s\n", rc, gts_last_error(ctx));
            return 97;
        }
        /* the context is reusable: the same stream again, this time to the end */
        if (gts_stream_reset(ctx) != GTS_OK) return 96;
        rc = run(ctx, argv[1], argv[2], (size_t)atoll(argv[10]), -1);
    }
    if (rc != GTS_OK) {
        fprintf(stderr, "stream: %d %s\n", rc, gts_last_error(ctx));
        return 95;
    }
    fprintf(stderr, "stream ok on %d device contexts\n", gts_num_devices(ctx));
    gts_destroy(ctx);
    return 0;
}
