/*
 * transeq_oracle_cli.c -- TEST INFRASTRUCTURE ONLY.
 * File -> file driver around the oracle, with the reference's flag names (main.go:21-37),
 * used for timing the restated CPU path (bench.py cpu_baseline / --impl reference).
 */
#include "transeq_oracle.h"

#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>

static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int main(int argc, char **argv) {
    const char *seq = NULL, *outp = NULL;
    to_options opt = {"1", 0, 0, 0, 0, 0};
    int timing = 0;
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        const char *v = (i + 1 < argc) ? argv[i + 1] : NULL;
        if ((!strcmp(a, "-s") || !strcmp(a, "--sequence")) && v) { seq = v; i++; }
        else if ((!strcmp(a, "-o") || !strcmp(a, "--outseq")) && v) { outp = v; i++; }
        else if ((!strcmp(a, "-f") || !strcmp(a, "--frame")) && v) { opt.frame = v; i++; }
        else if ((!strcmp(a, "-t") || !strcmp(a, "--table")) && v) { opt.table = atoi(v); i++; }
        else if ((!strcmp(a, "-n") || !strcmp(a, "--numcpu")) && v) { opt.num_worker = atoi(v); i++; }
        else if (!strcmp(a, "-c") || !strcmp(a, "--clean")) opt.clean = 1;
        else if (!strcmp(a, "-a") || !strcmp(a, "--alternative")) opt.alternative = 1;
        else if (!strcmp(a, "-T") || !strcmp(a, "--trim")) opt.trim = 1;
        else if (!strcmp(a, "--timing")) timing = 1;
        else { fprintf(stderr, "unknown argument %s\n", a); return 1; }
    }
    if (!seq || !outp) { fprintf(stderr, "usage: -s in.fna -o out.faa [-f F] [-t T] [-c] [-a] [-T] [-n N]\n"); return 1; }
    if (opt.num_worker == 0) opt.num_worker = (int)sysconf(_SC_NPROCESSORS_ONLN); /* main.go:48-50 */

    FILE *f = fopen(seq, "rb");
    if (!f) { perror(seq); return 1; }
    fseek(f, 0, SEEK_END);
    long n = ftell(f);
    fseek(f, 0, SEEK_SET);
    uint8_t *in = (uint8_t *)malloc(n > 0 ? (size_t)n : 1);
    if (fread(in, 1, (size_t)n, f) != (size_t)n) { perror("read"); return 1; }
    fclose(f);

    uint8_t *out; size_t out_n; uint64_t warns; char err[256];
    double t0 = now_s();
    if (to_translate(in, (size_t)n, &opt, &out, &out_n, &warns, err, sizeof err) != 0) {
        printf("fail to translate file:\n%s", err);
        return 0; /* main.go:87-90: error printed, exit code 0 */
    }
    double t1 = now_s();
    f = fopen(outp, "wb");
    if (!f) { perror(outp); return 1; }
    fwrite(out, 1, out_n, f);
    fclose(f);
    if (timing) fprintf(stderr, "{\"translate_s\": %.6f, \"in_bytes\": %ld, \"out_bytes\": %zu, \"workers\": %d, \"warnings\": %llu}\n",
                        t1 - t0, n, out_n, opt.num_worker, (unsigned long long)warns);
    to_free(out);
    free(in);
    return 0;
}
