// main.cpp -- `gotranseq` command line over the B200 path, flag for flag the reference's CLI
// (main.go:21-91, transeq/options.go:5-10): -s/--sequence, -o/--outseq, -f/--frame, -t/--table,
// -c/--clean, -a/--alternative, -T/--trim, -n/--numcpu, -h/--help, -v/--version.
// Like the reference, an error from the run is printed to stdout and the exit code stays 0
// (main.go:87-90); only flag parsing errors exit with 1 (main.go:73-76).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>

#include "../../include/gotranseq_b200.h"
#include "transeq.hpp"

static const char *kTool = "gotranseq";

static void usage() {
    std::printf(
        "%s (B200) version %s\n\nUsage:\n  %s --sequence file.fna --outseq out.faa\n\n"
        "required:\n  -s, --sequence=<filename>    Nucleotide sequence(s) filename\n"
        "  -o, --outseq=<filename>      Protein sequence filename\n\n"
        "optional:\n  -f, --frame=<code>           Frame to translate. Possible values:\n"
        "                               [1, 2, 3, F, -1, -2, -3, R, 6] (default: 1)\n"
        "  -t, --table=<code>           NCBI code to use (default: 0); available codes:\n"
        "                               0 2 3 4 5 6 9 10 11 12 13 14 16 21 22 23 24 25 26 29 30\n"
        "  -c, --clean                  Replace stop codon '*' by 'X'\n"
        "  -a, --alternative            Define frame '-1' as using the set of codons starting with the last codon of the sequence\n"
        "  -T, --trim                   Removes all 'X' and '*' characters from the right end of the translation\n"
        "  -n, --numcpu=<n>             Accepted for compatibility; the GPU path has no worker pool\n\n"
        "general:\n  -h, --help                   Show this help message\n"
        "  -v, --version                Print the tool version and exit\n",
        kTool, GTS_VERSION, kTool);
}

int main(int argc, char **argv) {
    transeq::Options opt;
    std::string seq, outseq;
    auto value = [&](int &i, const char *arg, const char *shortf, const char *longf, std::string &dst) -> bool {
        const size_t ll = std::strlen(longf);
        if (std::strcmp(arg, shortf) == 0 || std::strcmp(arg, longf) == 0) {
            if (i + 1 >= argc) { std::printf("expected argument for flag `%s'\n", arg); std::exit(1); }
            dst = argv[++i];
            return true;
        }
        if (std::strncmp(arg, longf, ll) == 0 && arg[ll] == '=') { dst = arg + ll + 1; return true; }
        if (std::strncmp(arg, shortf, 2) == 0 && arg[2] != 0) { dst = arg + 2 + (arg[2] == '=' ? 1 : 0); return true; }
        return false;
    };
    for (int i = 1; i < argc; i++) {
        const char *a = argv[i];
        std::string v;
        if (value(i, a, "-s", "--sequence", seq) || value(i, a, "-o", "--outseq", outseq) ||
            value(i, a, "-f", "--frame", opt.Frame)) continue;
        if (value(i, a, "-t", "--table", v)) { opt.Table = std::atoi(v.c_str()); continue; }
        if (value(i, a, "-n", "--numcpu", v)) { opt.NumWorker = std::atoi(v.c_str()); continue; }
        if (!std::strcmp(a, "-c") || !std::strcmp(a, "--clean")) { opt.Clean = true; continue; }
        if (!std::strcmp(a, "-a") || !std::strcmp(a, "--alternative")) { opt.Alternative = true; continue; }
        if (!std::strcmp(a, "-T") || !std::strcmp(a, "--trim")) { opt.Trim = true; continue; }
        if (!std::strcmp(a, "-h") || !std::strcmp(a, "--help")) { usage(); return 0; }
        if (!std::strcmp(a, "-v") || !std::strcmp(a, "--version")) { std::printf("%s (B200) version %s\n", kTool, GTS_VERSION); return 0; }
        std::printf("unknown flag `%s'\n", a);
        return 1;
    }
    std::string err;
    if (seq.empty()) {
        err = std::string("missing required parameter -s | -sequence, try ") + kTool + " --help for details";
    } else if (outseq.empty()) {
        err = std::string("missing required parameter -o | -outseq, try ") + kTool + " --help for details";
    } else {
        std::ifstream in(seq, std::ios::binary);
        if (!in) {
            err = "open " + seq + ": no such file or directory";
        } else {
            std::ofstream out(outseq, std::ios::binary | std::ios::trunc);
            if (!out) err = "open " + outseq + ": cannot create";
            else err = transeq::Translate(in, out, opt);
        }
    }
    if (!err.empty()) std::printf("%s\n", err.c_str());
    return 0;
}
