// main.cpp -- `gotranseq` command line over the B200 path, flag for flag the reference's CLI
// (main.go:21-91, transeq/options.go:5-10): -s/--sequence, -o/--outseq, -f/--frame, -t/--table,
// -c/--clean, -a/--alternative, -T/--trim, -n/--numcpu, -h/--help, -v/--version.
// Like the reference, an error from the run is printed to stdout as "fail to translate file:\n%v"
// and the exit code stays 0 (main.go:87-90); flag parsing errors print "wrong arguments: ..., try
// gotranseq --help for more informations" and exit with 1 (main.go:73-76).  The flag parser of the
// reference is go-flags v1.4.0 (go.mod:5, not vendored): its error texts are restated from memory
// of that library and are not pinned by any reference test.
// Environment (not flags, so that the flag set stays the reference's): GOTRANSEQ_B200_GPUS = number
// of GPUs (default: all visible), GOTRANSEQ_B200_CHUNK_MB = chunk size.
#include <fcntl.h>
#include <unistd.h>

#include <cerrno>
#include <climits>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

[[noreturn]] static void wrong_arguments(const std::string &what) {
    // main.go:74
    std::printf("wrong arguments: %s, try %s --help for more informations\n", what.c_str(), kTool);
    std::exit(1);
}

// Go's syscall.Errno texts are the C library's with a lower-case first letter
static std::string go_errno(int e) {
    std::string s = std::strerror(e);
    if (!s.empty() && s[0] >= 'A' && s[0] <= 'Z') s[0] = (char)(s[0] - 'A' + 'a');
    return s;
}

static int parse_int(const std::string &v, const char *flag) {
    errno = 0;
    char *end = nullptr;
    const long x = std::strtol(v.c_str(), &end, 10);
    if (v.empty() || *end != 0 || errno == ERANGE || x < INT_MIN || x > INT_MAX)
        wrong_arguments(std::string("invalid argument for flag `") + flag + "' (expected int): strconv.ParseInt: parsing \"" +
                        v + "\": " + (errno == ERANGE ? "value out of range" : "invalid syntax"));
    return (int)x;
}

int main(int argc, char **argv) {
    transeq::Options opt;
    std::string seq, outseq;
    bool help = false, version = false;
    auto value = [&](int &i, const char *arg, const char *shortf, const char *longf, std::string &dst) -> bool {
        const size_t ll = std::strlen(longf);
        if (std::strcmp(arg, shortf) == 0 || std::strcmp(arg, longf) == 0) {
            if (i + 1 >= argc) wrong_arguments(std::string("expected argument for flag `") + shortf + ", " + longf + "'");
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
        if (value(i, a, "-t", "--table", v)) { opt.Table = parse_int(v, "-t, --table"); continue; }
        if (value(i, a, "-n", "--numcpu", v)) { opt.NumWorker = parse_int(v, "-n, --numcpu"); continue; }
        if (!std::strcmp(a, "-c") || !std::strcmp(a, "--clean")) { opt.Clean = true; continue; }
        if (!std::strcmp(a, "-a") || !std::strcmp(a, "--alternative")) { opt.Alternative = true; continue; }
        if (!std::strcmp(a, "-T") || !std::strcmp(a, "--trim")) { opt.Trim = true; continue; }
        if (!std::strcmp(a, "-h") || !std::strcmp(a, "--help")) { help = true; continue; }
        if (!std::strcmp(a, "-v") || !std::strcmp(a, "--version")) { version = true; continue; }
        if (a[0] == '-' && a[1] != 0) wrong_arguments(std::string("unknown flag `") + (a + (a[1] == '-' ? 2 : 1)) + "'");
        // (positional arguments are ignored, as go-flags hands them back unused)
    }
    if (help) { usage(); return 0; }  // main.go:77-81
    if (version) { std::printf("%s (B200) version %s\n", kTool, GTS_VERSION); return 0; }
    if (const char *g = std::getenv("GOTRANSEQ_B200_GPUS")) opt.NumGpus = std::atoi(g);
    if (const char *t = std::getenv("GOTRANSEQ_B200_IO_THREADS")) opt.IoThreads = std::atoi(t);
    if (const char *a = std::getenv("GOTRANSEQ_B200_ANY_ORDER")) opt.AnyOrder = std::atoi(a) != 0;
    if (const char *c = std::getenv("GOTRANSEQ_B200_CHUNK_MB")) opt.ChunkBytes = (unsigned long long)(std::atof(c) * (1 << 20));

    // run(), main.go:39-65
    std::string err;
    if (seq.empty()) {
        err = std::string("missing required parameter -s | -sequence, try ") + kTool + " --help for details";
    } else if (outseq.empty()) {
        err = std::string("missing required parameter -o | -outseq, try ") + kTool + " --help for details";
    } else {
        const int in = ::open(seq.c_str(), O_RDONLY);
        if (in < 0) {
            err = "open " + seq + ": " + go_errno(errno);  // (*os.PathError: "open <path>: <errno text>")
        } else {
            const int out = ::open(outseq.c_str(), O_WRONLY | O_CREAT | O_TRUNC, 0666);
            if (out < 0) {
                err = "open " + outseq + ": " + go_errno(errno);
            } else {
                err = transeq::TranslateFd(in, out, opt, outseq.c_str());
                ::close(out);
            }
            ::close(in);
        }
    }
    if (!err.empty()) std::printf("fail to translate file:\n%s", err.c_str());  // main.go:89 (no newline)
    std::fflush(stdout);
    return 0;
}
