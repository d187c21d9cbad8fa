// transeq.hpp -- C++ host-side mirror of the reference's `transeq` package over the C ABI
// (include/gotranseq_b200.h).  The Go toolchain is absent in this image, so this is the compiled
// host layer a drop-in user links against; go/transeq/ holds the equivalent cgo shim.
//
//   transeq::Options    field for field `type Options struct` (transeq/options.go:4-11)
//   transeq::Translate  `func Translate(inputSequence io.Reader, out io.Writer, options Options) error`
//                       (transeq/transeq.go:114): returns an empty string on success, otherwise the
//                       error text (same texts as the reference: transeq.go:107, ncbicode/code.go:378,
//                       writer.go:175).  WARNING lines for invalid nucleotides go to stdout
//                       (encodedSequence.go:39).
//   transeq::TranslateFd  the same between two file descriptors, the way main.go:52-64 uses it
//                       (os.Open / os.Create): the calling thread reads, a second thread writes.
#pragma once
#include <istream>
#include <ostream>
#include <string>

namespace transeq {

struct Options {
    std::string Frame = "1";   // -f | --frame
    int Table = 0;             // -t | --table
    bool Clean = false;        // -c | --clean
    bool Alternative = false;  // -a | --alternative
    bool Trim = false;         // -T | --trim
    int NumWorker = 0;         // -n | --numcpu: accepted and ignored (output is in input order)
    int Device = -1;           // CUDA device ordinal when NumGpus == 0, -1 = current
    int NumGpus = -1;          // 0 = one GPU (Device), -1 = every visible GPU, n = GPUs 0..n-1
    unsigned long long ChunkBytes = 0;  // streaming chunk size, 0 = library default
    bool Warnings = true;      // print the reference's WARNING lines to stdout
    bool AnyOrder = false;     // results are written as they complete (the reference's -n > 1 order), not in input order
    int IoThreads = 0;         // TranslateFd: threads per file read / write (0 = 4)
};

std::string Translate(std::istream &inputSequence, std::ostream &out, const Options &options);
std::string TranslateFd(int fd_in, int fd_out, const Options &options, const char *out_path = "");

}  // namespace transeq
