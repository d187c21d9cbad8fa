"""gotranseq_b200 -- B200-native translation hot path of gotranseq behind its own API.

Host-side mirror of the reference's `transeq` package (transeq/transeq.go:114, options.go:4-11):

    from gotranseq_b200 import Options, Translate
    Translate(open("in.fna", "rb"), open("out.faa", "wb"), Options(Frame="6", Table=11, Clean=True))

All arithmetic runs in the CUDA kernels behind the C ABI of include/gotranseq_b200.h; this
package only moves bytes.  There is no CPU fallback.
"""
from .transeq import (Options, Translate, TranseqError, Translator, translate_bytes)  # noqa: F401

__all__ = ["Options", "Translate", "TranseqError", "Translator", "translate_bytes"]
