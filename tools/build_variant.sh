#!/bin/bash
# usage: build_variant.sh NAME [-DFLAG ...]   -> build/NAME.so (an A/B copy of the library; GOTRANSEQ_B200_LIB selects it)
name=$1; shift
mkdir -p build
cd gotranseq_b200/csrc
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-pthread -Xptxas -v "$@" -shared \
  -o ../../build/$name.so gts_parse.cu gts_size.cu gts_headers.cu gts_lines.cu gts_short.cu gts_api.cu 2> ../../build/$name.log || { tail -20 ../../build/$name.log; exit 1; }
grep -A1 "Compiling entry function '_ZN3gts[0-9]*k_[a-z_]*" ../../build/$name.log | grep -E "Compiling|registers" | sed -e 's/.*_ZN3gts[0-9]*\(k_[a-z_]*\)E.*/\1/' | paste - - | awk '{print $1, $5, $6}' | tr '\n' ';'; echo
