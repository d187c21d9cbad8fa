// which instructions share a pipe on sm_100a: pairs interleaved 1:1, warp instructions per clock per SM
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void op(int OP, unsigned &a, unsigned m, unsigned s) {
    switch (OP) {
        case 0: asm volatile("dp4a.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 1: asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 2: asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 3: asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 4: asm volatile("popc.b32 %0, %0;" : "+r"(a)); break;
        case 5: asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 6: asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 7: asm volatile("add.u32 %0, %0, %1;" : "+r"(a) : "r"(m)); break;
        case 8: asm volatile("{ .reg .pred p; setp.lt.u32 p, %0, %1; selp.u32 %0, %2, %0, p; }" : "+r"(a) : "r"(m), "r"(s)); break;
        case 9: asm volatile("shl.b32 %0, %0, 3;" : "+r"(a)); break;
        case 10: asm volatile("mul.lo.u32 %0, %0, 8;" : "+r"(a)); break;
        case 11: asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a) : "r"(m), "r"(s)); break;
        case 12: asm volatile("mad.lo.u32 %0, %0, 1, %1;" : "+r"(a) : "r"(m)); break;
    }
}
template <int A, int B>
__global__ void k(unsigned *out, unsigned seed, int iters) {
    unsigned a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 8 + i;
    unsigned m = seed * 3 + 1, s = seed | 0x3210;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) op((i & 1) ? B : A, a[i], m, s);
    }
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i];
    if (x == 0x12345678) out[0] = x;
}
template <int A, int B>
void run(const char *name) {
    unsigned *d; cudaMalloc(&d, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 4096;
    k<A, B><<<blocks, threads>>>(d, 1, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<A, B><<<blocks, threads>>>(d, 1, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double winstr = (double)blocks * threads / 32 * iters * 8;
    const double clocks = ms * 1e-3 * 1.965e9;
    printf("%-22s %.3f ms  %.2f warp-instr/clk/SM\n", name, ms, winstr / clocks / p.multiProcessorCount);
}
int main() {
    run<1, 2>("prmt+mad"); run<1, 0>("prmt+dp4a"); run<1, 3>("prmt+vabsdiff4"); run<2, 0>("mad+dp4a"); run<2, 3>("mad+vabsdiff4");
    run<1, 5>("prmt+lop3"); run<1, 6>("prmt+shf"); run<2, 6>("mad+shf"); run<1, 7>("prmt+add"); run<2, 7>("mad+add");
    run<1, 8>("prmt+setp/selp"); run<2, 8>("mad+setp/selp"); run<1, 9>("prmt+shl"); run<1, 10>("prmt+mul8"); run<0, 3>("dp4a+vabsdiff4");
    run<1, 11>("prmt+vabsdiff4.add"); run<1, 12>("prmt+mad1"); run<1, 4>("prmt+popc"); run<2, 4>("mad+popc");
    return 0;
}
