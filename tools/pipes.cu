// throughput of a few integer instructions on sm_100a: warp instructions per clock per SM
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(unsigned *out, unsigned seed, int iters) {
    unsigned a[8];
#pragma unroll
    for (int i = 0; i < 8; i++) a[i] = seed + threadIdx.x * 8 + i;
    unsigned m = seed * 3 + 1, s = seed | 0x3210;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++) {
            if (OP == 0) asm volatile("dp4a.u32.u32 %0, %1, %2, %0;" : "+r"(a[i]) : "r"(a[(i + 1) & 7]), "r"(m));
            if (OP == 1) asm volatile("prmt.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
            if (OP == 2) asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
            if (OP == 3) asm volatile("vabsdiff4.u32.u32.u32.add %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
            if (OP == 4) asm volatile("popc.b32 %0, %0;" : "+r"(a[i]));
            if (OP == 5) asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[i]) : "r"(m), "r"(s));
            if (OP == 6) asm volatile("vabsdiff4.u32.u32.u32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
            if (OP == 7) asm volatile("shf.l.wrap.b32 %0, %0, %1, %2;" : "+r"(a[i]) : "r"(m), "r"(s));
            if (OP == 8) asm volatile("bfind.u32 %0, %0;" : "+r"(a[i]));
        }
    }
    unsigned x = 0;
#pragma unroll
    for (int i = 0; i < 8; i++) x ^= a[i];
    if (x == 0x12345678) out[0] = x;
}
template <int OP>
void run(const char *name) {
    unsigned *d; cudaMalloc(&d, 4);
    cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
    const int blocks = p.multiProcessorCount * 2, threads = 1024, iters = 4096;
    k<OP><<<blocks, threads>>>(d, 1, 64);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<OP><<<blocks, threads>>>(d, 1, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    int khz; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double winstr = (double)blocks * threads / 32 * iters * 8;
    const double clocks = ms * 1e-3 * 1.965e9;
    printf("%-12s %.3f ms  %.2f warp-instr/clk/SM  (%.1f lanes/clk/SM)\n", name, ms, winstr / clocks / p.multiProcessorCount, 32 * winstr / clocks / p.multiProcessorCount);
}
int main() {
    run<0>("dp4a"); run<1>("prmt"); run<2>("mad.lo"); run<3>("vabsdiff4.add"); run<6>("vabsdiff4"); run<4>("popc"); run<5>("lop3"); run<7>("shf"); run<8>("bfind");
    return 0;
}
