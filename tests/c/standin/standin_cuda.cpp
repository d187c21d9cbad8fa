// standin_cuda.cpp -- TEST INFRASTRUCTURE ONLY: the few CUDA runtime calls of gotranseq_b200/csrc/gts_api.cu,
// carried out on the host, so that the HOST logic of the library (chunk engine: record-boundary cuts, carry-over,
// buffer pools, in-order / any-order hand-back, cancellation, write errors, warnings, over-long lines, retries)
// can be tested without a GPU (`-m "not gpu"`).  "Device memory" is host memory, every copy / memset happens at
// the call, streams and events are empty handles.  Together with standin_pass.cpp (the device pass restated on
// the host) and gts_api.cu compiled as plain C++ this makes tests/c/standin/libgotranseq_standin.so, which only
// tests/test_engine_standin.py loads (through GOTRANSEQ_B200_LIB, in a process of its own).  Nothing in
// gotranseq_b200/ refers to it; the product has no CPU path.
#include <cuda_runtime_api.h>

#include <cstdlib>
#include <cstring>

namespace {
thread_local int t_device = 0;
int device_count() {
    const char *e = std::getenv("GTS_STANDIN_GPUS");
    const int n = e ? std::atoi(e) : 2;
    return n < 1 ? 1 : n;
}
void *alloc(size_t size, size_t align) {
    if (size == 0) size = 1;
    const size_t rounded = (size + align - 1) / align * align;
    return std::aligned_alloc(align, rounded);
}
}  // namespace

extern "C" {

cudaError_t cudaGetDeviceCount(int *count) { *count = device_count(); return cudaSuccess; }
cudaError_t cudaSetDevice(int device) {
    if (device < 0 || device >= device_count()) return cudaErrorInvalidDevice;
    t_device = device;
    return cudaSuccess;
}
cudaError_t cudaGetDevice(int *device) { *device = t_device; return cudaSuccess; }
cudaError_t cudaGetDeviceProperties_v2(cudaDeviceProp *prop, int device) {
    if (device < 0 || device >= device_count()) return cudaErrorInvalidDevice;
    std::memset(prop, 0, sizeof(*prop));
    std::strcpy(prop->name, "host stand-in");
    prop->major = 10;
    prop->minor = 0;
    prop->multiProcessorCount = 148;
    return cudaSuccess;
}
cudaError_t cudaDeviceGetStreamPriorityRange(int *least, int *greatest) { *least = 0; *greatest = -5; return cudaSuccess; }
const char *cudaGetErrorString(cudaError_t e) { return e == cudaSuccess ? "no error" : "stand-in CUDA error"; }

cudaError_t cudaMalloc(void **p, size_t size) {
    *p = alloc(size, 256);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFree(void *p) { std::free(p); return cudaSuccess; }
cudaError_t cudaHostAlloc(void **p, size_t size, unsigned int) {
    *p = alloc(size, 4096);
    return *p ? cudaSuccess : cudaErrorMemoryAllocation;
}
cudaError_t cudaFreeHost(void *p) { std::free(p); return cudaSuccess; }

cudaError_t cudaMemcpy(void *dst, const void *src, size_t n, cudaMemcpyKind) { std::memcpy(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemcpyAsync(void *dst, const void *src, size_t n, cudaMemcpyKind, cudaStream_t) { std::memcpy(dst, src, n); return cudaSuccess; }
cudaError_t cudaMemset(void *p, int v, size_t n) { std::memset(p, v, n); return cudaSuccess; }
cudaError_t cudaMemsetAsync(void *p, int v, size_t n, cudaStream_t) { std::memset(p, v, n); return cudaSuccess; }

cudaError_t cudaStreamCreateWithFlags(cudaStream_t *s, unsigned int) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamCreateWithPriority(cudaStream_t *s, unsigned int, int) { *s = reinterpret_cast<cudaStream_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaStreamDestroy(cudaStream_t s) { std::free(s); return cudaSuccess; }
cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
cudaError_t cudaStreamWaitEvent(cudaStream_t, cudaEvent_t, unsigned int) { return cudaSuccess; }
cudaError_t cudaDeviceSynchronize(void) { return cudaSuccess; }

cudaError_t cudaEventCreate(cudaEvent_t *e) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaEventCreateWithFlags(cudaEvent_t *e, unsigned int) { *e = reinterpret_cast<cudaEvent_t>(std::malloc(8)); return cudaSuccess; }
cudaError_t cudaEventDestroy(cudaEvent_t e) { std::free(e); return cudaSuccess; }
cudaError_t cudaEventRecord(cudaEvent_t, cudaStream_t) { return cudaSuccess; }
cudaError_t cudaEventSynchronize(cudaEvent_t) { return cudaSuccess; }
cudaError_t cudaEventElapsedTime(float *ms, cudaEvent_t, cudaEvent_t) { *ms = 0.0f; return cudaSuccess; }

}  // extern "C"
