// gts_kernels.h -- host-callable launchers of the kernels in gts_kernels.cu.
#pragma once
#include <cuda_runtime.h>

#include "gts_device.h"

namespace gts {

// Per-device one-time kernel attribute setup (dynamic shared memory of k_translate).
cudaError_t init_kernels();

// k_parse + k_size, and unless sizes_only also k_translate + k_headers, on `stream`.
// max_symbols bounds the symbol stream (job.n + 4); *launches is advanced by the kernels started.
// ev, when not null, points to 5 events recorded before the first and after each kernel.
cudaError_t launch_pass(const Job &job, u64 max_symbols, int sm_count, cudaStream_t stream,
                        bool sizes_only, unsigned long long *launches, cudaEvent_t *ev = nullptr);

// k_translate + k_headers only (after a sizes_only pass, once the output buffer is known).
cudaError_t launch_emit(const Job &job, u64 max_symbols, int sm_count, cudaStream_t stream,
                        unsigned long long *launches);

}  // namespace gts
