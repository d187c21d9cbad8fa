// gts_kernels.h -- host-callable launchers of the kernels (gts_parse.cu, gts_size.cu,
// gts_lines.cu, gts_headers.cu).
#pragma once
#include <cuda_runtime.h>

#include "gts_device.h"

namespace gts {

// Per-device one-time kernel attribute setup (dynamic shared memory of k_lines).
cudaError_t init_kernels();

cudaError_t launch_parse(const Job &job, cudaStream_t stream);                // k_parse
cudaError_t launch_sizes(const Job &job, int sm_count, cudaStream_t stream);  // k_tile_scan, k_records, k_size
cudaError_t launch_tile_scan(const Job &job, int sm_count, cudaStream_t stream);  // k_tile_scan alone (piece scan)
cudaError_t launch_piece_plan(const PieceInfoDev *all, int rank, int world, PieceParams *out, cudaStream_t stream);  // k_piece_plan
cudaError_t launch_piece_fix(const Job &job, cudaStream_t stream);    // k_piece_fix: carry / end pads into reused slots
cudaError_t launch_piece_trim(const Job &job, cudaStream_t stream);   // k_piece_trim: one piece's part of the trim walk
// k_publish: the pass's totals into pinned host memory by stores from the device (no copy engine involved)
cudaError_t launch_publish(const Totals *d_totals, Totals *h_totals, cudaStream_t stream);
cudaError_t launch_warn(const Job &job, int sm_count, cudaStream_t stream);   // k_warn (only when job.warn is set)
// k_lines and k_headers; with a side stream (and two events to fork / join) the two run
// concurrently; ev_mid (optional) is recorded on `stream` after k_lines
struct EmitStreams {  // side streams of the emit stage (k_headers and k_short run next to k_lines) and their events
    cudaStream_t side = nullptr, side2 = nullptr;
    cudaEvent_t fork = nullptr, join = nullptr, join2 = nullptr;
};
cudaError_t launch_emit(const Job &job, int sm_count, cudaStream_t stream, cudaEvent_t ev_mid, const EmitStreams &es);

}  // namespace gts
