// gts_headers.cu -- k_headers and the launch of the emit stage (k_lines + k_headers).
//
// Reference behaviour reproduced (feliixx/gotranseq):
//   k_headers    writeHeader            transeq/writer.go:120-131
#include <cstdlib>

#include "gts_common.cuh"
#include "gts_kernels.h"

#ifndef GTS_HDR_BLOCKS
#define GTS_HDR_BLOCKS 8
#endif

namespace gts {

// =====================================================================================
// k_headers
// =====================================================================================
// Header line of one frame (writer.go:120-131): the header's bytes with '_' and the frame digit
// inserted before its first space (or appended), then '\n'.  A group of kHdrLanes lanes per record,
// each lane holding four consecutive bytes of the line per step: the line is assembled once (only the
// digit differs between frames) and stored once per requested frame, in front of the frame's residues
// (frame table of k_size).  Sixteen lanes per record measured best (profiles/r2_notes.md): the lines
// sit between residues that k_lines writes, so the kernel is bound by its scattered partial-sector
// accesses -- fewer lanes per record put more records, i.e. more sectors, into every warp instruction
// (8 / 4 / 2 lanes: 52 / 79 / 106 us against 42 us alone on the GPU).
#ifndef GTS_HDR_LANES
#define GTS_HDR_LANES 16
#endif
__global__ void __launch_bounds__(128) k_headers(const __grid_constant__ Job job) {
    constexpr int G = GTS_HDR_LANES;
    const int sub = threadIdx.x & (G - 1);
    const u64 R = job.totals->n_headers;
    if (job.totals->flags != 0 || piece_of(job).no_headers) return;
    if (job.totals->records != 0 && job.totals->n_short == job.totals->records) return;  // reads only: all k_short's
    const u64 ngroups = ((u64)gridDim.x * blockDim.x) / G;
    for (u64 s = ((u64)blockIdx.x * blockDim.x + threadIdx.x) / G; s <= R; s += ngroups) {
        // everything that is indexed by the record goes first, in one round of loads
        const RecInfo ri = job.rec[s];
        const u64 hinf = s >= 1 ? job.hinfo[s] : 0;
        const u64 hoff = s >= 1 ? job.hdr_off[s] : 0;
        const FrameTab *ft = job.ftab + s;  // each frame's header line ends where its residues begin
        u64 base[6];
#pragma unroll
        for (int f = 0; f < 6; f++) base[f] = ft->body[f];
        if (ri.emitted != 1) continue;  // (2: the whole record is written by k_short)
        const u32 H = ri.H;
        const u32 sp = (u32)(hinf >> 32);
        const uint8_t *hdr = job.in + hoff;
#pragma unroll
        for (int f = 0; f < 6; f++) base[f] -= H + 3;
        for (u32 i0 = 4u * (u32)sub; i0 < H + 3; i0 += 4u * G) {
            // bytes i0 .. i0+3 of the line (all loads first); the digit at sp+1 is patched per frame
            uint8_t c[4];
#pragma unroll
            for (int k = 0; k < 4; k++) {
                const u32 i = i0 + k;
                c[k] = '\n';
                if (i < sp) c[k] = hdr[i];
                else if (i == sp) c[k] = '_';
                else if (i > sp + 1 && i < H + 2) c[k] = hdr[i - 2];
            }
#pragma unroll
            for (int f = 0; f < 6; f++) {
                if (!((job.frame_mask >> f) & 1u)) continue;
                uint8_t *o = job.out + base[f] + i0;
#pragma unroll
                for (int k = 0; k < 4; k++)
                    if (i0 + k < H + 3) o[k] = (i0 + k == sp + 1) ? (uint8_t)('1' + f) : c[k];
            }
        }
    }
}

// =====================================================================================
// launch
// =====================================================================================
cudaError_t init_lines();
cudaError_t launch_lines(const Job &job, int sm_count, cudaStream_t stream);

cudaError_t init_short();
cudaError_t launch_short(const Job &job, int sm_count, cudaStream_t stream);

cudaError_t init_kernels() {
    cudaError_t e = init_lines();
    return e != cudaSuccess ? e : init_short();
}

// k_headers and k_short only need the record table, so they run on side streams next to k_lines:
// the persistent k_lines grid leaves room for their CTAs on every SM, and their launches overlap.
cudaError_t launch_emit(const Job &job, int sm_count, cudaStream_t stream, cudaEvent_t ev_mid, const EmitStreams &es) {
    cudaError_t e;
#define GTS_TRY(call) do { e = (call); if (e != cudaSuccess) return e; } while (0)
    // (A/B switch of the measurements in profiles/r2_notes.md: 0 = k_headers on a side stream next to
    // k_lines (the default), 1 = after k_lines, 2 = in front of k_lines, both on the main stream)
    static const int hdr_order = std::getenv("GTS_HDR_ORDER") ? std::atoi(std::getenv("GTS_HDR_ORDER")) : 0;
    GTS_TRY(cudaEventRecord(es.fork, stream));
    if (hdr_order == 0) {
        GTS_TRY(cudaStreamWaitEvent(es.side, es.fork, 0));
        k_headers<<<sm_count * GTS_HDR_BLOCKS, 128, 0, es.side>>>(job);
        GTS_TRY(cudaGetLastError());
        GTS_TRY(cudaEventRecord(es.join, es.side));
    } else if (hdr_order == 2) {
        k_headers<<<sm_count * GTS_HDR_BLOCKS, 128, 0, stream>>>(job);
        GTS_TRY(cudaGetLastError());
    }
    if (job.short_on) {
        GTS_TRY(cudaStreamWaitEvent(es.side2, es.fork, 0));
        GTS_TRY(launch_short(job, sm_count, es.side2));
        GTS_TRY(cudaEventRecord(es.join2, es.side2));
    }
    GTS_TRY(launch_lines(job, sm_count, stream));
    if (ev_mid) GTS_TRY(cudaEventRecord(ev_mid, stream));
    if (hdr_order == 0) {
        GTS_TRY(cudaStreamWaitEvent(stream, es.join, 0));
    } else if (hdr_order == 1) {
        k_headers<<<sm_count * GTS_HDR_BLOCKS, 128, 0, stream>>>(job);
        GTS_TRY(cudaGetLastError());
    }
    if (job.short_on) GTS_TRY(cudaStreamWaitEvent(stream, es.join2, 0));
#undef GTS_TRY
    return cudaSuccess;
}

}  // namespace gts
