// Internal declarations shared by the kernels and the C-ABI layer of libhtsgpu.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include "../../include/htsgpu.h"

#define HGPU_WARP 32

struct hgpu_ctx {
    int device;
    int sm_count;
    cudaStream_t stream;          // default stream for _host entry points
    cudaStream_t copy_stream[2];  // H2D / D2H overlap in the pipelined host paths
    cudaEvent_t ev[8];
    // device scratch, grown on demand
    uint8_t *d_scratch;  size_t d_scratch_cap;     // rANS per-warp scratch
    uint8_t *d_mrec;     size_t d_mrec_cap;        // inflate match-record scratch
    uint8_t *d_bam;      size_t d_bam_cap;         // BAM index / scan scratch
    uint8_t *d_stage;    size_t d_stage_cap;     // device staging for _host entry points
    uint8_t *h_pinned;   size_t h_pinned_cap;    // pinned host staging
    uint32_t *d_counter;                          // 64 work-queue counters, handed out round-robin
    uint32_t next_counter;
};

// error plumbing (hgpu_api.cu)
void hgpu_set_error(const char *fmt, ...);
int  hgpu_check(cudaError_t e, const char *what);
void hgpu_count_launch(int n = 1);
int  hgpu_ensure_scratch(hgpu_ctx *ctx, size_t bytes);
int  hgpu_ensure_stage(hgpu_ctx *ctx, size_t bytes);
int  hgpu_ensure_mrec(hgpu_ctx *ctx, size_t bytes);
int  hgpu_ensure_bam(hgpu_ctx *ctx, size_t bytes);
int  hgpu_ensure_pinned(hgpu_ctx *ctx, size_t bytes);
// process-wide context of the reference-named shims (one batch at a time); call between lock/unlock
void hgpu_shim_lock();
void hgpu_shim_unlock();
hgpu_ctx *hgpu_shim_ctx();
// a zeroed (stream-ordered) work counter; slots rotate so launches in flight on different streams never share one
uint32_t *hgpu_take_counter(hgpu_ctx *ctx, cudaStream_t st);

// kernel launchers (one per .cu)
int hgpu_launch_rans_nx16(hgpu_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off,
                          const uint32_t *d_in_len, uint32_t n, uint8_t *d_out,
                          const uint64_t *d_out_off, const uint32_t *d_out_len,
                          uint32_t *d_got_len, int32_t *d_status, uint32_t max_out_len,
                          cudaStream_t st);
int hgpu_launch_bgzf_inflate(hgpu_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off,
                             const uint32_t *d_in_len, uint32_t n, uint8_t *d_out,
                             const uint64_t *d_out_off, const uint32_t *d_out_cap,
                             uint32_t *d_out_len, int32_t *d_status, cudaStream_t st);
int hgpu_launch_gzip_inflate(hgpu_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off, const uint32_t *d_in_len, uint32_t n,
                             uint8_t *d_out, const uint64_t *d_out_off, const uint32_t *d_out_cap, uint32_t *d_out_len,
                             int32_t *d_status, cudaStream_t st);
int hgpu_launch_crc32(hgpu_ctx *ctx, const uint8_t *d_buf, size_t len, uint32_t *d_partial,
                      uint32_t *h_result, uint32_t crc0, cudaStream_t st);

int hgpu_launch_crc32_batch(hgpu_ctx *ctx, const uint8_t *d_buf, const uint64_t *d_off, const uint32_t *d_len, uint32_t n,
                            uint32_t *d_crc, cudaStream_t st);

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t hgpu_lanemask_lt()
{
    uint32_t m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ uint32_t hgpu_lane() { return threadIdx.x & 31; }
#endif
