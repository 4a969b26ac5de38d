// Shared between the tok3 host walk (tok3.cu) and the name-rebuild kernel (tok3_names.cu).
#pragma once
#include "hgpu_internal.h"

// enum name_type, tokenise_name3.c:121-122
enum { T_TYPE = 0, T_ALPHA, T_CHAR, T_DIGITS0, T_DZLEN, T_DUP, T_DIFF, T_DIGITS, T_DDELTA,
       T_DDELTA0, T_MATCH, T_NOP, T_END };
constexpr int TOK_MAX = 128;                       // MAX_TOKENS, tokenise_name3.c:115

struct Tok3Desc {                                  // one token stream ("descriptor", :143-148)
    uint64_t off;                                  // byte offset in the stream arena (16-byte aligned)
    uint32_t len;                                  // buf_a
    uint32_t synth;                                // 0: real bytes; else 0x100|type: [type, MATCH, MATCH, ...] (:1720-1729)
};

struct Tok3Block {
    uint64_t out_off;                              // names go to d_out + out_off
    uint64_t hist_off;                             // first uint2 of this block's history table
    uint64_t name_off;                             // first uint4 of this block's name table
    uint32_t out_cap;
    uint32_t desc_base;                            // first Tok3Desc of this block
    uint32_t max_tok;
    uint32_t nreads;                               // header field; the context holds nreads+1 names (:189-192)
    uint32_t ulen;                                 // header field
    uint32_t job0, njobs;                          // entropy-decoder jobs of this block
    int32_t  host_status;                          // framing already rejected on the host
};

// The general kernel: one warp per block, any number of token positions.  d_order[0..n): the blocks to
// run; max_ndesc: the largest max_tok*16 among them (sizes the per-warp shared-memory tables).
int hgpu_launch_tok3_names(hgpu_ctx *ctx, const Tok3Block *d_blocks, const uint32_t *d_order, uint32_t n, uint32_t max_ndesc,
                           const Tok3Desc *d_descs, const uint8_t *d_arena, const int32_t *d_job_status,
                           const uint32_t *d_job_got, const uint32_t *d_job_want, uint2 *d_hist, uint4 *d_names,
                           uint8_t *d_out, uint32_t *d_out_len, int32_t *d_status, cudaStream_t st);
// The common case, blocks with at most 16 token positions (max_tok <= 17): TWO blocks per warp, one per
// half-warp, sharing one instruction stream (tok3_names_h16.cu).
constexpr uint32_t TOK3_H16_MAX_TOK = 17;
int hgpu_launch_tok3_names_h16(hgpu_ctx *ctx, const Tok3Block *d_blocks, const uint32_t *d_order, uint32_t n,
                               const Tok3Desc *d_descs, const uint8_t *d_arena, const int32_t *d_job_status,
                               const uint32_t *d_job_got, const uint32_t *d_job_want, uint2 *d_hist, uint4 *d_names,
                               uint8_t *d_out, uint32_t *d_out_len, int32_t *d_status, cudaStream_t st);
