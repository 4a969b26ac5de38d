// Warp-level pieces of the htscodecs byte transforms shared by xform.cu (the hts_pack / hts_rle_* shims) and
// rans_nx16_enc.cu (PACK / RLE inside the rANS Nx16 container).
#pragma once
#include "hgpu_internal.h"

__device__ __forceinline__ int put_var(uint8_t *p, uint32_t v)       // var_put_u32, varint.h:206 (7 bits per byte, big end first)
{
    int n = 1;
    while (n < 5 && (v >> (7 * n))) n++;
    for (int k = n - 1; k >= 0; k--) *p++ = (uint8_t)(((v >> (7 * k)) & 0x7f) | (k ? 0x80 : 0));
    return n;
}

// RLE encode (rle.c:100-140) by one warp: 32 bytes per round; a byte is a literal unless it continues a run of a
// symbol in the set; lane 0 lays down the run lengths (varints) in order, carrying an open run across rounds.
// inset[256]: 1 for the symbols that carry run lengths.  Returns the literal count in nlit and the run bytes in nrun
// (the same on every lane).
__device__ inline void warp_rle_encode(const uint8_t *d, uint64_t len, const uint8_t *inset, uint8_t *lit, uint8_t *run,
                                       uint64_t &nlit, uint64_t &nrun)
{
    const uint32_t lane = threadIdx.x & 31;
    uint64_t k = 0, j = 0;
    bool open = false;                 // a run of a set symbol is still growing
    uint32_t ocount = 0;               // its length minus one so far
    uint32_t lastb = 256;
    for (uint64_t base = 0; base < len; base += 32) {
        const uint64_t i = base + lane;
        const bool valid = i < len;
        const uint32_t b = valid ? d[i] : 256u;
        uint32_t prev = __shfl_up_sync(0xffffffffu, b, 1);
        if (lane == 0) prev = lastb;
        const bool set = valid && inset[b];
        const bool cont = set && b == prev;
        const bool head = valid && !cont;
        const uint32_t hb = __ballot_sync(0xffffffffu, head), vb = __ballot_sync(0xffffffffu, valid), sb = __ballot_sync(0xffffffffu, set);
        if (head) lit[k + __popc(hb & hgpu_lanemask_lt())] = (uint8_t)b;
        if (lane == 0) {
            const int nvalid = __popc(vb);
            // the open run from the previous round grows by the leading continuation lanes
            if (open) {
                const int lead = hb ? __ffs(hb) - 1 : nvalid;
                ocount += (uint32_t)lead;
                if (hb) { j += put_var(run + j, ocount); open = false; }
            }
            uint32_t mm = hb;
            while (mm) {
                const int h = __ffs(mm) - 1;
                mm &= mm - 1;
                if (!((sb >> h) & 1u)) continue;                    // a literal of a symbol outside the set carries no run
                const int nxt = mm ? __ffs(mm) - 1 : nvalid;        // next literal, or the end of this round
                const uint32_t rl = (uint32_t)(nxt - h - 1);
                if (mm || base + 32 >= len) j += put_var(run + j, rl);      // closed inside the round, or the data ends here
                else { open = true; ocount = rl; }
            }
        }
        j = __shfl_sync(0xffffffffu, (unsigned long long)j, 0);
        open = __shfl_sync(0xffffffffu, (int)open, 0);
        ocount = __shfl_sync(0xffffffffu, ocount, 0);
        k += __popc(hb);
        lastb = __shfl_sync(0xffffffffu, b, 31);
    }
    if (open) { if (lane == 0) j += put_var(run + j, ocount); j = __shfl_sync(0xffffffffu, (unsigned long long)j, 0); }
    nlit = k; nrun = j;
}
