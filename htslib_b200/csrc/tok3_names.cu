// tok3 name rebuild: decode_name (htscodecs/htscodecs/tokenise_name3.c:1023-1210) for a batch of name
// blocks, one WARP per block, one LANE per token position.
//
// Why this shape: within one name, token position k reads only streams (k<<4 | type) and the k-th token
// record of the name it diffs against, so positions are independent and map to lanes; the fragment each
// lane produces is placed by a warp prefix sum.  Names of a block form a serial chain (name n needs the
// finished name n-dist), so a block cannot use more than one warp; batches supply the parallelism.
//
// Latency is what bounds a warp here (a handful of dependent loads per name), so everything the next
// name is likely to touch stays on chip:
//   * descriptors and stream cursors of the block: shared memory;
//   * the token-type stream of each lane (one byte per name): an 8-byte register window;
//   * the previous name's token records (the usual diff target): registers; any older name: the
//     history table in HBM;
//   * the previous name's bytes (MATCH of a string token, duplicates): a double-buffered 256-byte
//     shared-memory copy, written out to HBM with coalesced stores once the name is complete.
#include "tok3_internal.h"

namespace {

constexpr int NAME_BUF = 256;                      // bytes of each of the two per-warp name buffers
#ifndef TOK3_WARPS
#define TOK3_WARPS 2
#endif
#ifndef TOK3_MINB
#define TOK3_MINB 16                               // 64 registers: 32 warps per SM; the kernel is issue-bound, more warps win
#endif
constexpr int WARPS = TOK3_WARPS;                  // name blocks per CTA

__device__ __forceinline__ Tok3Desc as_desc(uint4 v)
{
    Tok3Desc d;
    d.off = (uint64_t)v.x | ((uint64_t)v.y << 32);
    d.len = v.z;
    d.synth = v.w;
    return d;
}
__device__ __forceinline__ int desc_byte(const uint8_t *arena, const Tok3Desc &d, uint32_t pos)
{
    if (d.synth) return pos == 0 ? (int)(d.synth & 0xff) : T_MATCH;
    return arena[d.off + pos];
}
// 32-bit little-endian read at a cursor; streams start 16-byte aligned, cursors of integer streams stay 4-aligned
__device__ __forceinline__ uint32_t desc_u32(const uint8_t *arena, const Tok3Desc &d, uint32_t pos)
{
    if (!d.synth && (pos & 3) == 0) return *reinterpret_cast<const uint32_t *>(arena + d.off + pos);
    return (uint32_t)desc_byte(arena, d, pos) | (uint32_t)desc_byte(arena, d, pos + 1) << 8 |
           (uint32_t)desc_byte(arena, d, pos + 2) << 16 | (uint32_t)desc_byte(arena, d, pos + 3) << 24;
}
// one byte through an 8-byte register window over a (real) stream; arena slots are padded to 16 bytes
__device__ __forceinline__ int window_byte(const uint8_t *arena, const Tok3Desc &d, uint32_t pos, uint64_t &w, uint32_t &wbase)
{
    if (d.synth) return pos == 0 ? (int)(d.synth & 0xff) : T_MATCH;
    if ((pos & ~7u) != wbase) { wbase = pos & ~7u; w = *reinterpret_cast<const uint64_t *>(arena + d.off + wbase); }
    return (int)((w >> ((pos & 7u) * 8)) & 0xff);
}

__constant__ uint32_t c_p10[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};

__global__ void __launch_bounds__(32 * WARPS, TOK3_MINB) tok3_names_kernel(const Tok3Block *blocks, const uint32_t *order, uint32_t nblocks,
        uint32_t max_ndesc, const Tok3Desc *descs, const uint8_t *arena, const int32_t *job_status, const uint32_t *job_got,
        const uint32_t *job_want, uint2 *hist_all, uint4 *names_all, uint8_t *out, uint32_t *out_len, int32_t *status)
{
    extern __shared__ uint4 smem4[];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t slot = blockIdx.x * WARPS + warp;
    if (slot >= nblocks) return;                                   // warps never synchronise with each other
    const uint32_t b = order[slot];                                // the blocks this kernel was given (those with > 16 token positions)
    uint4 *dsc = smem4 + (size_t)warp * max_ndesc;
    uint32_t *cur = reinterpret_cast<uint32_t *>(smem4 + (size_t)WARPS * max_ndesc) + (size_t)warp * max_ndesc;
    uint8_t *nbuf = reinterpret_cast<uint8_t *>(reinterpret_cast<uint32_t *>(smem4 + (size_t)WARPS * max_ndesc) + (size_t)WARPS * max_ndesc) +
                    (size_t)warp * 2 * NAME_BUF;

    const Tok3Block B = blocks[b];
    if (B.host_status) { if (lane == 0) { status[b] = B.host_status; out_len[b] = 0; } return; }

    // every token stream must have decoded to exactly the size its header announced (:1789-1793)
    bool bad = false;
    for (uint32_t j = lane; j < B.njobs; j += 32)
        bad |= job_status[B.job0 + j] != HGPU_OK || job_got[B.job0 + j] != job_want[B.job0 + j];
    if (__any_sync(0xffffffffu, bad)) { if (lane == 0) { status[b] = HGPU_TOK3_ERR; out_len[b] = 0; } return; }

    const uint32_t ndesc = B.max_tok * 16;
    for (uint32_t i = lane; i < ndesc; i += 32) {
        dsc[i] = __ldg(reinterpret_cast<const uint4 *>(descs + B.desc_base + i));
        cur[i] = 0;
    }
    __syncwarp();

    uint2 *H = hist_all + B.hist_off;                              // [name][max_tok] {val, type<<28 | aux}
    uint4 *NM = names_all + B.name_off;                            // {offset, ntok, history row, 0}
    uint8_t *O = out + B.out_off;
    const uint32_t kmax = B.max_tok < (uint32_t)TOK_MAX ? B.max_tok : (uint32_t)TOK_MAX;

    int64_t room = (int64_t)B.ulen + 1024;                         // name_len of decode_name (:1810-1818)
    uint64_t at = 0;
    uint32_t cnum = 0;
    int result = 0;                                                // 0 running, 1 finished, -1 error
    if ((int64_t)B.out_cap < room) result = -1;                    // caller's slot is smaller than the reference's buffer

    // per-warp caches
    const Tok3Desc d0 = as_desc(dsc[0]);
    uint64_t w0 = 0; uint32_t w0base = 0xffffffffu;                // window over token 0's type stream
    const uint32_t k1 = 1 + lane;                                  // this lane's token position in the first chunk
    const Tok3Desc dt1 = k1 < kmax ? as_desc(dsc[k1 << 4]) : Tok3Desc{0, 0, 0};
    uint64_t w1 = 0; uint32_t w1base = 0xffffffffu;                // window over this lane's type stream
    uint32_t pv_val = 0, pv_ta = 0, reg_row = 0xffffffffu;         // token record k1 of history row reg_row
    uint32_t last_cnum = 0xffffffffu;
    uint4 last_nm = make_uint4(0, 0, 0, 0);                        // NM[last_cnum]
    uint32_t cb = 0, buf_cnum = 0xffffffffu, buf_fill = 0;         // nbuf[cb^1] holds bytes [0, buf_fill) of name buf_cnum

    while (result == 0) {
        // ---- token 0: which earlier name to diff against (uniform; lane 0 owns these cursors)
        uint32_t c0 = __shfl_sync(0xffffffffu, cur[0], 0);
        int t0 = c0 < d0.len ? window_byte(arena, d0, c0, w0, w0base) : -1;
        if (lane == 0 && t0 >= 0) cur[0] = c0 + 1;
        if (cnum > B.nreads) { result = -1; break; }               // cnum >= max_names (:1028)
        if (t0 < 0 || (uint32_t)t0 >= ndesc) { result = 1; break; }
        __syncwarp();
        const Tok3Desc dd = as_desc(dsc[t0]);
        const uint32_t cd = __shfl_sync(0xffffffffu, cur[t0], 0);
        if ((uint64_t)cd + 4 > dd.len) { result = -1; break; }
        const uint32_t dist = desc_u32(arena, dd, cd);
        if (lane == 0) cur[t0] = cd + 4;
        __syncwarp();
        if (dist > cnum) { result = -1; break; }
        const uint32_t pnum = cnum - dist;
        const uint4 P = pnum == last_cnum ? last_nm : NM[pnum];    // only meaningful when pnum < cnum
        uint8_t *name = O + at;
        uint8_t *mybuf = nbuf + cb * NAME_BUF;
        const uint8_t *pvbuf = nbuf + (cb ^ 1) * NAME_BUF;

        if (t0 == T_DUP) {
            if (pnum == cnum) { result = -1; break; }
            // strcpy semantics: up to the first NUL of the earlier name (:1043-1045); that NUL exists
            // because every finished name ends in one this kernel wrote
            const bool in_buf = pnum == buf_cnum;
            const uint8_t *src = O + P.x;
            uint32_t l = 0;
            bool stop = false, over = false;
            while (!stop) {
                const uint32_t i = l + lane;
                const uint8_t ch = (in_buf && i < buf_fill) ? pvbuf[i] : src[i];
                const uint32_t z = __ballot_sync(0xffffffffu, ch == 0);
                const uint32_t n = z ? (uint32_t)__ffs(z) - 1 : 32u;
                if ((int64_t)(l + n) + 1 >= room) { over = true; break; }
                if (lane < n) { name[i] = ch; if (i < NAME_BUF) mybuf[i] = ch; }
                l += n;
                stop = z != 0;
            }
            if (over) { result = -1; break; }
            if (lane == 0) { name[l] = 0; if (l < NAME_BUF) mybuf[l] = 0; }
            last_nm = make_uint4((uint32_t)at, P.y, P.z, 0);
            if (lane == 0) NM[cnum] = last_nm;
            last_cnum = cnum;
            buf_cnum = cnum; buf_fill = l + 1 < (uint32_t)NAME_BUF ? l + 1 : (uint32_t)NAME_BUF; cb ^= 1;
            at += l + 1; room -= l + 1;
            cnum++;
            __syncwarp();
            continue;
        }

        const uint32_t pntok = pnum == cnum ? 0 : P.y;             // last_ntok is 0 while a name is in flight (:1071)
        const uint2 *HP = H + (uint64_t)P.z * B.max_tok;
        uint2 *HC = H + (uint64_t)cnum * B.max_tok;
        const bool prev_in_regs = P.z == reg_row;
        const bool prev_in_buf = pnum == buf_cnum;
        uint32_t len = 0, ntok = 0, fill = 0;
        bool ended = false, err = false;

        for (uint32_t base = 1; base < kmax && !ended && !err; base += 32) {
            const uint32_t k = base + lane;
            const bool active = k < kmax, first = base == 1;
            const uint4 *S = dsc + (k << 4);
            uint32_t *C = cur + (k << 4);
            int tok = -1;
            uint32_t ct = 0;
            if (active) {
                ct = C[T_TYPE];
                if (first) { if (ct < dt1.len) tok = window_byte(arena, dt1, ct, w1, w1base); }
                else { const Tok3Desc dt = as_desc(S[T_TYPE]); if (ct < dt.len) tok = desc_byte(arena, dt, ct); }
            }
            const bool payload = tok == T_ALPHA || tok == T_CHAR || tok == T_DIGITS0 || tok == T_DIGITS ||
                                 tok == T_DDELTA || tok == T_DDELTA0 || tok == T_MATCH || tok == T_NOP;
            const uint32_t endmask = __ballot_sync(0xffffffffu, active && !payload);
            const uint32_t e = endmask ? (uint32_t)__ffs(endmask) - 1 : 32u;   // first END / dry type stream
            const bool mine = active && lane <= e;

            // ---- operands.  All stream reads happen up front, outside any per-type branch, so that lanes
            // holding different token types overlap their load latencies instead of serialising them.
            const bool body = mine && lane != e;
            const bool hasq = body && k < pntok;
            const uint32_t cls = body ? (uint32_t)tok : 0u;
            // byte operand: CHAR, DDELTA, DDELTA0 read their own stream, DIGITS0 reads its width from DZLEN;
            // word operand: DIGITS and DIGITS0.  One nibble per token type.
            const uint32_t bs = (uint32_t)(0x0000009800004200ull >> (4 * cls)) & 15u;
            const uint32_t ws = (uint32_t)(0x0000000070003000ull >> (4 * cls)) & 15u;
            if (mine && tok >= 0) C[T_TYPE] = ct + 1;                             // decode_token_type consumed it
            uint2 q = make_uint2(0, 0);
            if (hasq) q = (first && prev_in_regs) ? make_uint2(pv_val, pv_ta) : HP[k];
            bool lerr = false;
            uint32_t bval = 0, wval = 0;
            if (bs) {
                const Tok3Desc db = as_desc(S[bs]);
                const uint32_t c = C[bs];
                if (c >= db.len) lerr = true;
                else { bval = (uint32_t)desc_byte(arena, db, c); C[bs] = c + 1; }
            }
            if (ws) {
                const Tok3Desc dw = as_desc(S[ws]);
                const uint32_t c = C[ws];
                if ((uint64_t)c + 4 > dw.len) lerr = true;
                else { wval = desc_u32(arena, dw, c); C[ws] = c + 4; }
            }
            const uint32_t qtype = q.y >> 28, qaux = q.y & 0x0fffffffu;

            // ---- classify: a single byte, a number, or a copied string; and the token record to keep.
            // MATCH adopts the type of the earlier name's token (:1133-1183), so it shares the arms below.
            uint32_t flen = 0, need = 0, rtype = T_NOP, rval = 0, raux = 0;
            uint32_t chr = 0, num = 0, numw = 0, numkind = 0;                     // numkind 1: variable width, 2: fixed width numw
            bool is_chr = false, is_copy = false, alpha_open = false, synth_alpha = false;
            const uint8_t *csrc = nullptr;
            Tok3Desc ds = Tok3Desc{0, 0, 0};
            uint32_t spos = 0;
            if (mine && lane == e) {                                              // N_END (:1186-1204)
                is_chr = true; flen = 1; need = 1; rtype = T_END;
            } else if (body && !lerr) {
                const uint32_t t = (uint32_t)tok;
                if ((t == T_MATCH || t == T_DDELTA || t == T_DDELTA0) && !hasq) lerr = true;
                else {
                    const uint32_t et = t == T_MATCH ? (0x100u | qtype) : t;
                    switch (et) {
                    case T_CHAR: case 0x100 | T_CHAR:
                        chr = et == T_CHAR ? bval : (q.x & 0xff);
                        rval = et == T_CHAR ? (uint32_t)(int32_t)(int8_t)bval : q.x;   // token_int = (char) (:1078)
                        is_chr = true; flen = 1; need = 1; rtype = T_CHAR;
                        break;
                    case T_DIGITS: case T_DDELTA: case 0x100 | T_DIGITS:
                        num = et == T_DIGITS ? wval : et == T_DDELTA ? bval + q.x : q.x;
                        numkind = 1; need = 20; rtype = T_DIGITS; rval = num;         // flen once the digits are counted
                        break;
                    case T_DIGITS0: case T_DDELTA0: case 0x100 | T_DIGITS0:
                        num = et == T_DIGITS0 ? wval : et == T_DDELTA0 ? bval + q.x : q.x;
                        numw = et == T_DIGITS0 ? bval : qaux;
                        need = et == T_DIGITS0 ? 20 + numw : et == T_DDELTA0 ? numw + 1 : (numw ? numw : 0x80000000u);
                        numkind = 2; flen = numw; rtype = T_DIGITS0; rval = num; raux = numw;
                        break;
                    case T_ALPHA: {
                        ds = as_desc(S[T_ALPHA]);
                        const uint32_t c = C[T_ALPHA];
                        if (c >= ds.len) { lerr = true; break; }
                        uint32_t n = 0;                                           // bytes consumed incl. the NUL
                        int ch;
                        do { ch = desc_byte(arena, ds, c + n); n++; } while (ch && c + n < ds.len);
                        C[T_ALPHA] = c + n;
                        flen = n - 1;                                             // a missing NUL drops the last char (:432-437)
                        need = n; alpha_open = true;                              // needs n <= room - len
                        spos = c;
                        if (ds.synth) synth_alpha = true; else { is_copy = true; csrc = arena + ds.off + c; }
                        rtype = T_ALPHA; rval = flen;
                        break; }
                    case 0x100 | T_ALPHA:
                        if ((int32_t)q.x < 0) { lerr = true; break; }
                        flen = q.x; need = q.x ? q.x : 0x80000000u;               // "len + 0 >= room" still applies: marker
                        is_copy = true;
                        csrc = (prev_in_buf && qaux + q.x <= buf_fill) ? pvbuf + qaux : O + P.x + qaux;
                        rtype = T_ALPHA; rval = q.x;
                        break;
                    case T_NOP:
                        rtype = T_NOP;
                        break;
                    default:                                                       // MATCH of a token that is none of the above (:1180)
                        lerr = true;
                    }
                }
            }

            // ---- decimal text of every numeric lane at once: straight-line, no per-digit loop.  The string is
            // built zero-padded to 10 characters and the wanted suffix shifted down into (slo, shi).
            uint64_t slo = chr;                                                   // bytes 0..7 of the register string
            uint32_t shi = 0, rlen = is_chr ? 1u : 0u;                            // bytes 8..9; bytes to store from it
            if (__any_sync(0xffffffffu, numkind != 0)) {
                const uint32_t hi5 = num / 100000u, lo5 = num - hi5 * 100000u;
                uint32_t a = hi5;
                const uint32_t d9 = a / 10000u; a -= d9 * 10000u;
                const uint32_t d8 = a / 1000u;  a -= d8 * 1000u;
                const uint32_t d7 = a / 100u;   a -= d7 * 100u;
                const uint32_t d6 = a / 10u;    const uint32_t d5 = a - d6 * 10u;
                a = lo5;
                const uint32_t d4 = a / 10000u; a -= d4 * 10000u;
                const uint32_t d3 = a / 1000u;  a -= d3 * 1000u;
                const uint32_t d2 = a / 100u;   a -= d2 * 100u;
                const uint32_t d1 = a / 10u;    const uint32_t d0 = a - d1 * 10u;
                const uint32_t w0_ = (d9 | d8 << 8 | d7 << 16 | d6 << 24) + 0x30303030u;
                const uint32_t w1_ = (d5 | d4 << 8 | d3 << 16 | d2 << 24) + 0x30303030u;
                const uint64_t lo = (uint64_t)w0_ | (uint64_t)w1_ << 32;
                const uint32_t hi = (d1 | d0 << 8) + 0x3030u;
                const uint32_t sig = num >= 100000u ? (num >= 10000000u ? (num >= 1000000000u ? 10u : num >= 100000000u ? 9u : 8u)
                                                                       : (num >= 1000000u ? 7u : 6u))
                                                    : (num >= 100u ? (num >= 10000u ? 5u : num >= 1000u ? 4u : 3u)
                                                                   : (num >= 10u ? 2u : num >= 1u ? 1u : 0u));
                if (numkind == 1) flen = sig;                                      // append_uint32_var (:249-286): nothing for 0
                // append_uint32_fixed (:233-247) writes w chars for w in 1..9 and nothing otherwise (flen stays w)
                const uint32_t wr = numkind == 1 ? sig : numkind == 2 ? ((numw >= 1 && numw <= 9) ? numw : 0u) : 0u;
                if (numkind) {
                    const uint32_t sh = 10u - wr;                                  // drop the leading sh characters
                    if (sh >= 8u) { slo = (uint64_t)(hi >> (8u * (sh - 8u))); shi = 0; }
                    else if (sh == 0u) { slo = lo; shi = hi; }
                    else { slo = (lo >> (8u * sh)) | ((uint64_t)hi << (64u - 8u * sh)); shi = hi >> (8u * sh); }
                    // the fixed form stores its leading digit unreduced: a value too wide for w digits shows there
                    if (numkind == 2 && wr && num >= c_p10[wr]) slo = (slo & ~0xffull) | (uint8_t)(num / c_p10[wr - 1] + '0');
                    rlen = wr;
                }
            }

            // exclusive prefix of the fragment lengths -> where each lane writes
            uint32_t incl = flen;
            #pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                const uint32_t t = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= (uint32_t)s) incl += t;
            }
            const uint32_t off = len + incl - flen;
            // the reference's "len + need >= name_len" guards, evaluated with this lane's own len
            if (mine && !lerr) {
                if (alpha_open) { if ((int64_t)need > room - (int64_t)off) lerr = true; }
                else if (need == 0x80000000u) { if ((int64_t)off >= room) lerr = true; }
                else if (need && (int64_t)off + (int64_t)need >= room) lerr = true;
                if (rtype == T_ALPHA) raux = off;                                  // token_str = offset in the name
                if (raux >> 28) lerr = true;                                       // beyond the packed field (names of 256 MB)
            }
            if (__any_sync(0xffffffffu, lerr)) { err = true; break; }

            // fragments that fit the shared buffer are staged there (and flushed below), the rest go straight out
            const bool fits = (uint64_t)off + flen <= (uint64_t)NAME_BUF;
            uint8_t *dst = fits ? mybuf + off : name + off;
            if (!mine) rlen = 0;
            // (1) register strings: single bytes and numbers, at most 10 bytes per lane
            const uint32_t maxr = __reduce_max_sync(0xffffffffu, rlen);
            for (uint32_t i = 0; i < maxr; i++) {
                if (i < rlen) dst[i] = (uint8_t)slo;
                slo = (slo >> 8) | ((uint64_t)shi << 56); shi >>= 8;
            }
            // (2) copied strings (from a token stream, or from the earlier name): the whole warp moves each one
            uint32_t cmask = __ballot_sync(0xffffffffu, mine && is_copy && flen > 0);
            while (cmask) {
                const int j = __ffs(cmask) - 1;
                cmask &= cmask - 1;
                const uint8_t *s8 = reinterpret_cast<const uint8_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(csrc), j));
                uint8_t *d8 = reinterpret_cast<uint8_t *>(__shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(dst), j));
                const uint32_t L = __shfl_sync(0xffffffffu, flen, j);
                for (uint32_t i = lane; i < L; i += 32) d8[i] = s8[i];
            }
            if (mine) {
                if (synth_alpha) for (uint32_t i = 0; i < flen; i++) dst[i] = (uint8_t)desc_byte(arena, ds, spos + i);
                const uint32_t rec = rtype << 28 | raux;
                HC[k] = make_uint2(rval, rec);
                if (first) { pv_val = rval; pv_ta = rec; }
                if (fits && flen) fill = off + flen;                               // offsets only grow: the last fitting end
            }
            len += __shfl_sync(0xffffffffu, incl, 31);
            if (e < 32) { ended = true; ntok = base + e; }
        }
        if (err || !ended) { result = -1; break; }
        // flush the staged prefix with coalesced stores
        fill = __reduce_max_sync(0xffffffffu, fill);
        __syncwarp();
        for (uint32_t i = lane; i < fill; i += 32) name[i] = mybuf[i];
        last_nm = make_uint4((uint32_t)at, ntok, cnum, 0);
        if (lane == 0) NM[cnum] = last_nm;
        last_cnum = cnum; reg_row = cnum;
        buf_cnum = cnum; buf_fill = fill; cb ^= 1;
        at += len; room -= len;
        cnum++;
        __syncwarp();                                                              // history and name bytes visible to the next name
    }

    if (lane == 0) {
        status[b] = result == 1 ? HGPU_OK : HGPU_TOK3_ERR;
        out_len[b] = result == 1 ? (uint32_t)at : 0;
    }
}

}  // namespace

int hgpu_launch_tok3_names(hgpu_ctx *ctx, const Tok3Block *d_blocks, const uint32_t *d_order, uint32_t n, uint32_t max_ndesc,
                           const Tok3Desc *d_descs, const uint8_t *d_arena, const int32_t *d_job_status,
                           const uint32_t *d_job_got, const uint32_t *d_job_want, uint2 *d_hist, uint4 *d_names,
                           uint8_t *d_out, uint32_t *d_out_len, int32_t *d_status, cudaStream_t st)
{
    (void)ctx;
    if (n == 0) return HGPU_OK;
    if (max_ndesc < 16) max_ndesc = 16;
    const size_t smem = (size_t)WARPS * ((size_t)max_ndesc * (sizeof(uint4) + sizeof(uint32_t)) + 2 * NAME_BUF);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        if (hgpu_check(cudaFuncSetAttribute(tok3_names_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "tok3 smem attr"))
            return HGPU_ERR_CUDA;
        smem_set = smem;
    }
    tok3_names_kernel<<<(n + WARPS - 1) / WARPS, 32 * WARPS, smem, st>>>(d_blocks, d_order, n, max_ndesc, d_descs, d_arena, d_job_status,
                                                                        d_job_got, d_job_want, d_hist, d_names, d_out, d_out_len, d_status);
    hgpu_count_launch();
    return hgpu_check(cudaGetLastError(), "tok3_names_kernel");
}
