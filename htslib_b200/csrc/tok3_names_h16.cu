// tok3 name rebuild, common case: blocks with at most 16 token positions (max_tok <= 17) — TWO blocks
// per warp, one per half-warp.
//
// tok3_names_kernel (tok3_names.cu) is bound by instruction issue, and a typical name (Illumina: 9-15
// tokens) leaves half of its lanes idle.  Here lanes 0-15 rebuild one block and lanes 16-31 another with
// the SAME instruction stream, so the per-name instruction cost is shared by two names.  Everything that
// was warp-uniform there is half-uniform here (each half carries its own block state, cursors, name
// buffers in shared memory); warp collectives are issued by the full warp and each half reads its own
// 16 bits; nothing returns or breaks early — a half that is finished (or failed) just stops being
// `live` — so the two chains never disturb each other's control flow.  Semantics per block are those of
// decode_name (tokenise_name3.c:1023-1210), statement for statement as in tok3_names.cu.
#include "tok3_internal.h"

namespace {

constexpr int NAME_BUF = 256;                      // bytes of each of the two per-half name buffers
constexpr uint32_t NDESC = TOK3_H16_MAX_TOK * 16;  // descriptor slots per block
constexpr int WARPS = 2;                           // warps per CTA: four blocks
constexpr uint32_t FULL = 0xffffffffu;
constexpr uint32_t SYNTH = 0xfffffff0u;            // D8.offu >= SYNTH: implied [type, MATCH, MATCH, ...] stream, type in the low nibble

struct D8 { uint32_t offu, len; };                 // shared-memory descriptor: arena offset in 16-byte units, length

__device__ __forceinline__ int d_byte(const uint8_t *arena, D8 d, uint32_t pos)
{
    if (d.offu >= SYNTH) return pos == 0 ? (int)(d.offu & 15u) : T_MATCH;
    return arena[(uint64_t)d.offu * 16 + pos];
}
__device__ __forceinline__ uint32_t d_u32(const uint8_t *arena, D8 d, uint32_t pos)
{
    if (d.offu < SYNTH && (pos & 3) == 0) return *reinterpret_cast<const uint32_t *>(arena + (uint64_t)d.offu * 16 + pos);
    return (uint32_t)d_byte(arena, d, pos) | (uint32_t)d_byte(arena, d, pos + 1) << 8 |
           (uint32_t)d_byte(arena, d, pos + 2) << 16 | (uint32_t)d_byte(arena, d, pos + 3) << 24;
}
__device__ __forceinline__ int d_window(const uint8_t *arena, D8 d, uint32_t pos, uint64_t &w, uint32_t &wbase)
{
    if (d.offu >= SYNTH) return pos == 0 ? (int)(d.offu & 15u) : T_MATCH;
    if ((pos & ~7u) != wbase) { wbase = pos & ~7u; w = *reinterpret_cast<const uint64_t *>(arena + (uint64_t)d.offu * 16 + wbase); }
    return (int)((w >> ((pos & 7u) * 8)) & 0xff);
}

__constant__ uint32_t c_p10[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};

__global__ void __launch_bounds__(32 * WARPS, 15) tok3_names_h16_kernel(const Tok3Block *blocks, const uint32_t *order, uint32_t nblocks,
        const Tok3Desc *descs, const uint8_t *arena, const int32_t *job_status, const uint32_t *job_got,
        const uint32_t *job_want, uint2 *hist_all, uint4 *names_all, uint8_t *out, uint32_t *out_len, int32_t *status)
{
    __shared__ D8 s_dsc[WARPS * 2][NDESC];
    __shared__ uint32_t s_cur[WARPS * 2][NDESC];
    __shared__ uint8_t s_nbuf[WARPS * 2][2 * NAME_BUF];
    const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t half = lane >> 4, sub = lane & 15, hshift = half * 16;
    const uint32_t slot = (blockIdx.x * WARPS + warp) * 2 + half;
    const bool has = slot < nblocks;
    if (!__any_sync(FULL, has)) return;                             // whole warp without work
    const uint32_t b = has ? order[slot] : 0;
    D8 *dsc = s_dsc[warp * 2 + half];
    uint32_t *cur = s_cur[warp * 2 + half];
    uint8_t *nbuf = s_nbuf[warp * 2 + half];

    Tok3Block B;
    if (has) B = blocks[b];
    else { B.host_status = 1; B.njobs = 0; B.max_tok = 1; B.nreads = 0; B.ulen = 0; B.out_cap = 0; B.desc_base = 0; B.job0 = 0;
           B.hist_off = 0; B.name_off = 0; B.out_off = 0; }
    int result = 0;                                                 // 0 running, 1 finished, -1 error, 2 nothing to do / host verdict
    if (B.host_status) result = 2;
    {   // every token stream must have decoded to exactly the size its header announced (:1789-1793)
        bool bad = false;
        if (result == 0)
            for (uint32_t j = sub; j < B.njobs; j += 16)
                bad |= job_status[B.job0 + j] != HGPU_OK || job_got[B.job0 + j] != job_want[B.job0 + j];
        if ((__ballot_sync(FULL, bad) >> hshift) & 0xffffu) result = -1;
    }
    const uint32_t ndesc = B.max_tok * 16;                          // <= NDESC for the blocks given to this kernel
    if (result == 0) {
        for (uint32_t i = sub; i < ndesc; i += 16) {
            const uint4 v = __ldg(reinterpret_cast<const uint4 *>(descs + B.desc_base + i));
            D8 d;
            d.offu = v.w ? (SYNTH | (v.w & 15u)) : (uint32_t)((((uint64_t)v.y << 32) | v.x) >> 4);
            d.len = v.z;
            dsc[i] = d;
            cur[i] = 0;
        }
    }
    __syncwarp();

    uint2 *H = hist_all + B.hist_off;                               // [name][max_tok] {val, type<<28 | aux}
    uint4 *NM = names_all + B.name_off;                             // {offset, ntok, history row, 0}
    uint8_t *O = out + B.out_off;
    const uint32_t kmax = B.max_tok;                                // token positions 1 .. kmax-1, one per lane of the half

    int64_t room = (int64_t)B.ulen + 1024;                          // name_len of decode_name (:1810-1818)
    uint64_t at = 0;
    uint32_t cnum = 0;
    if (result == 0 && (int64_t)B.out_cap < room) result = -1;      // caller's slot is smaller than the reference's buffer

    // per-half caches (identical in the 16 lanes of a half unless noted)
    const D8 d0 = result == 0 ? dsc[0] : D8{0, 0};
    uint64_t w0 = 0; uint32_t w0base = 0xffffffffu;                 // window over token 0's type stream
    const uint32_t k = 1 + sub;                                     // this lane's token position
    const D8 dt1 = (result == 0 && k < kmax) ? dsc[k << 4] : D8{0, 0};
    uint64_t w1 = 0; uint32_t w1base = 0xffffffffu;                 // per lane: window over its type stream
    uint32_t pv_val = 0, pv_ta = 0, reg_row = 0xffffffffu;          // per lane: token record k of history row reg_row
    uint32_t last_cnum = 0xffffffffu;
    uint4 last_nm = make_uint4(0, 0, 0, 0);                         // NM[last_cnum]
    uint32_t cb = 0, buf_cnum = 0xffffffffu, buf_fill = 0;          // nbuf[cb^1] holds bytes [0, buf_fill) of name buf_cnum
    const D8 *S = dsc + (k << 4);                                   // only dereferenced when k < kmax
    uint32_t *C = cur + (k << 4);

    while (__any_sync(FULL, result == 0)) {
        bool live = result == 0;
        // ---- token 0: which earlier name to diff against
        const uint32_t c0 = live ? cur[0] : 0;
        int t0 = -1;
        if (live && c0 < d0.len) t0 = d_window(arena, d0, c0, w0, w0base);
        __syncwarp();
        if (live && sub == 0 && t0 >= 0) cur[0] = c0 + 1;
        if (live && cnum > B.nreads) { result = -1; live = false; }                     // cnum >= max_names (:1028)
        if (live && (t0 < 0 || (uint32_t)t0 >= ndesc)) { result = 1; live = false; }
        __syncwarp();
        D8 dd = D8{0, 0};
        uint32_t cd = 0, dist = 0;
        if (live) {
            dd = dsc[t0]; cd = cur[t0];
            if ((uint64_t)cd + 4 > dd.len) { result = -1; live = false; }
        }
        if (live) dist = d_u32(arena, dd, cd);
        __syncwarp();
        if (live && sub == 0) cur[t0] = cd + 4;
        __syncwarp();
        if (live && dist > cnum) { result = -1; live = false; }
        const uint32_t pnum = cnum - dist;
        uint4 P = make_uint4(0, 0, 0, 0);
        if (live) P = pnum == last_cnum ? last_nm : NM[pnum];                            // only meaningful when pnum < cnum
        uint8_t *name = O + at;
        uint8_t *mybuf = nbuf + cb * NAME_BUF;
        const uint8_t *pvbuf = nbuf + (cb ^ 1) * NAME_BUF;

        // ---- duplicate of an earlier name (:1038-1058): strcpy up to its first NUL
        bool dup = live && t0 == T_DUP;
        if (dup && pnum == cnum) { result = -1; live = false; dup = false; }
        if (__any_sync(FULL, dup)) {
            const bool in_buf = pnum == buf_cnum;
            const uint8_t *src = O + P.x;
            uint32_t l = 0;
            bool going = dup, over = false;
            while (__any_sync(FULL, going)) {
                const uint32_t i = l + sub;
                uint8_t ch = 1;
                if (going) ch = (in_buf && i < buf_fill) ? pvbuf[i] : src[i];
                const uint32_t z = (__ballot_sync(FULL, going && ch == 0) >> hshift) & 0xffffu;
                const uint32_t n = z ? (uint32_t)__ffs(z) - 1 : 16u;
                if (going) {
                    if ((int64_t)(l + n) + 1 >= room) { over = true; going = false; }
                    else {
                        if (sub < n) { name[i] = ch; if (i < NAME_BUF) mybuf[i] = ch; }
                        l += n;
                        if (z) going = false;
                    }
                }
            }
            if (dup) {
                if (over) { result = -1; live = false; }
                else {
                    if (sub == 0) { name[l] = 0; if (l < NAME_BUF) mybuf[l] = 0; }
                    last_nm = make_uint4((uint32_t)at, P.y, P.z, 0);
                    if (sub == 0) NM[cnum] = last_nm;
                    last_cnum = cnum;
                    buf_cnum = cnum; buf_fill = l + 1 < (uint32_t)NAME_BUF ? l + 1 : (uint32_t)NAME_BUF; cb ^= 1;
                    at += l + 1; room -= l + 1;
                    cnum++;
                    live = false;                                                        // this half is done with this name
                }
            }
        }

        // ---- a name built from tokens: one lane per token position
        const bool tokn = live;                                                          // half-uniform
        if (__any_sync(FULL, tokn)) {
            const uint32_t pntok = pnum == cnum ? 0 : P.y;                               // last_ntok is 0 while a name is in flight (:1071)
            const uint2 *HP = H + (uint64_t)P.z * B.max_tok;
            uint2 *HC = H + (uint64_t)cnum * B.max_tok;
            const bool prev_in_regs = P.z == reg_row;
            const bool prev_in_buf = pnum == buf_cnum;
            const bool active = tokn && k < kmax;
            int tok = -1;
            uint32_t ct = 0;
            if (active) {
                ct = C[T_TYPE];
                if (ct < dt1.len) tok = d_window(arena, dt1, ct, w1, w1base);
            }
            const bool payload = tok == T_ALPHA || tok == T_CHAR || tok == T_DIGITS0 || tok == T_DIGITS ||
                                 tok == T_DDELTA || tok == T_DDELTA0 || tok == T_MATCH || tok == T_NOP;
            const uint32_t endmask = (__ballot_sync(FULL, active && !payload) >> hshift) & 0xffffu;
            const uint32_t e = endmask ? (uint32_t)__ffs(endmask) - 1 : 16u;             // first END / dry type stream of this half
            const bool mine = active && sub <= e;

            // operands: all stream reads up front
            const bool body = mine && sub != e;
            const bool hasq = body && k < pntok;
            const uint32_t cls = body ? (uint32_t)tok : 0u;
            const uint32_t bs = (uint32_t)(0x0000009800004200ull >> (4 * cls)) & 15u;
            const uint32_t ws = (uint32_t)(0x0000000070003000ull >> (4 * cls)) & 15u;
            if (mine && tok >= 0) C[T_TYPE] = ct + 1;
            uint2 q = make_uint2(0, 0);
            if (hasq) q = prev_in_regs ? make_uint2(pv_val, pv_ta) : HP[k];
            bool lerr = false;
            uint32_t bval = 0, wval = 0;
            if (bs) {
                const D8 db = S[bs];
                const uint32_t c = C[bs];
                if (c >= db.len) lerr = true;
                else { bval = (uint32_t)d_byte(arena, db, c); C[bs] = c + 1; }
            }
            if (ws) {
                const D8 dw = S[ws];
                const uint32_t c = C[ws];
                if ((uint64_t)c + 4 > dw.len) lerr = true;
                else { wval = d_u32(arena, dw, c); C[ws] = c + 4; }
            }
            const uint32_t qtype = q.y >> 28, qaux = q.y & 0x0fffffffu;

            // classify (MATCH adopts the earlier token's type, :1133-1183)
            uint32_t flen = 0, need = 0, rtype = T_NOP, rval = 0, raux = 0;
            uint32_t chr = 0, num = 0, numw = 0, numkind = 0;
            bool is_chr = false, is_copy = false, alpha_open = false, synth_alpha = false;
            const uint8_t *csrc = nullptr;
            D8 ds = D8{0, 0};
            uint32_t spos = 0;
            if (mine && sub == e) {                                                      // N_END (:1186-1204)
                is_chr = true; flen = 1; need = 1; rtype = T_END;
            } else if (body && !lerr) {
                const uint32_t t = (uint32_t)tok;
                if ((t == T_MATCH || t == T_DDELTA || t == T_DDELTA0) && !hasq) lerr = true;
                else {
                    const uint32_t et = t == T_MATCH ? (0x100u | qtype) : t;
                    switch (et) {
                    case T_CHAR: case 0x100 | T_CHAR:
                        chr = et == T_CHAR ? bval : (q.x & 0xff);
                        rval = et == T_CHAR ? (uint32_t)(int32_t)(int8_t)bval : q.x;
                        is_chr = true; flen = 1; need = 1; rtype = T_CHAR;
                        break;
                    case T_DIGITS: case T_DDELTA: case 0x100 | T_DIGITS:
                        num = et == T_DIGITS ? wval : et == T_DDELTA ? bval + q.x : q.x;
                        numkind = 1; need = 20; rtype = T_DIGITS; rval = num;
                        break;
                    case T_DIGITS0: case T_DDELTA0: case 0x100 | T_DIGITS0:
                        num = et == T_DIGITS0 ? wval : et == T_DDELTA0 ? bval + q.x : q.x;
                        numw = et == T_DIGITS0 ? bval : qaux;
                        need = et == T_DIGITS0 ? 20 + numw : et == T_DDELTA0 ? numw + 1 : (numw ? numw : 0x80000000u);
                        numkind = 2; flen = numw; rtype = T_DIGITS0; rval = num; raux = numw;
                        break;
                    case T_ALPHA: {
                        ds = S[T_ALPHA];
                        const uint32_t c = C[T_ALPHA];
                        if (c >= ds.len) { lerr = true; break; }
                        uint32_t n = 0;
                        int ch;
                        do { ch = d_byte(arena, ds, c + n); n++; } while (ch && c + n < ds.len);
                        C[T_ALPHA] = c + n;
                        flen = n - 1;
                        need = n; alpha_open = true;
                        spos = c;
                        if (ds.offu >= SYNTH) synth_alpha = true; else { is_copy = true; csrc = arena + (uint64_t)ds.offu * 16 + c; }
                        rtype = T_ALPHA; rval = flen;
                        break; }
                    case 0x100 | T_ALPHA:
                        if ((int32_t)q.x < 0) { lerr = true; break; }
                        flen = q.x; need = q.x ? q.x : 0x80000000u;
                        is_copy = true;
                        csrc = (prev_in_buf && qaux + q.x <= buf_fill) ? pvbuf + qaux : O + P.x + qaux;
                        rtype = T_ALPHA; rval = q.x;
                        break;
                    case T_NOP:
                        rtype = T_NOP;
                        break;
                    default:
                        lerr = true;
                    }
                }
            }

            // decimal text for every numeric lane of both halves at once
            uint64_t slo = chr;
            uint32_t shi = 0, rlen = is_chr ? 1u : 0u;
            if (__any_sync(FULL, numkind != 0)) {
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
                const uint32_t d1 = a / 10u;    const uint32_t d0_ = a - d1 * 10u;
                const uint32_t x0 = (d9 | d8 << 8 | d7 << 16 | d6 << 24) + 0x30303030u;
                const uint32_t x1 = (d5 | d4 << 8 | d3 << 16 | d2 << 24) + 0x30303030u;
                const uint64_t lo = (uint64_t)x0 | (uint64_t)x1 << 32;
                const uint32_t hi = (d1 | d0_ << 8) + 0x3030u;
                const uint32_t sig = num >= 100000u ? (num >= 10000000u ? (num >= 1000000000u ? 10u : num >= 100000000u ? 9u : 8u)
                                                                       : (num >= 1000000u ? 7u : 6u))
                                                    : (num >= 100u ? (num >= 10000u ? 5u : num >= 1000u ? 4u : 3u)
                                                                   : (num >= 10u ? 2u : num >= 1u ? 1u : 0u));
                if (numkind == 1) flen = sig;
                const uint32_t wr = numkind == 1 ? sig : numkind == 2 ? ((numw >= 1 && numw <= 9) ? numw : 0u) : 0u;
                if (numkind) {
                    const uint32_t sh = 10u - wr;
                    if (sh >= 8u) { slo = (uint64_t)(hi >> (8u * (sh - 8u))); shi = 0; }
                    else if (sh == 0u) { slo = lo; shi = hi; }
                    else { slo = (lo >> (8u * sh)) | ((uint64_t)hi << (64u - 8u * sh)); shi = hi >> (8u * sh); }
                    if (numkind == 2 && wr && num >= c_p10[wr]) slo = (slo & ~0xffull) | (uint8_t)(num / c_p10[wr - 1] + '0');
                    rlen = wr;
                }
            }

            // exclusive prefix of the fragment lengths within the half
            uint32_t incl = flen;
            #pragma unroll
            for (int s = 1; s < 16; s <<= 1) {
                const uint32_t t = __shfl_up_sync(FULL, incl, s, 16);
                if (sub >= (uint32_t)s) incl += t;
            }
            const uint32_t off = incl - flen;
            const uint32_t len = __shfl_sync(FULL, incl, 15, 16);                        // the name's length, NUL included
            if (mine && !lerr) {
                if (alpha_open) { if ((int64_t)need > room - (int64_t)off) lerr = true; }
                else if (need == 0x80000000u) { if ((int64_t)off >= room) lerr = true; }
                else if (need && (int64_t)off + (int64_t)need >= room) lerr = true;
                if (rtype == T_ALPHA) raux = off;
                if (raux >> 28) lerr = true;
            }
            // this half fails on any lane error, or when no END came within its token positions (:1207-1209)
            const bool herr = ((__ballot_sync(FULL, lerr) >> hshift) & 0xffffu) != 0 || (tokn && e == 16u);
            const bool good = tokn && !herr;
            const bool wr_ok = mine && good;

            const bool fits = (uint64_t)off + flen <= (uint64_t)NAME_BUF;
            uint8_t *dst = fits ? mybuf + off : name + off;
            if (!wr_ok) rlen = 0;
            const uint32_t maxr = __reduce_max_sync(FULL, rlen);
            for (uint32_t i = 0; i < maxr; i++) {
                if (i < rlen) dst[i] = (uint8_t)slo;
                slo = (slo >> 8) | ((uint64_t)shi << 56); shi >>= 8;
            }
            uint32_t cmask = __ballot_sync(FULL, wr_ok && is_copy && flen > 0);
            while (cmask) {                                                              // all 32 lanes move each copied string
                const int j = __ffs(cmask) - 1;
                cmask &= cmask - 1;
                const uint8_t *s8 = reinterpret_cast<const uint8_t *>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(csrc), j));
                uint8_t *d8 = reinterpret_cast<uint8_t *>(__shfl_sync(FULL, reinterpret_cast<unsigned long long>(dst), j));
                const uint32_t L = __shfl_sync(FULL, flen, j);
                for (uint32_t i = lane; i < L; i += 32) d8[i] = s8[i];
            }
            uint32_t fill = 0;
            if (wr_ok) {
                if (synth_alpha) for (uint32_t i = 0; i < flen; i++) dst[i] = (uint8_t)d_byte(arena, ds, spos + i);
                const uint32_t rec = rtype << 28 | raux;
                HC[k] = make_uint2(rval, rec);
                pv_val = rval; pv_ta = rec;
                if (fits && flen) fill = off + flen;
            }
            #pragma unroll
            for (int s = 8; s >= 1; s >>= 1) { const uint32_t t = __shfl_xor_sync(FULL, fill, s, 16); fill = t > fill ? t : fill; }
            __syncwarp();
            if (good) for (uint32_t i = sub; i < fill; i += 16) name[i] = mybuf[i];      // flush the staged prefix
            if (tokn) {
                if (herr) result = -1;
                else {
                    last_nm = make_uint4((uint32_t)at, e + 1, cnum, 0);                  // ntok = position of the END token
                    if (sub == 0) NM[cnum] = last_nm;
                    last_cnum = cnum; reg_row = cnum;
                    buf_cnum = cnum; buf_fill = fill; cb ^= 1;
                    at += len; room -= len;
                    cnum++;
                }
            }
        }
        __syncwarp();                                                                    // history and name bytes visible to the next name
    }

    if (has && sub == 0) {
        status[b] = result == 1 ? HGPU_OK : result == 2 ? B.host_status : HGPU_TOK3_ERR;
        out_len[b] = result == 1 ? (uint32_t)at : 0;
    }
}

}  // namespace

int hgpu_launch_tok3_names_h16(hgpu_ctx *ctx, const Tok3Block *d_blocks, const uint32_t *d_order, uint32_t n,
                               const Tok3Desc *d_descs, const uint8_t *d_arena, const int32_t *d_job_status,
                               const uint32_t *d_job_got, const uint32_t *d_job_want, uint2 *d_hist, uint4 *d_names,
                               uint8_t *d_out, uint32_t *d_out_len, int32_t *d_status, cudaStream_t st)
{
    (void)ctx;
    if (n == 0) return HGPU_OK;
    const uint32_t per_cta = WARPS * 2;
    tok3_names_h16_kernel<<<(n + per_cta - 1) / per_cta, 32 * WARPS, 0, st>>>(d_blocks, d_order, n, d_descs, d_arena, d_job_status,
                                                                             d_job_got, d_job_want, d_hist, d_names, d_out, d_out_len, d_status);
    hgpu_count_launch();
    return hgpu_check(cudaGetLastError(), "tok3_names_h16_kernel");
}
