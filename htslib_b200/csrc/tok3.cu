// CRAM 3.1 read-name tokeniser ("tok3", block method 8) — decode side.
//
// Replaces tok3_decode_names (htscodecs/htscodecs/tokenise_name3.c:1679-1834) as called from
// cram_uncompress_block (cram/cram_io.c:1753-1765) for a BATCH of name blocks:
//   host   : walks each block's descriptor framing (ttype byte, dup links, varint sizes) and turns the
//            compressed token streams of all blocks into one job list;
//   device : the existing rANS-Nx16 / adaptive-arithmetic batch decoders expand every token stream
//            into an arena, then tok3_names_kernel rebuilds the names, one WARP per block with one
//            LANE per token position (token positions own disjoint streams, so the lanes of a name
//            are independent; names themselves are a serial chain through the name they diff against).
//
// Layout per block in HBM: a descriptor table of max_tok*16 {offset,len,synth} entries into the stream
// arena; a history table (nreads+1) x max_tok of {value, type|aux} so that any earlier name can be the
// reference of a later one (decode_name :1023-1210 keeps the same per-name token history); a name
// table {offset, ntok, history row}.  A duplicate name aliases its source's history row.
#include "hgpu_internal.h"
#include <vector>
#include <string.h>
#include <stdlib.h>
#include <mutex>

namespace {

enum { T_TYPE = 0, T_ALPHA, T_CHAR, T_DIGITS0, T_DZLEN, T_DUP, T_DIFF, T_DIGITS, T_DDELTA,
       T_DDELTA0, T_MATCH, T_NOP, T_END };
constexpr int TOK_MAX = 128;                       // MAX_TOKENS, tokenise_name3.c:115

struct Tok3Desc {                                  // one token stream ("descriptor", :143-148)
    uint64_t off;                                  // byte offset in the stream arena
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

__device__ __forceinline__ int desc_byte(const uint8_t *arena, const Tok3Desc &d, uint32_t pos)
{
    if (d.synth) return pos == 0 ? (int)(d.synth & 0xff) : T_MATCH;
    return arena[d.off + pos];
}
__device__ __forceinline__ Tok3Desc load_desc(const Tok3Desc *p)
{
    uint4 v = __ldg(reinterpret_cast<const uint4 *>(p));
    Tok3Desc d;
    d.off = (uint64_t)v.x | ((uint64_t)v.y << 32);
    d.len = v.z;
    d.synth = v.w;
    return d;
}

__constant__ uint32_t c_p10[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};

// append_uint32_var (:249-286): no leading zeros, nothing at all for 0
__device__ __forceinline__ int var_digits(uint32_t v)
{
    int n = 0;
    #pragma unroll
    for (int k = 0; k < 10; k++) n += v >= c_p10[k] ? 1 : 0;
    return n;                                      // v == 0 -> 0
}
__device__ __forceinline__ void put_var(uint8_t *o, uint32_t v, int n)
{
    for (int k = n - 1; k >= 0; k--) { o[k] = '0' + v % 10; v /= 10; }
}
// append_uint32_fixed (:233-247): w digits; the leading one is stored unreduced; w == 0 or w > 9 writes nothing
__device__ __forceinline__ void put_fixed(uint8_t *o, uint32_t v, uint32_t w)
{
    if (w == 0 || w > 9) return;
    uint32_t p = c_p10[w - 1];
    o[0] = (uint8_t)(v / p + '0');
    v %= p;
    for (uint32_t k = w - 1; k >= 1; k--) { o[k] = '0' + v % 10; v /= 10; }
}

enum { W_NONE = 0, W_BYTE, W_STREAM, W_NAME, W_VAR, W_FIXED, W_NUL };

// One warp per name block.  cur[] (shared) holds the read cursor of each of the block's max_tok*16 streams.
__global__ void __launch_bounds__(32) tok3_names_kernel(const Tok3Block *blocks, const Tok3Desc *descs,
        const uint8_t *arena, const int32_t *job_status, const uint32_t *job_got, const uint32_t *job_want,
        uint2 *hist_all, uint4 *names_all, uint8_t *out, uint32_t *out_len, int32_t *status)
{
    __shared__ uint32_t cur[TOK_MAX * 16];
    const uint32_t lane = threadIdx.x;
    const Tok3Block B = blocks[blockIdx.x];
    const uint32_t b = blockIdx.x;
    if (B.host_status) { if (lane == 0) { status[b] = B.host_status; out_len[b] = 0; } return; }

    // every token stream must have decoded to exactly the size its header announced (:1789-1793)
    bool bad = false;
    for (uint32_t j = lane; j < B.njobs; j += 32)
        bad |= job_status[B.job0 + j] != HGPU_OK || job_got[B.job0 + j] != job_want[B.job0 + j];
    if (__any_sync(0xffffffffu, bad)) { if (lane == 0) { status[b] = HGPU_TOK3_ERR; out_len[b] = 0; } return; }

    const uint32_t ndesc = B.max_tok * 16;
    for (uint32_t i = lane; i < ndesc; i += 32) cur[i] = 0;
    __syncwarp();

    const Tok3Desc *D = descs + B.desc_base;
    uint2 *H = hist_all + B.hist_off;                              // [name][max_tok] {val, type<<28 | aux}
    uint4 *NM = names_all + B.name_off;                            // {offset, ntok, history row, 0}
    uint8_t *O = out + B.out_off;
    const uint32_t kmax = B.max_tok < (uint32_t)TOK_MAX ? B.max_tok : (uint32_t)TOK_MAX;

    int64_t room = (int64_t)B.ulen + 1024;                         // name_len of decode_name (:1810-1818)
    if ((int64_t)B.out_cap < room) room = -1;                      // caller's slot is too small: rejected below
    uint64_t at = 0;
    uint32_t cnum = 0;
    int result = 0;                                                // 0 running, 1 finished, -1 error
    if (room < 0) result = -1;

    while (result == 0) {
        // ---- token 0: which earlier name to diff against (uniform across the warp)
        Tok3Desc d0 = load_desc(&D[0]);
        uint32_t c0 = cur[0];
        int t0 = c0 < d0.len ? desc_byte(arena, d0, c0) : -1;
        __syncwarp();
        if (lane == 0 && t0 >= 0) cur[0] = c0 + 1;
        __syncwarp();
        if (cnum > B.nreads) { result = -1; break; }               // cnum >= max_names (:1028)
        if (t0 < 0 || (uint32_t)t0 >= ndesc) { result = 1; break; }
        Tok3Desc dd = load_desc(&D[t0]);
        uint32_t cd = cur[t0];
        if ((uint64_t)cd + 4 > dd.len) { result = -1; break; }
        uint32_t dist = (uint32_t)desc_byte(arena, dd, cd) | (uint32_t)desc_byte(arena, dd, cd + 1) << 8 |
                        (uint32_t)desc_byte(arena, dd, cd + 2) << 16 | (uint32_t)desc_byte(arena, dd, cd + 3) << 24;
        __syncwarp();
        if (lane == 0) cur[t0] = cd + 4;
        __syncwarp();
        if (dist > cnum) { result = -1; break; }
        const uint32_t pnum = cnum - dist;
        const uint4 P = NM[pnum];                                  // only meaningful when pnum < cnum
        uint8_t *name = O + at;

        if (t0 == T_DUP) {
            if (pnum == cnum) { result = -1; break; }
            // strcpy semantics: up to the first NUL of the earlier name (:1043-1045)
            const uint8_t *src = O + P.x;
            uint32_t l = 0;
            bool stop = false, over = false;
            while (!stop) {
                uint32_t i = l + lane;
                // the earlier name always ends in a NUL this kernel wrote, so the scan terminates
                uint8_t ch = src[i];
                uint32_t z = __ballot_sync(0xffffffffu, ch == 0);
                uint32_t n = z ? (uint32_t)__ffs(z) - 1 : 32u;
                if ((int64_t)(l + n) + 1 >= room) { over = true; break; }
                if (lane < n) name[i] = ch;
                l += n;
                stop = z != 0;
            }
            if (over) { result = -1; break; }
            if (lane == 0) { name[l] = 0; NM[cnum] = make_uint4((uint32_t)at, P.y, P.z, 0); }
            at += l + 1; room -= l + 1;
            cnum++;
            __syncwarp();
            continue;
        }

        const uint32_t pntok = pnum == cnum ? 0 : P.y;             // last_ntok is 0 while a name is in flight (:1071)
        const uint2 *HP = H + (uint64_t)P.z * B.max_tok;
        uint2 *HC = H + (uint64_t)cnum * B.max_tok;
        uint32_t len = 0, ntok = 0;
        bool ended = false, err = false;

        for (uint32_t base = 1; base < kmax && !ended && !err; base += 32) {
            const uint32_t k = base + lane;
            const bool active = k < kmax;
            const Tok3Desc *S = D + (k << 4);
            uint32_t *C = cur + (k << 4);
            int tok = -1;
            Tok3Desc dt;
            uint32_t ct = 0;
            if (active) {
                dt = load_desc(&S[T_TYPE]);
                ct = C[T_TYPE];
                if (ct < dt.len) tok = desc_byte(arena, dt, ct);
            }
            const bool payload = tok == T_ALPHA || tok == T_CHAR || tok == T_DIGITS0 || tok == T_DIGITS ||
                                 tok == T_DDELTA || tok == T_DDELTA0 || tok == T_MATCH || tok == T_NOP;
            const uint32_t endmask = __ballot_sync(0xffffffffu, active && !payload);
            const uint32_t e = endmask ? (uint32_t)__ffs(endmask) - 1 : 32u;   // first END / dry type stream
            const bool mine = active && lane <= e;

            // what this lane contributes
            uint32_t flen = 0, need = 0, wmode = W_NONE, v = 0, w = 0;
            uint64_t srcoff = 0;
            Tok3Desc ds;
            uint32_t rtype = T_NOP, rval = 0, raux = 0;
            bool lerr = false, alpha_open = false;
            if (mine) {
                if (tok >= 0) C[T_TYPE] = ct + 1;                                 // decode_token_type consumed it
                if (lane == e) {                                                  // N_END (:1186-1204)
                    flen = 1; need = 1; wmode = W_NUL; rtype = T_END;
                } else {
                    const bool hasq = k < pntok;
                    uint2 q = hasq ? HP[k] : make_uint2(0, 0);
                    const uint32_t qtype = q.y >> 28, qaux = q.y & 0x0fffffffu;
                    switch (tok) {
                    case T_CHAR: {
                        ds = load_desc(&S[T_CHAR]);
                        uint32_t c = C[T_CHAR];
                        if (c >= ds.len) { lerr = true; break; }
                        v = (uint32_t)desc_byte(arena, ds, c); C[T_CHAR] = c + 1;
                        flen = 1; need = 1; wmode = W_BYTE;
                        rtype = T_CHAR; rval = (uint32_t)(int32_t)(int8_t)v;      // token_int = (char) (:1078)
                        break; }
                    case T_ALPHA: {
                        ds = load_desc(&S[T_ALPHA]);
                        uint32_t c = C[T_ALPHA];
                        if (c >= ds.len) { lerr = true; break; }
                        uint32_t n = 0;                                           // bytes consumed incl. the NUL
                        int ch;
                        do { ch = desc_byte(arena, ds, c + n); n++; } while (ch && c + n < ds.len);
                        C[T_ALPHA] = c + n;
                        flen = n - 1;                                             // a missing NUL drops the last char (:432-437)
                        need = n; alpha_open = true;                              // needs n <= room - len
                        wmode = W_STREAM; srcoff = c;
                        rtype = T_ALPHA; rval = flen;
                        break; }
                    case T_DIGITS0: {
                        ds = load_desc(&S[T_DZLEN]);
                        uint32_t c = C[T_DZLEN];
                        if (c >= ds.len) { lerr = true; break; }
                        w = (uint32_t)desc_byte(arena, ds, c); C[T_DZLEN] = c + 1;
                        ds = load_desc(&S[T_DIGITS0]);
                        c = C[T_DIGITS0];
                        if ((uint64_t)c + 4 > ds.len) { lerr = true; break; }
                        v = (uint32_t)desc_byte(arena, ds, c) | (uint32_t)desc_byte(arena, ds, c + 1) << 8 |
                            (uint32_t)desc_byte(arena, ds, c + 2) << 16 | (uint32_t)desc_byte(arena, ds, c + 3) << 24;
                        C[T_DIGITS0] = c + 4;
                        flen = w; need = 20 + w; wmode = W_FIXED;
                        rtype = T_DIGITS0; rval = v; raux = w;
                        break; }
                    case T_DDELTA0: {
                        if (!hasq) { lerr = true; break; }
                        ds = load_desc(&S[T_DDELTA0]);
                        uint32_t c = C[T_DDELTA0];
                        if (c >= ds.len) { lerr = true; break; }
                        v = (uint32_t)desc_byte(arena, ds, c) + q.x; C[T_DDELTA0] = c + 1;
                        w = qaux;
                        flen = w; need = w + 1; wmode = W_FIXED;
                        rtype = T_DIGITS0; rval = v; raux = w;
                        break; }
                    case T_DIGITS: {
                        ds = load_desc(&S[T_DIGITS]);
                        uint32_t c = C[T_DIGITS];
                        if ((uint64_t)c + 4 > ds.len) { lerr = true; break; }
                        v = (uint32_t)desc_byte(arena, ds, c) | (uint32_t)desc_byte(arena, ds, c + 1) << 8 |
                            (uint32_t)desc_byte(arena, ds, c + 2) << 16 | (uint32_t)desc_byte(arena, ds, c + 3) << 24;
                        C[T_DIGITS] = c + 4;
                        flen = (uint32_t)var_digits(v); need = 20; wmode = W_VAR;
                        rtype = T_DIGITS; rval = v;
                        break; }
                    case T_DDELTA: {
                        if (!hasq) { lerr = true; break; }
                        ds = load_desc(&S[T_DDELTA]);
                        uint32_t c = C[T_DDELTA];
                        if (c >= ds.len) { lerr = true; break; }
                        v = (uint32_t)desc_byte(arena, ds, c) + q.x; C[T_DDELTA] = c + 1;
                        flen = (uint32_t)var_digits(v); need = 20; wmode = W_VAR;
                        rtype = T_DIGITS; rval = v;
                        break; }
                    case T_NOP:
                        rtype = T_NOP;
                        break;
                    default:                                                       // T_MATCH (:1133-1183)
                        if (!hasq) { lerr = true; break; }
                        switch (qtype) {
                        case T_CHAR:
                            v = q.x & 0xff; flen = 1; need = 1; wmode = W_BYTE;
                            rtype = T_CHAR; rval = q.x;
                            break;
                        case T_ALPHA:
                            if ((int32_t)q.x < 0) { lerr = true; break; }
                            flen = q.x; need = q.x; wmode = W_NAME; srcoff = (uint64_t)P.x + qaux;
                            if (q.x == 0) need = 0x80000000u;                      // "len + 0 >= room" still applies: marker
                            rtype = T_ALPHA; rval = q.x;
                            break;
                        case T_DIGITS:
                            v = q.x; flen = (uint32_t)var_digits(v); need = 20; wmode = W_VAR;
                            rtype = T_DIGITS; rval = v;
                            break;
                        case T_DIGITS0:
                            v = q.x; w = qaux; flen = w; need = w; wmode = W_FIXED;
                            if (w == 0) need = 0x80000000u;
                            rtype = T_DIGITS0; rval = v; raux = w;
                            break;
                        default:
                            lerr = true;
                        }
                    }
                }
            }
            // exclusive prefix of the fragment lengths -> where each lane writes
            uint32_t incl = flen;
            #pragma unroll
            for (int s = 1; s < 32; s <<= 1) {
                uint32_t t = __shfl_up_sync(0xffffffffu, incl, s);
                if (lane >= (uint32_t)s) incl += t;
            }
            const uint32_t off = len + incl - flen;
            // the reference's "len + need >= name_len" guards, evaluated with this lane's own len
            if (mine && !lerr) {
                if (alpha_open) { if ((int64_t)need > room - (int64_t)off) lerr = true; }
                else if (need == 0x80000000u) { if ((int64_t)off >= room) lerr = true; }
                else if (need && (int64_t)off + (int64_t)need >= room) lerr = true;
            }
            if (__any_sync(0xffffffffu, lerr)) { err = true; break; }
            if (mine) {
                uint8_t *o = name + off;
                switch (wmode) {
                case W_BYTE: o[0] = (uint8_t)v; break;
                case W_NUL: o[0] = 0; break;
                case W_STREAM: for (uint32_t i = 0; i < flen; i++) o[i] = (uint8_t)desc_byte(arena, ds, (uint32_t)srcoff + i); break;
                case W_NAME: { const uint8_t *s = O + srcoff; for (uint32_t i = 0; i < flen; i++) o[i] = s[i]; break; }
                case W_VAR: put_var(o, v, (int)flen); break;
                case W_FIXED: put_fixed(o, v, w); break;
                default: break;
                }
                if (rtype == T_ALPHA) raux = off;                                  // token_str = offset in the name
                if (raux >> 28) lerr = true;                                       // beyond the packed field (names of 256 MB)
                HC[k] = make_uint2(rval, rtype << 28 | raux);
            }
            if (__any_sync(0xffffffffu, lerr)) { err = true; break; }
            len += __shfl_sync(0xffffffffu, incl, 31);
            if (e < 32) { ended = true; ntok = base + e; }
        }
        if (err || !ended) { result = -1; break; }
        if (lane == 0) NM[cnum] = make_uint4((uint32_t)at, ntok, cnum, 0);
        at += len; room -= len;
        cnum++;
        __syncwarp();                                                              // history and name bytes visible to the next name
    }

    if (lane == 0) {
        status[b] = result == 1 ? HGPU_OK : HGPU_TOK3_ERR;
        out_len[b] = result == 1 ? (uint32_t)at : 0;
    }
}

// big-endian 7-bit varint, var_get_u32 (varint.h:267-299)
int h_vget(const uint8_t *p, const uint8_t *end, uint32_t *v)
{
    const uint8_t *s = p;
    uint32_t acc = 0;
    uint8_t c;
    if (end - p >= 6) {
        int n = 5;
        do { c = *p++; acc = (acc << 7) | (c & 0x7f); } while ((c & 0x80) && n-- > 0);
    } else {
        if (p >= end) { *v = 0; return 0; }
        if (*p < 128) { *v = *p; return 1; }
        do { c = *p++; acc = (acc << 7) | (c & 0x7f); } while ((c & 0x80) && p < end);
    }
    *v = acc;
    return (int)(p - s);
}

float g_tok3_ms[2];
struct Job { uint64_t in_off; uint32_t in_len; uint64_t out_off; uint32_t out_len; };

}  // namespace

extern "C" int hgpu_arith_decode_batch_dev(hgpu_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off,
        const uint32_t *d_in_len, uint32_t n, uint8_t *d_out, const uint64_t *d_out_off, const uint32_t *d_out_len,
        uint32_t *d_got_len, int32_t *d_status, uint32_t max_out_len, void *stream);

extern "C" uint32_t hgpu_tok3_out_bound(const uint8_t *in, uint32_t len)
{
    if (!in || len < 9) return 0;
    uint32_t ulen = in[0] | in[1] << 8 | in[2] << 16 | (uint32_t)in[3] << 24;
    if ((int32_t)ulen < 0 || ulen >= 0x7fffffffu - 1024) return 0;
    return ulen + 1024;
}

extern "C" int hgpu_tok3_decode_batch_host(hgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off,
        const uint32_t *in_len, uint32_t n, uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
        uint32_t *out_len, int32_t *status)
{
    if (!ctx || (n && (!in || !in_off || !in_len || !out || !out_off || !out_cap || !out_len || !status))) {
        hgpu_set_error("bad argument");
        return HGPU_ERR_ARG;
    }
    if (n == 0) return HGPU_OK;
    if (hgpu_check(cudaSetDevice(ctx->device), "cudaSetDevice")) return HGPU_ERR_CUDA;

    // ---- host: descriptor framing of every block (tok3_decode_names :1679-1806)
    std::vector<Tok3Block> blocks(n);
    std::vector<Tok3Desc> descs;
    std::vector<Job> rjobs, ajobs;                 // rANS-Nx16 / adaptive arithmetic
    std::vector<uint32_t> rjob_block_first(n, 0), ajob_block_first(n, 0);
    uint64_t arena = 0, hist = 0, names = 0, in_end = 0, out_end = 0;
    uint32_t max_stream = 0;
    for (uint32_t b = 0; b < n; b++) {
        Tok3Block &B = blocks[b];
        memset(&B, 0, sizeof(B));
        B.out_off = out_off[b]; B.out_cap = out_cap[b];
        B.host_status = HGPU_TOK3_ERR;
        if (in_off[b] + in_len[b] > in_end) in_end = in_off[b] + in_len[b];
        if (out_off[b] + out_cap[b] > out_end) out_end = out_off[b] + out_cap[b];
        const uint8_t *p = in + in_off[b];
        const uint32_t sz = in_len[b];
        if (sz < 9) continue;
        uint32_t ulen = p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24;
        if ((int32_t)ulen < 0 || ulen >= 0x7fffffffu - 1024) continue;
        int32_t nreads = (int32_t)(p[4] | p[5] << 8 | p[6] << 16 | (uint32_t)p[7] << 24);
        const int use_arith = p[8];
        if (nreads <= 0 || nreads > 10000000) continue;                       // create_context :172-187
        std::vector<Job> &jobs = use_arith ? ajobs : rjobs;
        const size_t job_mark = jobs.size(), desc_mark = descs.size();
        const uint64_t arena_mark = arena;
        descs.resize(desc_mark + 16, Tok3Desc{0, 0, 0});
        std::vector<uint8_t> present(TOK_MAX * 16, 0);                        // desc[i].buf != NULL
        uint32_t o = 9;
        int tnum = -1;
        bool ok = true, limit = false;
        while (ok && o < sz) {
            const uint8_t tt = p[o++];
            const bool dup = tt & 64;
            int j = 0;
            if (dup) {
                if (o + 2 > sz) { ok = false; break; }
                j = (p[o] << 4) + p[o + 1]; o += 2;
            }
            if (tt & 128) {
                if (++tnum >= TOK_MAX) { ok = false; break; }
                descs.resize(desc_mark + (size_t)(tnum + 1) * 16, Tok3Desc{0, 0, 0});
                for (int k = 0; k < 16; k++) { descs[desc_mark + (tnum << 4) + k] = Tok3Desc{0, 0, 0}; present[(tnum << 4) + k] = 0; }
            }
            if ((tt & 15) != 0 && (tt & 128)) {
                descs[desc_mark + (tnum << 4)] = Tok3Desc{0, (uint32_t)nreads, 0x100u | (tt & 15u)};
                present[tnum << 4] = 1;
            }
            if (tnum < 0) { ok = false; break; }
            const int i = (tnum << 4) | (tt & 15);
            if (dup) {
                if (j >= i || !present[j]) { ok = false; break; }
                descs[desc_mark + i] = descs[desc_mark + j];
                present[i] = 1;
                continue;
            }
            const uint8_t *s = p + o, *e = p + sz;
            uint32_t clen, usz;
            const int nb = h_vget(s, e, &clen);
            h_vget(s + nb + 1 <= e ? s + nb + 1 : e, e, &usz);
            if ((int32_t)usz < 0 || usz >= 0x7fffffffu) { ok = false; break; }
            // No encoder writes a token stream longer than 4 bytes per name (integers) or two per name
            // byte (strings + NUL); beyond that the reference would still malloc(usz) and decode, this
            // implementation refuses instead of sizing device arenas from a corrupt field.
            if ((uint64_t)usz > 4ull * (uint64_t)nreads + 2ull * ulen + 1024) { ok = false; limit = true; break; }
            if ((uint64_t)o + nb > sz) { ok = false; break; }                 // nothing left for the sub-decoder
            Job jb;
            jb.in_off = in_off[b] + o + nb;
            jb.in_len = sz - o - nb;                                          // the sub-decoder is handed the rest of the block (:1436)
            jb.out_off = arena;
            jb.out_len = usz;
            jobs.push_back(jb);
            descs[desc_mark + i] = Tok3Desc{arena, usz, 0};
            present[i] = 1;
            arena += ((uint64_t)usz + 15) & ~(uint64_t)15;
            if (usz > max_stream) max_stream = usz;
            if ((uint64_t)o + clen + nb > 0xffffffffull) { ok = false; break; }
            o += clen + nb;
        }
        if (!ok) {                                                            // drop what this block queued
            jobs.resize(job_mark); descs.resize(desc_mark); arena = arena_mark;
            if (limit) B.host_status = HGPU_TOK3_ERR_LIMIT;
            continue;
        }
        B.host_status = HGPU_OK;
        B.desc_base = (uint32_t)desc_mark;
        B.max_tok = (uint32_t)(tnum + 1 > 1 ? tnum + 1 : 1);
        B.nreads = (uint32_t)nreads;
        B.ulen = ulen;
        B.job0 = (uint32_t)job_mark | (use_arith ? 0x80000000u : 0);          // rebased below
        B.njobs = (uint32_t)(jobs.size() - job_mark);
        B.hist_off = hist; hist += (uint64_t)(nreads + 1) * B.max_tok;
        B.name_off = names; names += (uint64_t)nreads + 1;
    }
    const uint32_t nr = (uint32_t)rjobs.size(), na = (uint32_t)ajobs.size(), nj = nr + na;
    for (uint32_t b = 0; b < n; b++) {
        Tok3Block &B = blocks[b];
        if (B.host_status) continue;
        B.job0 = (B.job0 & 0x80000000u) ? (B.job0 & 0x7fffffffu) + nr : B.job0;
    }
    if (descs.empty()) descs.push_back(Tok3Desc{0, 0, 0});

    // ---- device layout in the staging buffer
    auto up = [](uint64_t x) { return (x + 255) & ~(uint64_t)255; };
    const uint64_t o_in = 0, o_arena = o_in + up(in_end + 8), o_out = o_arena + up(arena + 16),
                   o_hist = o_out + up(out_end), o_names = o_hist + up(hist * 8), o_blocks = o_names + up(names * 16),
                   o_descs = o_blocks + up((uint64_t)n * sizeof(Tok3Block)), o_jio = o_descs + up(descs.size() * sizeof(Tok3Desc)),
                   o_joo = o_jio + up((uint64_t)nj * 8), o_jil = o_joo + up((uint64_t)nj * 8), o_jol = o_jil + up((uint64_t)nj * 4),
                   o_jgot = o_jol + up((uint64_t)nj * 4), o_jst = o_jgot + up((uint64_t)nj * 4), o_olen = o_jst + up((uint64_t)nj * 4),
                   o_st = o_olen + up((uint64_t)n * 4), total = o_st + up((uint64_t)n * 4);
    int rc = hgpu_ensure_stage(ctx, total + 256);
    if (rc) return rc;
    uint8_t *base = ctx->d_stage;
    cudaStream_t s = ctx->stream;
    std::vector<uint64_t> jio(nj), joo(nj);
    std::vector<uint32_t> jil(nj), jol(nj);
    for (uint32_t k = 0; k < nj; k++) {
        const Job &jb = k < nr ? rjobs[k] : ajobs[k - nr];
        jio[k] = jb.in_off; joo[k] = jb.out_off; jil[k] = jb.in_len; jol[k] = jb.out_len;
    }
    if (hgpu_check(cudaMemcpyAsync(base + o_in, in, in_end, cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaMemcpyAsync(base + o_blocks, blocks.data(), (size_t)n * sizeof(Tok3Block), cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaMemcpyAsync(base + o_descs, descs.data(), descs.size() * sizeof(Tok3Desc), cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
    if (nj) {
        if (hgpu_check(cudaMemcpyAsync(base + o_jio, jio.data(), (size_t)nj * 8, cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
        if (hgpu_check(cudaMemcpyAsync(base + o_joo, joo.data(), (size_t)nj * 8, cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
        if (hgpu_check(cudaMemcpyAsync(base + o_jil, jil.data(), (size_t)nj * 4, cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
        if (hgpu_check(cudaMemcpyAsync(base + o_jol, jol.data(), (size_t)nj * 4, cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
    }
    const uint64_t *d_jio = (const uint64_t *)(base + o_jio), *d_joo = (const uint64_t *)(base + o_joo);
    const uint32_t *d_jil = (const uint32_t *)(base + o_jil), *d_jol = (const uint32_t *)(base + o_jol);
    uint32_t *d_jgot = (uint32_t *)(base + o_jgot);
    int32_t *d_jst = (int32_t *)(base + o_jst);
    cudaEvent_t tev[3];
    for (int k = 0; k < 3; k++) if (hgpu_check(cudaEventCreate(&tev[k]), "event")) return HGPU_ERR_CUDA;
    cudaEventRecord(tev[0], s);
    if (nr) {
        rc = hgpu_launch_rans_nx16(ctx, base + o_in, d_jio, d_jil, nr, base + o_arena, d_joo, d_jol, d_jgot, d_jst, max_stream, s);
        if (rc) return rc;
    }
    if (na) {
        rc = hgpu_arith_decode_batch_dev(ctx, base + o_in, d_jio + nr, d_jil + nr, na, base + o_arena, d_joo + nr, d_jol + nr,
                                         d_jgot + nr, d_jst + nr, max_stream, s);
        if (rc) return rc;
    }
    cudaEventRecord(tev[1], s);
    tok3_names_kernel<<<n, 32, 0, s>>>((const Tok3Block *)(base + o_blocks), (const Tok3Desc *)(base + o_descs), base + o_arena,
                                      d_jst, d_jgot, d_jol, (uint2 *)(base + o_hist), (uint4 *)(base + o_names),
                                      base + o_out, (uint32_t *)(base + o_olen), (int32_t *)(base + o_st));
    if (hgpu_check(cudaGetLastError(), "tok3_names_kernel")) return HGPU_ERR_CUDA;
    cudaEventRecord(tev[2], s);
    hgpu_count_launch();
    if (hgpu_check(cudaMemcpyAsync(out_len, base + o_olen, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "D2H")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaMemcpyAsync(status, base + o_st, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "D2H")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaMemcpyAsync(out, base + o_out, out_end, cudaMemcpyDeviceToHost, s), "D2H")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaStreamSynchronize(s), "sync")) return HGPU_ERR_CUDA;
    cudaEventElapsedTime(&g_tok3_ms[0], tev[0], tev[1]);
    cudaEventElapsedTime(&g_tok3_ms[1], tev[1], tev[2]);
    for (int k = 0; k < 3; k++) cudaEventDestroy(tev[k]);
    return HGPU_OK;
}

// device time of the last hgpu_tok3_decode_batch_host call: [0] token-stream entropy decode, [1] name rebuild
extern "C" void hgpu_tok3_last_ms(float *ms2) { ms2[0] = g_tok3_ms[0]; ms2[1] = g_tok3_ms[1]; }

// Drop-in for the reference symbol (tokenise_name3.h:59): one block, malloc'd result, NULL on failure.
static std::mutex g_tok3_mu;
static hgpu_ctx *g_tok3_ctx;
extern "C" uint8_t *tok3_decode_names(uint8_t *in, uint32_t sz, uint32_t *out_len)
{
    if (!in || !out_len) return nullptr;
    uint32_t cap = hgpu_tok3_out_bound(in, sz);
    if (!cap) return nullptr;
    std::lock_guard<std::mutex> lock(g_tok3_mu);
    if (!g_tok3_ctx) g_tok3_ctx = hgpu_create(-1);
    if (!g_tok3_ctx) return nullptr;
    uint8_t *out = (uint8_t *)malloc(cap);
    if (!out) return nullptr;
    uint64_t ioff = 0, ooff = 0;
    uint32_t got = 0;
    int32_t st = 0;
    int rc = hgpu_tok3_decode_batch_host(g_tok3_ctx, in, &ioff, &sz, 1, out, &ooff, &cap, &got, &st);
    if (rc != HGPU_OK || st != HGPU_OK) { free(out); return nullptr; }
    *out_len = got;
    return out;
}
