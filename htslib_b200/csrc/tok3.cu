// CRAM 3.1 read-name tokeniser ("tok3", block method 8) — decode side.
//
// Replaces tok3_decode_names (htscodecs/htscodecs/tokenise_name3.c:1679-1834) as called from
// cram_uncompress_block (cram/cram_io.c:1753-1765) for a BATCH of name blocks:
//   host   : walks each block's descriptor framing (ttype byte, dup links, varint sizes) and turns the
//            compressed token streams of all blocks into one job list;
//   device : the existing rANS-Nx16 / adaptive-arithmetic batch decoders expand every token stream
//            into an arena, then tok3_names_kernel (tok3_names.cu) rebuilds the names, one warp per block.
//
// Layout per block in HBM: a descriptor table of max_tok*16 {offset,len,synth} entries into the stream
// arena; a history table (nreads+1) x max_tok of {value, type|aux} so that any earlier name can be the
// reference of a later one (decode_name :1023-1210 keeps the same per-name token history); a name
// table {offset, ntok, history row}.  A duplicate name aliases its source's history row.
#include "tok3_internal.h"
#include <vector>
#include <string.h>
#include <stdlib.h>
#include <mutex>

namespace {

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

static int hgpu_tok3_decode_batch_host_impl(hgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off,
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
    uint32_t max_stream = 0, max_ndesc = 16;
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
        if (B.max_tok * 16 > max_ndesc) max_ndesc = B.max_tok * 16;
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
                   o_st = o_olen + up((uint64_t)n * 4), o_order = o_st + up((uint64_t)n * 4), total = o_order + up((uint64_t)n * 4);
    int rc = hgpu_ensure_stage(ctx, total + 256);
    if (rc) return rc;
    uint8_t *base = ctx->d_stage;
    // blocks with at most 16 token positions (every Illumina-style block) share a warp in pairs; the rest,
    // and the blocks the framing walk already rejected (their status still has to be written), take the
    // general kernel.  order[] = [paired blocks ..., general blocks ...]
    std::vector<uint32_t> order(n);
    uint32_t n16 = 0, max_ndesc_gen = 16;
    const bool h16_ok = arena < (1ull << 36) - 4096;                 // its shared-memory descriptors hold offsets in 16-byte units
    for (uint32_t b = 0; b < n; b++)
        if (h16_ok && blocks[b].host_status == HGPU_OK && blocks[b].max_tok <= TOK3_H16_MAX_TOK) order[n16++] = b;
    {
        uint32_t g = n16;
        for (uint32_t b = 0; b < n; b++)
            if (!(h16_ok && blocks[b].host_status == HGPU_OK && blocks[b].max_tok <= TOK3_H16_MAX_TOK)) {
                order[g++] = b;
                if (blocks[b].host_status == HGPU_OK && blocks[b].max_tok * 16 > max_ndesc_gen) max_ndesc_gen = blocks[b].max_tok * 16;
            }
    }
    (void)max_ndesc;
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
    struct Events {                                                  // released on every return path
        cudaEvent_t e[3] = {nullptr, nullptr, nullptr};
        ~Events() { for (cudaEvent_t x : e) if (x) cudaEventDestroy(x); }
    } evs;
    cudaEvent_t *tev = evs.e;
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
    if (hgpu_check(cudaMemcpyAsync(base + o_order, order.data(), (size_t)n * 4, cudaMemcpyHostToDevice, s), "H2D")) return HGPU_ERR_CUDA;
    const uint32_t *d_order = (const uint32_t *)(base + o_order);
    rc = hgpu_launch_tok3_names_h16(ctx, (const Tok3Block *)(base + o_blocks), d_order, n16, (const Tok3Desc *)(base + o_descs), base + o_arena,
                                    d_jst, d_jgot, d_jol, (uint2 *)(base + o_hist), (uint4 *)(base + o_names),
                                    base + o_out, (uint32_t *)(base + o_olen), (int32_t *)(base + o_st), s);
    if (rc) return rc;
    rc = hgpu_launch_tok3_names(ctx, (const Tok3Block *)(base + o_blocks), d_order + n16, n - n16, max_ndesc_gen, (const Tok3Desc *)(base + o_descs),
                                base + o_arena, d_jst, d_jgot, d_jol, (uint2 *)(base + o_hist), (uint4 *)(base + o_names),
                                base + o_out, (uint32_t *)(base + o_olen), (int32_t *)(base + o_st), s);
    if (rc) return rc;
    cudaEventRecord(tev[2], s);
    if (hgpu_check(cudaMemcpyAsync(out_len, base + o_olen, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "D2H")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaMemcpyAsync(status, base + o_st, (size_t)n * 4, cudaMemcpyDeviceToHost, s), "D2H")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaMemcpyAsync(out, base + o_out, out_end, cudaMemcpyDeviceToHost, s), "D2H")) return HGPU_ERR_CUDA;
    if (hgpu_check(cudaStreamSynchronize(s), "sync")) return HGPU_ERR_CUDA;
    cudaEventElapsedTime(&g_tok3_ms[0], tev[0], tev[1]);
    cudaEventElapsedTime(&g_tok3_ms[1], tev[1], tev[2]);
    return HGPU_OK;
}

// no C++ exception may cross the C ABI (host buffers are sized from untrusted input: std::bad_alloc)
extern "C" int hgpu_tok3_decode_batch_host(hgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off,
        const uint32_t *in_len, uint32_t n, uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap,
        uint32_t *out_len, int32_t *status)
{
    try {
        return hgpu_tok3_decode_batch_host_impl(ctx, in, in_off, in_len, n, out, out_off, out_cap, out_len, status);
    } catch (const std::bad_alloc &) {
        hgpu_set_error("out of host memory");
        return HGPU_ERR_NOMEM;
    } catch (...) {
        hgpu_set_error("internal error");
        return HGPU_ERR_CUDA;
    }
}

// device time of the last hgpu_tok3_decode_batch_host call: [0] token-stream entropy decode, [1] name rebuild
extern "C" void hgpu_tok3_last_ms(float *ms2) { ms2[0] = g_tok3_ms[0]; ms2[1] = g_tok3_ms[1]; }

// Drop-in for the reference symbol (tokenise_name3.h:59): one block, malloc'd result, NULL on failure.
namespace { struct ShimLock { ShimLock() { hgpu_shim_lock(); } ~ShimLock() { hgpu_shim_unlock(); } }; }
extern "C" uint8_t *tok3_decode_names(uint8_t *in, uint32_t sz, uint32_t *out_len)
{
    if (!in || !out_len) return nullptr;
    uint32_t cap = hgpu_tok3_out_bound(in, sz);
    if (!cap) return nullptr;
    ShimLock lock;
    hgpu_ctx *g_tok3_ctx = hgpu_shim_ctx();
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
