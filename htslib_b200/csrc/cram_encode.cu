// hgpu_cram_encode_records_host: bam1_t records -> a complete CRAM 3.x file image — cram_encode_container /
// cram_encode_slice (cram/cram_encode.c:1950-2420), cram_encode_compression_header (:380-1030), cram_write_container and
// the file framing (cram_io.c:3958-4100, :4694, :4889, :5512), in the no-reference shape described in cram_encode.cuh.
//
// Device: cram_enc_count_kernel (one thread per record: bytes per series), cram_enc_scan_kernel (one warp per
// (slice, series): counts -> offsets), cram_enc_write_kernel (one thread per record: the bytes).  The series blocks are
// then compressed by the existing device encoders — the method trial of hgpu_cram_compress_blocks_host for every series
// (rANS Nx16 family for CRAM 3.1, rANS 4x8 for 3.0) and the tok3 encoder for read names (3.1) — and framed with the
// device CRC-32.  Host: the tag dictionary (one walk over the aux field headers), compression / slice / container headers.
// One slice per container; slices are independent, which is the axis that shards across GPUs.
//
// Built a second time by tests/hostsim (-DHGPU_HOSTSIM): kernels -> loops, blocks stored RAW (the codecs are GPU-only);
// the reference must read that file back to the input records.  libhtsgpu.so never contains that variant.
#ifdef HGPU_HOSTSIM
#include "../../include/htsgpu.h"
#include <stdarg.h>
#include <stdio.h>
static char g_enc_err[256];
static void hgpu_set_error(const char *fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_enc_err, sizeof g_enc_err, fmt, ap); va_end(ap); }
extern "C" const char *hostsim_enc_last_error(void) { return g_enc_err; }
struct hgpu_ctx;
#else
#include "hgpu_internal.h"
#endif
#include "cram_encode.cuh"
#include <new>
#include <map>
#include <string>
#include <vector>
#include <stdlib.h>
#include <string.h>

using namespace cramenc;

// (declared at file scope: inside the unnamed namespace they would get hidden visibility and drag the definitions with them)
#ifndef HGPU_HOSTSIM
extern "C" int hgpu_cram_compress_blocks_host(hgpu_ctx *ctx, const uint8_t *const *payload, const uint32_t *payload_len,
        const uint32_t *method_mask, const int32_t *content_id, const uint8_t *content_type, uint32_t n,
        uint8_t *out, uint64_t cap, uint64_t *out_off, uint64_t *out_len, int32_t *chosen);
extern "C" int hgpu_tok3_encode_batch_host(hgpu_ctx *ctx, const uint8_t *in, const uint64_t *in_off, const uint32_t *in_len, uint32_t n,
        uint8_t *out, const uint64_t *out_off, const uint32_t *out_cap, uint32_t *out_len, int32_t *status);
#endif


namespace {

static_assert(sizeof(Core) == 48 && sizeof(hgpu_bam1_core) == 48, "bam1_core_t mirror");

uint32_t host_crc32(const uint8_t *p, size_t n, uint32_t crc = 0)         // container headers (a few dozen bytes each)
{
    static uint32_t tab[256];
    static bool init = false;
    if (!init) { for (uint32_t i = 0; i < 256; i++) { uint32_t c = i; for (int k = 0; k < 8; k++) c = (c & 1) ? (c >> 1) ^ 0xEDB88320u : c >> 1; tab[i] = c; } init = true; }
    crc = ~crc;
    for (size_t i = 0; i < n; i++) crc = tab[(crc ^ p[i]) & 0xff] ^ (crc >> 8);
    return ~crc;
}

struct Buf {
    std::vector<uint8_t> v;
    void u8(uint8_t b) { v.push_back(b); }
    void itf8(int32_t x) { uint8_t t[5]; const int n = itf8_put(t, (uint32_t)x); v.insert(v.end(), t, t + n); }
    void ltf8(int64_t x)                                                   // ltf8_put, cram/cram_io.c
    {
        const uint64_t u = (uint64_t)x;
        int n = 0;
        while (n < 8 && (u >> (7 * (n + 1))) != 0) n++;                    // n extra bytes: value < 2^(7 (n + 1)), 9 bytes for 64 bits
        if (n == 8) { v.push_back(0xff); for (int k = 7; k >= 0; k--) v.push_back((uint8_t)(u >> (8 * k))); return; }
        v.push_back((uint8_t)((0xff00u >> n) & 0xff) | (uint8_t)(u >> (8 * n)));
        for (int k = n - 1; k >= 0; k--) v.push_back((uint8_t)(u >> (8 * k)));
    }
    void le32(uint32_t x) { for (int k = 0; k < 4; k++) v.push_back((uint8_t)(x >> (8 * k))); }
    void bytes(const void *p, size_t n) { v.insert(v.end(), (const uint8_t *)p, (const uint8_t *)p + n); }
};

// one block, cram_write_block's layout (cram_io.c:1511-1563); crc = CRC-32 of everything before it
void frame_block(Buf &o, int method, int ctype, int32_t cid, const uint8_t *payload, uint32_t comp, uint32_t uncomp)
{
    const size_t at = o.v.size();
    o.u8((uint8_t)method); o.u8((uint8_t)ctype); o.itf8(cid); o.itf8((int32_t)comp); o.itf8((int32_t)uncomp);
    o.bytes(payload, comp);
    o.le32(host_crc32(o.v.data() + at, o.v.size() - at));
}

const char *const k_keys[S_COUNT] = {"BF", "CF", "RI", "RL", "AP", "RG", "RN", "MF", "NS", "NP", "TS", "TL", "FN", "FC", "FP", "DL", "RS", "HC", "PD",
                                     nullptr, "BB", nullptr, "SC", nullptr, "IN", "BA", "QS", "MQ", nullptr, nullptr, "BS"};

void enc_external(Buf &o, int id) { o.itf8(1); Buf t; t.itf8(id); o.itf8((int32_t)t.v.size()); o.bytes(t.v.data(), t.v.size()); }
void enc_byte_array_len(Buf &o, int len_id, int val_id)
{
    Buf t;
    enc_external(t, len_id);
    enc_external(t, val_id);
    o.itf8(4); o.itf8((int32_t)t.v.size()); o.bytes(t.v.data(), t.v.size());
}

// cram_encode_compression_header :380-1030 for this writer's fixed layout
void compression_header(Buf &o, const std::vector<std::string> &tag_lines, const std::vector<uint32_t> &tag_keys, bool ref_required)
{
    Buf pm;                                                                // preservation map
    pm.itf8(5);
    pm.u8('R'); pm.u8('N'); pm.u8(1);
    pm.u8('A'); pm.u8('P'); pm.u8(0);
    pm.u8('R'); pm.u8('R'); pm.u8(ref_required ? 1 : 0);
    pm.u8('S'); pm.u8('M'); { const uint8_t sm[5] = {0x1b, 0x1b, 0x1b, 0x1b, 0x1b}; pm.bytes(sm, 5); }   // the default matrix (:165): CGTN AGTN ACTN ACGN ACGT
    pm.u8('T'); pm.u8('D');
    { Buf td; for (const std::string &l : tag_lines) { td.bytes(l.data(), l.size()); td.u8(0); } if (tag_lines.empty()) td.u8(0);
      pm.itf8((int32_t)td.v.size()); pm.bytes(td.v.data(), td.v.size()); }
    o.itf8((int32_t)pm.v.size()); o.bytes(pm.v.data(), pm.v.size());
    Buf rm;                                                                // record encoding map
    int cnt = 0;
    Buf body;
    for (int s = 0; s < S_COUNT; s++) {
        if (!k_keys[s]) continue;
        body.u8((uint8_t)k_keys[s][0]); body.u8((uint8_t)k_keys[s][1]);
        if (s == S_RN) { body.itf8(5); Buf t; t.u8(0); t.itf8(s + 1); body.itf8((int32_t)t.v.size()); body.bytes(t.v.data(), t.v.size()); }
        else if (s == S_BB || s == S_SC || s == S_IN) enc_byte_array_len(body, s, s + 1);              // the length stream sits just before its value stream
        else enc_external(body, s + 1);
        cnt++;
    }
    rm.itf8(cnt); rm.bytes(body.v.data(), body.v.size());
    o.itf8((int32_t)rm.v.size()); o.bytes(rm.v.data(), rm.v.size());
    Buf tm;                                                                // tag encoding map
    tm.itf8((int32_t)tag_keys.size());
    for (uint32_t k : tag_keys) { tm.itf8((int32_t)k); enc_byte_array_len(tm, S_TAG_LEN + 1, S_TAG_VAL + 1); }
    o.itf8((int32_t)tm.v.size()); o.bytes(tm.v.data(), tm.v.size());
}

struct EArgs {
    const Core *core; const uint8_t *data; const uint64_t *data_off; const int32_t *tl;
    const uint8_t *ref_bases; const uint64_t *ref_off; int32_t n_ref;      // reference sequences (ref_bases == nullptr: the no-reference shape)
    uint64_t n; uint32_t rps;               // records, records per slice
    uint32_t *cnt;                          // [slice][stream][rps]: counts, then exclusive offsets
    uint32_t *tot;                          // [slice][stream]
    const uint64_t *base;                   // [slice][stream] -> byte offset in arena
    uint8_t *arena;
    int32_t *status;                        // per record
};

CRAMREC_HD inline void count_body(const EArgs &A, uint64_t g)
{
    const uint32_t sl = (uint32_t)(g / A.rps), r = (uint32_t)(g % A.rps);
    uint32_t n[S_COUNT];
    for (int s = 0; s < S_COUNT; s++) n[s] = 0;
    Emit<false> E{n, nullptr};
    const uint8_t *ref = nullptr; int64_t rl = 0;
    if (A.ref_bases && A.core[g].tid >= 0 && A.core[g].tid < A.n_ref) { ref = A.ref_bases + A.ref_off[A.core[g].tid]; rl = (int64_t)(A.ref_off[A.core[g].tid + 1] - A.ref_off[A.core[g].tid]); }
    const int rc = walk<false>(A.core[g], A.data + A.data_off[g], (uint32_t)(A.data_off[g + 1] - A.data_off[g]), A.tl[g], ref, rl, E);
    A.status[g] = rc;
    for (int s = 0; s < S_COUNT; s++) A.cnt[((size_t)sl * S_COUNT + s) * A.rps + r] = rc == ENC_OK ? n[s] : 0;
}

CRAMREC_HD inline void write_body(const EArgs &A, uint64_t g)
{
    if (A.status[g] != ENC_OK) return;
    const uint32_t sl = (uint32_t)(g / A.rps), r = (uint32_t)(g % A.rps);
    uint32_t n[S_COUNT];
    uint8_t *base[S_COUNT];
    for (int s = 0; s < S_COUNT; s++) { n[s] = A.cnt[((size_t)sl * S_COUNT + s) * A.rps + r]; base[s] = A.arena + A.base[(size_t)sl * S_COUNT + s]; }
    Emit<true> E{n, base};
    const uint8_t *ref = nullptr; int64_t rl = 0;
    if (A.ref_bases && A.core[g].tid >= 0 && A.core[g].tid < A.n_ref) { ref = A.ref_bases + A.ref_off[A.core[g].tid]; rl = (int64_t)(A.ref_off[A.core[g].tid + 1] - A.ref_off[A.core[g].tid]); }
    walk<true>(A.core[g], A.data + A.data_off[g], (uint32_t)(A.data_off[g + 1] - A.data_off[g]), A.tl[g], ref, rl, E);
}

#ifndef HGPU_HOSTSIM
__global__ void __launch_bounds__(128) cram_enc_count_kernel(EArgs A)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < A.n) count_body(A, g);
}
// one warp per (slice, stream): exclusive scan of that row in place, total to tot[]
__global__ void __launch_bounds__(128) cram_enc_scan_kernel(EArgs A, uint32_t rows)
{
    const uint32_t row = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
    if (row >= rows) return;
    const uint32_t sl = row / S_COUNT;
    const uint64_t first = (uint64_t)sl * A.rps;
    const uint32_t nr = (uint32_t)(A.n - first < A.rps ? A.n - first : A.rps);
    uint32_t *p = A.cnt + (size_t)row * A.rps;
    uint32_t run = 0;
    for (uint32_t b = 0; b < nr; b += 32) {
        const uint32_t i = b + lane, v = i < nr ? p[i] : 0;
        uint32_t inc = v;
        for (int d = 1; d < 32; d <<= 1) { const uint32_t t = __shfl_up_sync(0xffffffffu, inc, d); if (lane >= (uint32_t)d) inc += t; }
        if (i < nr) p[i] = run + inc - v;
        run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) A.tot[row] = run;
}
__global__ void __launch_bounds__(128) cram_enc_write_kernel(EArgs A)
{
    const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g < A.n) write_body(A, g);
}
#endif

inline size_t up256(size_t x) { return (x + 255) & ~(size_t)255; }

int encode_impl(hgpu_ctx *ctx, const char *header_text, uint32_t header_len, const hgpu_bam1_core *core, const uint8_t *data,
                const uint64_t *data_off, uint64_t n, const hgpu_cram_refs *refs, uint32_t rps, int minor, uint8_t **out_file, uint64_t *out_len)
{
    // reference-based shape only when every mapped record's reference sequence was supplied (the reader will need them all)
    bool use_ref = refs && refs->bases && refs->off && refs->n_ref > 0;
    if (use_ref)
        for (uint64_t g = 0; g < n && use_ref; g++) {
            const int32_t t = core[g].tid;
            if (!(core[g].flag & 4) && (t < 0 || t >= refs->n_ref || refs->off[t + 1] == refs->off[t])) use_ref = false;
        }
    const uint64_t ref_bytes = use_ref ? refs->off[refs->n_ref] : 0;
    if (!out_file || !out_len || (n && (!core || !data || !data_off)) || (header_len && !header_text)) { hgpu_set_error("cram encode: null argument"); return HGPU_ERR_ARG; }
    *out_file = nullptr; *out_len = 0;
    if (rps == 0) rps = 10000;
    if (minor != 0 && minor != 1) { hgpu_set_error("cram encode: CRAM 3.0 or 3.1"); return HGPU_ERR_ARG; }
    const uint32_t ns = (uint32_t)((n + rps - 1) / rps);
    // ---- tag dictionary per slice (host: one walk over the aux field headers) ----
    std::vector<int32_t> tl(n ? n : 1, 0);
    std::vector<std::vector<std::string>> lines(ns);
    std::vector<std::vector<uint32_t>> keys(ns);
    for (uint32_t sl = 0; sl < ns; sl++) {
        std::map<std::string, int32_t> seen;
        std::map<uint32_t, int> kseen;
        const uint64_t a = (uint64_t)sl * rps, b = a + rps < n ? a + rps : n;
        for (uint64_t g = a; g < b; g++) {
            const hgpu_bam1_core &c = core[g];
            const uint8_t *d = data + data_off[g], *end = data + data_off[g + 1];
            const uint64_t fixed = (uint64_t)c.l_qname + 4ull * c.n_cigar + ((uint64_t)(c.l_qseq < 0 ? 0 : c.l_qseq) + 1) / 2 + (uint64_t)(c.l_qseq < 0 ? 0 : c.l_qseq);
            std::string line;
            if (fixed <= (uint64_t)(end - d)) {
                for (const uint8_t *p = d + fixed; p < end;) {
                    uint32_t vlen = 0;
                    if (!aux_field(p, end, vlen)) break;                   // the count pass flags the record
                    line.append((const char *)p, 3);
                    const uint32_t key = (uint32_t)p[0] << 16 | (uint32_t)p[1] << 8 | p[2];
                    if (!kseen.count(key)) { kseen[key] = 1; keys[sl].push_back(key); }
                    p += 3 + vlen;
                }
            }
            auto it = seen.find(line);
            if (it == seen.end()) { const int32_t k = (int32_t)lines[sl].size(); seen[line] = k; lines[sl].push_back(line); tl[g] = k; }
            else tl[g] = it->second;
        }
    }
    // ---- device: count, scan, write ----
    const size_t rows = (size_t)ns * S_COUNT;
    std::vector<uint32_t> tot(rows ? rows : 1, 0);
    std::vector<uint64_t> base(rows ? rows : 1, 0);
    std::vector<int32_t> status(n ? n : 1, 0);
    std::vector<uint8_t> arena_h;
    const uint64_t data_bytes = n ? data_off[n] : 0;
    if (n) {
        struct Seg { size_t off, bytes; };
        size_t total = 0;
        auto seg = [&](size_t bytes) { Seg s{total, bytes}; total += up256(bytes + 16); return s; };
        const Seg s_ref = seg(ref_bytes), s_roff = seg(use_ref ? ((size_t)refs->n_ref + 1) * 8 : 0);
        const Seg s_core = seg(n * 48), s_data = seg(data_bytes), s_doff = seg((n + 1) * 8), s_tl = seg(n * 4), s_cnt = seg(rows * rps * 4),
                  s_tot = seg(rows * 4), s_base = seg(rows * 8), s_st = seg(n * 4);
        // every series byte comes from the record data, ITF8 at most 5 bytes per value: bound the arena before the scan
        const size_t arena_cap = up256(2 * data_bytes + 200 * n + 4096);
        const Seg s_arena = seg(arena_cap);
#ifdef HGPU_HOSTSIM
        (void)ctx;
        std::vector<uint8_t> image(total);
        uint8_t *b0 = image.data();
        memcpy(b0 + s_core.off, core, n * 48); memcpy(b0 + s_data.off, data, data_bytes); memcpy(b0 + s_doff.off, data_off, (n + 1) * 8);
        memcpy(b0 + s_tl.off, tl.data(), n * 4);
        if (use_ref) { memcpy(b0 + s_ref.off, refs->bases, ref_bytes); memcpy(b0 + s_roff.off, refs->off, ((size_t)refs->n_ref + 1) * 8); }
#else
        if (!ctx) { hgpu_set_error("null context"); return HGPU_ERR_ARG; }
        if (cudaSetDevice(ctx->device) != cudaSuccess) return HGPU_ERR_CUDA;
        int rc0 = hgpu_ensure_stage(ctx, total + 256);
        if (rc0) return rc0;
        uint8_t *b0 = ctx->d_stage;
        cudaStream_t st = ctx->stream;
        if (hgpu_check(cudaMemcpyAsync(b0 + s_core.off, core, n * 48, cudaMemcpyHostToDevice, st), "H2D") ||
            hgpu_check(cudaMemcpyAsync(b0 + s_data.off, data, data_bytes, cudaMemcpyHostToDevice, st), "H2D") ||
            hgpu_check(cudaMemcpyAsync(b0 + s_doff.off, data_off, (n + 1) * 8, cudaMemcpyHostToDevice, st), "H2D") ||
            hgpu_check(cudaMemcpyAsync(b0 + s_tl.off, tl.data(), n * 4, cudaMemcpyHostToDevice, st), "H2D")) return HGPU_ERR_CUDA;
        if (use_ref && (hgpu_check(cudaMemcpyAsync(b0 + s_ref.off, refs->bases, ref_bytes, cudaMemcpyHostToDevice, st), "H2D") ||
                        hgpu_check(cudaMemcpyAsync(b0 + s_roff.off, refs->off, ((size_t)refs->n_ref + 1) * 8, cudaMemcpyHostToDevice, st), "H2D"))) return HGPU_ERR_CUDA;
#endif
        EArgs A;
        A.core = reinterpret_cast<const Core *>(b0 + s_core.off); A.data = b0 + s_data.off; A.data_off = reinterpret_cast<const uint64_t *>(b0 + s_doff.off);
        A.tl = reinterpret_cast<const int32_t *>(b0 + s_tl.off); A.n = n; A.rps = rps;
        A.ref_bases = use_ref ? b0 + s_ref.off : nullptr; A.ref_off = reinterpret_cast<const uint64_t *>(b0 + s_roff.off); A.n_ref = use_ref ? refs->n_ref : 0;
        A.cnt = reinterpret_cast<uint32_t *>(b0 + s_cnt.off); A.tot = reinterpret_cast<uint32_t *>(b0 + s_tot.off);
        A.base = reinterpret_cast<const uint64_t *>(b0 + s_base.off); A.arena = b0 + s_arena.off; A.status = reinterpret_cast<int32_t *>(b0 + s_st.off);
#ifdef HGPU_HOSTSIM
        for (uint64_t g = 0; g < n; g++) count_body(A, g);
        for (size_t row = 0; row < rows; row++) {
            const uint64_t first = (uint64_t)(row / S_COUNT) * rps;
            const uint32_t nr = (uint32_t)(n - first < rps ? n - first : rps);
            uint32_t *p = A.cnt + row * rps, run = 0;
            for (uint32_t i = 0; i < nr; i++) { const uint32_t v = p[i]; p[i] = run; run += v; }
            A.tot[row] = run;
        }
        memcpy(tot.data(), A.tot, rows * 4);
        memcpy(status.data(), A.status, n * 4);
#else
        cram_enc_count_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A);
        cram_enc_scan_kernel<<<(unsigned)((rows + 3) / 4), 128, 0, st>>>(A, (uint32_t)rows);
        hgpu_count_launch(2);
        if (hgpu_check(cudaGetLastError(), "cram encode launch")) return HGPU_ERR_CUDA;
        if (hgpu_check(cudaMemcpyAsync(tot.data(), A.tot, rows * 4, cudaMemcpyDeviceToHost, st), "D2H") ||
            hgpu_check(cudaMemcpyAsync(status.data(), A.status, n * 4, cudaMemcpyDeviceToHost, st), "D2H") ||
            hgpu_check(cudaStreamSynchronize(st), "cram encode count")) return HGPU_ERR_CUDA;
#endif
        for (uint64_t g = 0; g < n; g++)
            if (status[g] != ENC_OK) {
                hgpu_set_error("cram encode: record %llu cannot be written by this encoder (%s)", (unsigned long long)g,
                               status[g] == ENC_UNSUPPORTED ? "a mapped read without SEQ / position, or a zero-length CIGAR op: host library" : "malformed bam1_t");
                return status[g] == ENC_UNSUPPORTED ? HGPU_CRAM_UNSUPPORTED : HGPU_CRAM_ERR_DECODE;
            }
        uint64_t at = 0;
        for (size_t row = 0; row < rows; row++) { base[row] = at; at += ((uint64_t)tot[row] + 15) & ~15ull; }
        if (at > arena_cap) { hgpu_set_error("cram encode: series arena bound exceeded"); return HGPU_ERR_NOMEM; }
        arena_h.resize(at + 16);
#ifdef HGPU_HOSTSIM
        memcpy(b0 + s_base.off, base.data(), rows * 8);
        for (uint64_t g = 0; g < n; g++) write_body(A, g);
        memcpy(arena_h.data(), A.arena, at);
#else
        if (hgpu_check(cudaMemcpyAsync(b0 + s_base.off, base.data(), rows * 8, cudaMemcpyHostToDevice, st), "H2D")) return HGPU_ERR_CUDA;
        cram_enc_write_kernel<<<(unsigned)((n + 127) / 128), 128, 0, st>>>(A);
        hgpu_count_launch();
        if (hgpu_check(cudaGetLastError(), "cram encode write launch")) return HGPU_ERR_CUDA;
        if (at && hgpu_check(cudaMemcpyAsync(arena_h.data(), A.arena, at, cudaMemcpyDeviceToHost, st), "D2H")) return HGPU_ERR_CUDA;
        if (hgpu_check(cudaStreamSynchronize(st), "cram encode write")) return HGPU_ERR_CUDA;
#endif
    }
    // ---- compress the series blocks (device codecs) ----
    struct Blk { uint32_t slice; int stream; int method; std::vector<uint8_t> comp; uint32_t usize; };
    std::vector<Blk> blks;
    for (uint32_t sl = 0; sl < ns; sl++)
        for (int s = 0; s < S_COUNT; s++)
            if (tot[(size_t)sl * S_COUNT + s]) blks.push_back({sl, s, 0, {}, tot[(size_t)sl * S_COUNT + s]});
#ifndef HGPU_HOSTSIM
    if (!blks.empty()) {
        // names through the tok3 encoder (CRAM 3.1); everything through the method trial; the smaller wins
        const uint32_t nb = (uint32_t)blks.size();
        std::vector<const uint8_t *> pay(nb);
        std::vector<uint32_t> plen(nb), mask(nb);
        std::vector<int32_t> cid(nb), chosen(nb);
        std::vector<uint8_t> ctype(nb, 4);
        uint64_t cap = 0;
        const uint32_t m31 = (1u << 5) | (1u << 17) | (1u << 18) | (1u << 20) | (1u << 21) | (1u << 22) | (1u << 23);    // RANS_PR0/1/64/128/129/192/193
        const uint32_t m30 = (1u << 4) | (1u << 16);                                                                  // RANS0 / RANS1
        for (uint32_t i = 0; i < nb; i++) {
            pay[i] = arena_h.data() + base[(size_t)blks[i].slice * S_COUNT + blks[i].stream];
            plen[i] = blks[i].usize; mask[i] = minor ? m31 : m30; cid[i] = blks[i].stream + 1;
            cap += (uint64_t)plen[i] + 64;
        }
        std::vector<uint8_t> framed(cap + 4096);
        std::vector<uint64_t> foff(nb);
        uint64_t flen = 0;
        int rc = hgpu_cram_compress_blocks_host(ctx, pay.data(), plen.data(), mask.data(), cid.data(), ctype.data(), nb, framed.data(), framed.size(), foff.data(), &flen, chosen.data());
        if (rc) return rc;
        for (uint32_t i = 0; i < nb; i++) {
            // un-frame: method, type, id, comp size, uncomp size, payload (the CRC is rebuilt when the container is laid out)
            const uint8_t *p = framed.data() + foff[i];
            blks[i].method = p[0];
            const uint8_t *q = p + 2;
            auto rd = [&](void) { uint32_t c = *q; int k = c < 0x80 ? 0 : c < 0xc0 ? 1 : c < 0xe0 ? 2 : c < 0xf0 ? 3 : 4; uint32_t v;
                switch (k) { case 0: v = c; break; case 1: v = ((c & 0x3f) << 8) | q[1]; break; case 2: v = ((c & 0x1f) << 16) | (q[1] << 8) | q[2]; break;
                             case 3: v = ((c & 0x0f) << 24) | (q[1] << 16) | (q[2] << 8) | q[3]; break;
                             default: v = ((c & 0x0f) << 28) | (q[1] << 20) | (q[2] << 12) | (q[3] << 4) | (q[4] & 0x0f); break; }
                q += k + 1; return v; };
            rd();
            const uint32_t csz = rd();
            rd();
            blks[i].comp.assign(q, q + csz);
        }
        if (minor == 1) {
            std::vector<uint32_t> idx;
            for (uint32_t i = 0; i < nb; i++) if (blks[i].stream == S_RN) idx.push_back(i);
            if (!idx.empty()) {
                const uint32_t k = (uint32_t)idx.size();
                std::vector<uint64_t> ioff(k), ooff(k);
                std::vector<uint32_t> ilen(k), ocap(k), olen(k);
                std::vector<int32_t> ost(k);
                uint64_t ipos = 0, opos = 0;
                for (uint32_t j = 0; j < k; j++) { ioff[j] = ipos; ilen[j] = blks[idx[j]].usize; ipos += ((uint64_t)ilen[j] + 15) & ~15ull;
                                                   ooff[j] = opos; ocap[j] = ilen[j] + ilen[j] / 2 + 65536; opos += ((uint64_t)ocap[j] + 15) & ~15ull; }
                std::vector<uint8_t> ibuf(ipos + 16), obuf(opos + 16);
                for (uint32_t j = 0; j < k; j++) memcpy(ibuf.data() + ioff[j], pay[idx[j]], ilen[j]);
                rc = hgpu_tok3_encode_batch_host(ctx, ibuf.data(), ioff.data(), ilen.data(), k, obuf.data(), ooff.data(), ocap.data(), olen.data(), ost.data());
                if (rc == HGPU_OK)
                    for (uint32_t j = 0; j < k; j++) {
                        Blk &b = blks[idx[j]];
                        const size_t cur = b.method == 0 ? b.usize : b.comp.size();
                        if (ost[j] == HGPU_OK && olen[j] && olen[j] < cur) { b.method = 8; b.comp.assign(obuf.data() + ooff[j], obuf.data() + ooff[j] + olen[j]); }
                    }
            }
        }
    }
#endif
    // ---- file image ----
    Buf f;
    { const uint8_t def[6] = {'C', 'R', 'A', 'M', 3, (uint8_t)minor}; f.bytes(def, 6); uint8_t id[20] = "htslib_b200"; f.bytes(id, 20); }
    auto container = [&](int32_t ref_id, int32_t start, int32_t span, int32_t nrec, int64_t counter, int64_t bases, int32_t nblocks,
                         const std::vector<int32_t> &landmarks, const Buf &body) {
        Buf h;
        h.le32((uint32_t)body.v.size());
        h.itf8(ref_id); h.itf8(start); h.itf8(span); h.itf8(nrec); h.ltf8(counter); h.ltf8(bases); h.itf8(nblocks);
        h.itf8((int32_t)landmarks.size());
        for (int32_t l : landmarks) h.itf8(l);
        h.le32(host_crc32(h.v.data(), h.v.size()));
        f.bytes(h.v.data(), h.v.size());
        f.bytes(body.v.data(), body.v.size());
    };
    {   // SAM header container (cram_write_SAM_hdr :4889): one RAW FILE_HEADER block = int32 length + text
        Buf pl; pl.le32(header_len); pl.bytes(header_text, header_len);
        Buf body; frame_block(body, 0, 0, 0, pl.v.data(), (uint32_t)pl.v.size(), (uint32_t)pl.v.size());
        container(0, 0, 0, 0, 0, 0, 1, std::vector<int32_t>{0}, body);
    }
    size_t bi = 0;
    for (uint32_t sl = 0; sl < ns; sl++) {
        const uint64_t a = (uint64_t)sl * rps, b = a + rps < n ? a + rps : n;
        int64_t bases = 0;
        for (uint64_t g = a; g < b; g++) bases += core[g].l_qseq;
        Buf ch; compression_header(ch, lines[sl], keys[sl], use_ref);
        Buf body;
        frame_block(body, 0, 1, 0, ch.v.data(), (uint32_t)ch.v.size(), (uint32_t)ch.v.size());
        const int32_t landmark = (int32_t)body.v.size();
        size_t e = bi;
        while (e < blks.size() && blks[e].slice == sl) e++;
        const int32_t next = (int32_t)(e - bi);
        Buf sh;                                                            // slice header (cram_encode_slice_header :2870)
        sh.itf8(-2); sh.itf8(0); sh.itf8(0); sh.itf8((int32_t)(b - a)); sh.ltf8((int64_t)a); sh.itf8(next + 1); sh.itf8(next + 1);
        sh.itf8(0);                                                        // content ids: the CORE block, then the external blocks
        for (size_t k = bi; k < e; k++) sh.itf8(blks[k].stream + 1);
        sh.itf8(-1);                                                       // no embedded reference
        { const uint8_t md5[16] = {0}; sh.bytes(md5, 16); }
        frame_block(body, 0, 2, 0, sh.v.data(), (uint32_t)sh.v.size(), (uint32_t)sh.v.size());
        frame_block(body, 0, 5, 0, nullptr, 0, 0);                         // CORE: every series is external
        for (size_t k = bi; k < e; k++) {
            const Blk &bk = blks[k];
            const uint8_t *raw = arena_h.data() + base[(size_t)sl * S_COUNT + bk.stream];
            if (bk.method == 0) frame_block(body, 0, 4, bk.stream + 1, raw, bk.usize, bk.usize);
            else frame_block(body, bk.method, 4, bk.stream + 1, bk.comp.data(), (uint32_t)bk.comp.size(), bk.usize);
        }
        container(-2, 0, 0, (int32_t)(b - a), (int64_t)a, bases, next + 3, std::vector<int32_t>{landmark}, body);
        bi = e;
    }
    { static const uint8_t eof[38] = {0x0f, 0x00, 0x00, 0x00, 0xff, 0xff, 0xff, 0xff, 0x0f, 0xe0, 0x45, 0x4f, 0x46, 0x00, 0x00, 0x00, 0x00, 0x01, 0x00,
                                      0x05, 0xbd, 0xd9, 0x4f, 0x00, 0x01, 0x00, 0x06, 0x06, 0x01, 0x00, 0x01, 0x00, 0x01, 0x00, 0xee, 0x63, 0x01, 0x4b};
      f.bytes(eof, 38); }                                                  // the CRAM 3 end-of-file container (CRAM specification, section 9)
    uint8_t *res = (uint8_t *)malloc(f.v.size() + 1);
    if (!res) { hgpu_set_error("out of host memory"); return HGPU_ERR_NOMEM; }
    memcpy(res, f.v.data(), f.v.size());
    *out_file = res; *out_len = f.v.size();
    return HGPU_OK;
}

}  // namespace

#ifdef HGPU_HOSTSIM
extern "C" int hostsim_cram_encode_records(const char *header_text, uint32_t header_len, const hgpu_bam1_core *core, const uint8_t *data,
        const uint64_t *data_off, uint64_t n, const hgpu_cram_refs *refs, uint32_t rps, int minor, uint8_t **out_file, uint64_t *out_len)
{
    try { return encode_impl(nullptr, header_text, header_len, core, data, data_off, n, refs, rps, minor, out_file, out_len); }
    catch (...) { hgpu_set_error("internal error"); return HGPU_ERR_NOMEM; }
}
#else
extern "C" int hgpu_cram_encode_records_host(hgpu_ctx *ctx, const char *header_text, uint32_t header_len, const hgpu_bam1_core *core,
        const uint8_t *data, const uint64_t *data_off, uint64_t n, const hgpu_cram_refs *refs, uint32_t records_per_slice, int minor_version,
        uint8_t **out_file, uint64_t *out_len)
{
    try { return encode_impl(ctx, header_text, header_len, core, data, data_off, n, refs, records_per_slice, minor_version, out_file, out_len); }
    catch (const std::bad_alloc &) { hgpu_set_error("out of host memory"); return HGPU_ERR_NOMEM; }
    catch (...) { hgpu_set_error("internal error"); return HGPU_ERR_NOMEM; }
}
#endif
