// Host-side tables of a CRAM 3.x file image above the block level: the container headers
// (cram_read_container, cram/cram_io.c:3760-3900) and the slice headers (cram_decode_slice_header,
// cram/cram_decode.c:959-1046).  Containers and slices are the units that shard across GPUs with no
// exchange step (contiguous ranges of slices per rank, so every rank's output is a
// contiguous record range); a slice header also lists the content ids of its blocks, which together with
// the compression header's encoding map (cram_header.cu) tells a record decoder where every data series
// lives.  Pure framing: ITF8 / LTF8 integers; pinned against the reference's own accessors
// (htslib/cram.h:190-216, :418-441) in tests/test_cram_header.py.
#include "hgpu_internal.h"
#include <string.h>

namespace {

struct Rd {
    const uint8_t *p, *e;
    bool err = false;
    uint8_t byte() { if (p >= e) { err = true; return 0; } return *p++; }
    int32_t itf8()
    {
        if (p >= e) { err = true; return 0; }
        const uint8_t c = *p;
        const int n = c < 0x80 ? 0 : c < 0xc0 ? 1 : c < 0xe0 ? 2 : c < 0xf0 ? 3 : 4;
        if (e - p < n + 1) { err = true; p = e; return 0; }
        uint32_t v;
        switch (n) {
        case 0: v = c; break;
        case 1: v = ((c & 0x3fu) << 8) | p[1]; break;
        case 2: v = ((c & 0x1fu) << 16) | (p[1] << 8) | p[2]; break;
        case 3: v = ((c & 0x0fu) << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; break;
        default: v = ((c & 0x0fu) << 28) | (p[1] << 20) | (p[2] << 12) | (p[3] << 4) | (p[4] & 0x0f); break;
        }
        p += n + 1;
        return (int32_t)v;
    }
    int64_t ltf8()                                                   // ltf8_get, cram/cram_io.c: leading ones give the length
    {
        if (p >= e) { err = true; return 0; }
        const uint8_t c = *p;
        int n = 0;
        while (n < 8 && (c & (0x80u >> n))) n++;                      // n extra bytes
        if (e - p < n + 1) { err = true; p = e; return 0; }
        uint64_t v = n >= 8 ? 0 : (uint64_t)(c & (0xffu >> (n + 1)));
        if (n == 7) v = 0;
        for (int k = 1; k <= n; k++) v = (v << 8) | p[k];
        p += n + 1;
        return (int64_t)v;
    }
};

}  // namespace

extern "C" long hgpu_cram_scan_containers(const uint8_t *file, uint64_t len, hgpu_cram_container *out, long cap,
                                          int32_t *landmarks, long landmark_cap)
{
    if (!file || len < 26 || memcmp(file, "CRAM", 4) != 0) { hgpu_set_error("not a CRAM file"); return -1; }
    if (file[4] != 3) { hgpu_set_error("CRAM major version %d not supported (3.x only)", file[4]); return -1; }
    const uint8_t *end = file + len;
    Rd r{file + 26, end};
    long n = 0, nl = 0;
    uint32_t block = 0;
    while (r.p < end) {
        hgpu_cram_container c;
        memset(&c, 0, sizeof(c));
        c.offset = (uint64_t)(r.p - file);
        if (end - r.p < 4) { hgpu_set_error("truncated container header"); return -1; }
        c.length = (int32_t)(r.p[0] | r.p[1] << 8 | r.p[2] << 16 | (uint32_t)r.p[3] << 24);
        r.p += 4;
        c.ref_id = r.itf8(); c.start = r.itf8(); c.span = r.itf8(); c.n_records = r.itf8();
        c.record_counter = r.ltf8(); c.bases = r.ltf8();
        c.n_blocks = r.itf8(); c.n_landmarks = r.itf8();
        if (r.err || c.n_landmarks < 0 || c.n_blocks < 0) { hgpu_set_error("container %ld: bad header", n); return -1; }
        c.landmark0 = (uint32_t)nl;
        for (int32_t k = 0; k < c.n_landmarks; k++) {
            const int32_t v = r.itf8();
            if (r.err) break;                                        // a corrupt count must not spin through 2^31 reads
            if (landmarks && nl < landmark_cap) landmarks[nl] = v;
            nl++;
        }
        if (r.err || end - r.p < 4) { hgpu_set_error("container %ld: truncated header", n); return -1; }
        c.crc32 = r.p[0] | r.p[1] << 8 | r.p[2] << 16 | (uint32_t)r.p[3] << 24;
        r.p += 4;
        c.data_off = (uint64_t)(r.p - file);
        c.first_block = block;
        if (c.length < 0 || (uint64_t)(end - r.p) < (uint64_t)c.length) { hgpu_set_error("container %ld runs past the file", n); return -1; }
        if (out && n < cap) out[n] = c;
        n++;
        block += (uint32_t)c.n_blocks;
        r.p += c.length;
    }
    return n;
}

extern "C" long hgpu_cram_parse_slice_header(const uint8_t *payload, uint32_t len, int major_version,
                                             hgpu_cram_slice *out, int32_t *content_ids, long cap)
{
    if (!payload || !out || major_version != 3) { hgpu_set_error("slice header: CRAM 3.x only"); return -1; }
    Rd r{payload, payload + len};
    memset(out, 0, sizeof(*out));
    out->ref_id = r.itf8(); out->start = r.itf8(); out->span = r.itf8();
    out->n_records = r.itf8();
    out->record_counter = r.ltf8();
    out->n_blocks = r.itf8();
    out->n_content_ids = r.itf8();
    if (r.err || out->n_content_ids < 1 || out->n_content_ids >= 10000) { hgpu_set_error("slice header: bad content id count"); return -1; }
    if (len < (uint32_t)out->n_content_ids) { hgpu_set_error("slice header: truncated"); return -1; }
    for (int32_t k = 0; k < out->n_content_ids; k++) {
        const int32_t v = r.itf8();
        if (content_ids && k < cap) content_ids[k] = v;
    }
    out->ref_base_id = r.itf8();
    if (r.err || r.e - r.p < 16) { hgpu_set_error("slice header: truncated"); return -1; }
    memcpy(out->md5, r.p, 16);
    return out->n_content_ids;
}
