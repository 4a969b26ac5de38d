// Host-side walk of a CRAM 3.x file image down to its blocks — the container/block framing of
// cram_read_container (cram/cram_io.c:3760-3900) and cram_read_block (:1414-1483), so that the
// payloads of every entropy-coded block can be handed to the batch decoders in one launch
// (what cram_decode_slice does block by block through cram_uncompress_block, cram_io.c:1576).
// Pure framing: ITF8/LTF8 integers, no decompression, no CRC.
#include "hgpu_internal.h"
#include <string.h>

namespace {

// ITF8 (cram_io.c:138-200): 1-5 bytes, length in the leading ones of the first byte
int itf8(const uint8_t *p, const uint8_t *end, int32_t *v)
{
    if (p >= end) return 0;
    uint8_t b = p[0];
    int n = b < 0x80 ? 1 : b < 0xc0 ? 2 : b < 0xe0 ? 3 : b < 0xf0 ? 4 : 5;
    if (end - p < n) return 0;
    uint32_t u;
    switch (n) {
    case 1: u = b; break;
    case 2: u = ((b & 0x3fu) << 8) | p[1]; break;
    case 3: u = ((b & 0x1fu) << 16) | (p[1] << 8) | p[2]; break;
    case 4: u = ((b & 0x0fu) << 24) | (p[1] << 16) | (p[2] << 8) | p[3]; break;
    default: u = ((b & 0x0fu) << 28) | (p[1] << 20) | (p[2] << 12) | (p[3] << 4) | (p[4] & 0x0f); break;
    }
    *v = (int32_t)u;
    return n;
}

// LTF8 (cram_io.c:~400): 1-9 bytes; only the length matters here
int ltf8_len(const uint8_t *p, const uint8_t *end)
{
    if (p >= end) return 0;
    uint8_t b = p[0];
    int n = b < 0x80 ? 1 : b < 0xc0 ? 2 : b < 0xe0 ? 3 : b < 0xf0 ? 4 : b < 0xf8 ? 5 : b < 0xfc ? 6 : b < 0xfe ? 7 : b < 0xff ? 8 : 9;
    return end - p < n ? 0 : n;
}

} // namespace

extern "C" long hgpu_cram_scan_blocks(const uint8_t *file, uint64_t len, hgpu_cram_block *blocks, long cap,
                                      int *major, int *minor)
{
    if (!file || len < 26 || memcmp(file, "CRAM", 4) != 0) { hgpu_set_error("not a CRAM file"); return -1; }
    int maj = file[4], min = file[5];
    if (major) *major = maj;
    if (minor) *minor = min;
    if (maj != 3) { hgpu_set_error("CRAM major version %d not supported (3.x only)", maj); return -1; }
    const uint8_t *end = file + len, *p = file + 26;
    long n = 0;
    uint32_t container = 0;
    while (p < end) {
        if (end - p < 4) { hgpu_set_error("truncated container header"); return -1; }
        int32_t clen = (int32_t)(p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24), v, nblk, nland;
        p += 4;
        int k;
        for (int f = 0; f < 4; f++) { if (!(k = itf8(p, end, &v))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k; }       // ref id, start, span, n records
        if (!(k = ltf8_len(p, end))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;                                          // record counter
        if (!(k = ltf8_len(p, end))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;                                          // bases
        if (!(k = itf8(p, end, &nblk))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;
        if (!(k = itf8(p, end, &nland))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;
        for (int f = 0; f < nland; f++) { if (!(k = itf8(p, end, &v))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k; }
        if (end - p < 4) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; }
        p += 4;                                                                                  // container CRC32
        if (clen < 0 || (uint64_t)(end - p) < (uint64_t)clen) { hgpu_set_error("container %u runs past the file", container); return -1; }
        const uint8_t *cend = p + clen;
        while (p < cend) {
            if (cend - p < 2) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; }
            hgpu_cram_block b;
            const uint8_t *hdr = p;
            b.method = p[0]; b.content_type = p[1]; b.container = container;
            p += 2;
            int32_t cs, us;
            if (!(k = itf8(p, cend, &b.content_id))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;
            if (!(k = itf8(p, cend, &cs))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;
            if (!(k = itf8(p, cend, &us))) { hgpu_set_error("truncated or malformed CRAM container/block header"); return -1; } p += k;
            if (cs < 0 || us < 0 || (uint64_t)(cend - p) < (uint64_t)cs + 4) { hgpu_set_error("block runs past its container"); return -1; }
            b.data_off = (uint64_t)(p - file);
            b.hdr_len = (uint16_t)(p - hdr);
            b.comp_size = (uint32_t)cs; b.uncomp_size = (uint32_t)us;
            if (blocks && n < cap) blocks[n] = b;
            n++;
            p += cs + 4;                                                                         // payload + block CRC32
        }
        container++;
    }
    return n;
}
