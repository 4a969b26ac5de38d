/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_rans_nx16.c header for the usage rule).
 *
 * CPU restatement of the rANS 4x8 decoder (CRAM 3.0 block method 4, "RANS") of htscodecs 1.6.6:
 *   rans_uncompress      rANS_static.c:840-850
 *   rans_uncompress_O0   rANS_static.c:221-384
 *   rans_uncompress_O1   rANS_static.c:599-827
 *   byte renormalisation rANS_byte.h:512-554 (L = 2^23, up to two bytes per state and step,
 *                        never reading past the input)
 * Parity status: PINNED — tests/test_oracle_rans4x8.py: all 8 golden streams of
 * htscodecs/tests/dat/r4x8 and the nine RANS blocks of a reference-written CRAM 3.0 file decode to
 * the expected bytes, and the oracle agrees with the compiled reference on seeded inputs.
 * Table oddities kept: a frequency byte of 0 means 4096 in order-1 tables (:672-673); a table may
 * sum to 4095, in which case slot 4095 repeats the last symbol (order 0, :299-305).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define L8 (1u << 23)

static void renorm8(uint32_t *x, const uint8_t **p, const uint8_t *end)
{
    if (*x >= L8 || *p >= end) return;
    *x = (*x << 8) | *(*p)++;
    if (*x < L8 && *p < end) *x = (*x << 8) | *(*p)++;
}

int orc_rans_4x8_decode(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size)
{
    const uint8_t *cp = in + 9, *end = in + in_size;
    uint32_t in_sz, out_sz, R[4], i, x;
    int order, z;
    if (in_size < 26) return -1;
    order = in[0];
    if (order > 1) return -1;
    in_sz = in[1] | in[2] << 8 | in[3] << 16 | (uint32_t)in[4] << 24;
    out_sz = in[5] | in[6] << 8 | in[7] << 16 | (uint32_t)in[8] << 24;
    if (in_sz != in_size - 9 || out_sz > *out_size) return -1;
    if (order && in_size < 27) return -1;
    if (!order) {
        uint8_t sym[4096]; uint16_t fr[4096], ba[4096];
        int j = *cp++, rle = 0;
        x = 0;
        do {
            uint32_t F, y;
            if (cp > end - 16) return -1;
            F = *cp++;
            if (F >= 128) F = ((F & 127) << 8) | *cp++;
            if (x + F > 4096) return -1;
            for (y = 0; y < F; y++) { sym[x + y] = (uint8_t)j; fr[x + y] = (uint16_t)F; ba[x + y] = (uint16_t)y; }
            x += F;
            if (!rle && j + 1 == *cp) { j = *cp++; rle = *cp++; }
            else if (rle) { rle--; if (++j > 255) return -1; }
            else j = *cp++;
        } while (j);
        if (x < 4095 || x > 4096) return -1;
        if (x != 4096) { sym[x] = sym[x - 1]; fr[x] = fr[x - 1]; ba[x] = ba[x - 1] + 1; }
        if (cp > end - 16) return -1;
        for (z = 0; z < 4; z++) { R[z] = cp[0] | cp[1] << 8 | cp[2] << 16 | (uint32_t)cp[3] << 24; cp += 4; if (R[z] < L8) return -1; }
        for (i = 0; i + 4 <= out_sz; i += 4) {
            uint32_t m[4];
            for (z = 0; z < 4; z++) { m[z] = R[z] & 4095; R[z] = fr[m[z]] * (R[z] >> 12) + ba[m[z]]; }
            for (z = 0; z < 4; z++) renorm8(&R[z], &cp, end);
            for (z = 0; z < 4; z++) out[i + z] = sym[m[z]];
        }
        for (z = 0; i + z < out_sz; z++) out[i + z] = sym[R[z] & 4095];
    } else {
        /* rows are numbered in order of first appearance as context or symbol (map[], :636-652) */
        int16_t map[256];
        uint8_t *lut = calloc(256, 4096);
        uint16_t (*sf)[256] = calloc(256, sizeof(*sf)), (*ss)[256] = calloc(256, sizeof(*ss));
        int mi = 0, ci, rle_i = 0, rc = -1;
        uint32_t l[4] = {0, 0, 0, 0}, pos[4], q;
        uint8_t c[4];
        if (!lut || !sf || !ss) goto done;
        memset(map, -1, sizeof(map));
        ci = *cp++;
        do {
            int row, j, rle_j = 0;
            if (map[ci] == -1) map[ci] = (int16_t)mi++;
            row = map[ci];
            x = 0;
            j = *cp++;
            do {
                uint32_t F;
                if (map[j] == -1) map[j] = (int16_t)mi++;
                if (cp > end - 16) goto done;
                F = *cp++;
                if (F >= 128) F = ((F & 127) << 8) | *cp++;
                if (!F) F = 4096;
                if (x + F > 4096) goto done;
                sf[row][j] = (uint16_t)F; ss[row][j] = (uint16_t)x;
                memset(lut + (size_t)row * 4096 + x, j, F);
                x += F;
                if (!rle_j && j + 1 == *cp) { j = *cp++; rle_j = *cp++; }
                else if (rle_j) { rle_j--; if (++j > 255) goto done; }
                else j = *cp++;
            } while (j);
            if (x < 4095 || x > 4096) goto done;
            if (!rle_i && ci + 1 == *cp) { ci = *cp++; rle_i = *cp++; }
            else if (rle_i) { rle_i--; if (++ci > 255) goto done; }
            else ci = *cp++;
        } while (ci);
        for (z = 0; z < 256; z++) if (map[z] == -1) map[z] = 0;
        if (cp > end - 16) goto done;
        for (z = 0; z < 4; z++) { R[z] = cp[0] | cp[1] << 8 | cp[2] << 16 | (uint32_t)cp[3] << 24; cp += 4; if (R[z] < L8) goto done; }
        q = out_sz >> 2;
        for (z = 0; z < 4; z++) { pos[z] = z * q; l[z] = (uint32_t)map[0]; }
        for (i = 0; i < q; i++) {
            for (z = 0; z < 4; z++) {
                uint32_t m = R[z] & 4095;
                c[z] = lut[(size_t)l[z] * 4096 + m];
                out[pos[z]++] = c[z];
                R[z] = sf[l[z]][c[z]] * (R[z] >> 12) + m - ss[l[z]][c[z]];
            }
            for (z = 0; z < 4; z++) { renorm8(&R[z], &cp, end); l[z] = (uint32_t)map[c[z]]; }
        }
        while (pos[3] < out_sz) {
            uint32_t m = R[3] & 4095;
            uint8_t c3 = lut[(size_t)l[3] * 4096 + m];
            out[pos[3]++] = c3;
            R[3] = sf[l[3]][c3] * (R[3] >> 12) + m - ss[l[3]][c3];
            renorm8(&R[3], &cp, end);
            l[3] = (uint32_t)map[c3];
        }
        rc = 0;
    done:
        free(lut); free(sf); free(ss);
        if (rc) return -1;
    }
    *out_size = out_sz;
    return 0;
}
