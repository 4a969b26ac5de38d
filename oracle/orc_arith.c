/* TEST INFRASTRUCTURE — CPU oracle for the adaptive arithmetic coder ("ARITH_PR", CRAM 3.1 block method 6),
 * decode side.  Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file.
 *
 * Restates htscodecs/htscodecs/arith_dynamic.c:
 *   arith_uncompress_to         :1033-1278  (format byte, size, PACK / RLE / CAT / NOSZ wrappers, STRIPE)
 *   arith_uncompress_O0 / _O1   :137-165, :227-272
 *   arith_uncompress_O0_RLE / _O1_RLE :520-580, :660-728
 * with hts_unpack_meta / hts_unpack (pack.c:161-330), unstripe (utils.h:79-138), the range coder
 * (c_range_coder.h:62-164) and the adaptive model (c_simple_model.h:85-169).  The four symbol loops are
 * one routine here (order and run-length flags as parameters); one model routine serves the 256-symbol
 * byte models and the 258-symbol run models.  X_EXT (bzip2 payload) is an error, as in a reference built
 * without libbz2 (:1218-1226), which is how oracle/_ref is built.
 *
 * Parity pinned: tests/test_oracle_arith.py — the golden streams of htscodecs/tests/dat/arith/ decode to
 * their raw inputs (as tests/arith.test checks), and oracle == compiled reference on seeded streams for
 * every flag combination and on corrupted streams.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TOPV (1u << 24)
#define MAXF ((1u << 16) - 17)
#define STEPV 16u
#define NRUN 258
#define MAXRUN 4

typedef struct { const uint8_t *p, *end; uint32_t range, code; int err; } rc_t;
typedef struct { uint32_t tot; int n; uint16_t f[NRUN + 1], s[NRUN + 1]; } mdl_t;

static void rc_init(rc_t *r, const uint8_t *p, const uint8_t *end)
{
    r->range = 0xffffffffu; r->code = 0; r->err = 0; r->p = p; r->end = end;
    if (p + 5 > end) { r->p = end; return; }
    for (int i = 0; i < 5; i++) r->code = (r->code << 8) | *r->p++;
}

/* SIMPLE_MODEL_init(m, max_sym) over NSYM slots: only the first `live` slots have weight */
static void mdl_init(mdl_t *m, int live)
{
    m->n = live; m->tot = (uint32_t)live;
    for (int i = 0; i < live; i++) { m->f[i] = 1; m->s[i] = (uint16_t)i; }
    m->f[live] = 0; m->s[live] = 0;
}

static unsigned mdl_get(mdl_t *m, rc_t *r)
{
    uint32_t tot = m->tot;
    uint32_t freq = (tot && r->range >= tot) ? r->code / (r->range /= tot) : 0;
    if (freq > MAXF) return 0;
    uint32_t acc = 0;
    int i = 0;
    while (i < m->n && acc + m->f[i] <= freq) acc += m->f[i++];
    if (i >= m->n) return 0;                       /* walked past the live slots: the reference's error return */
    uint32_t f = m->f[i];
    r->code -= acc * r->range;
    r->range *= f;
    while (r->range < TOPV) {
        if (r->p >= r->end) { r->err = -1; break; }
        r->code = (r->code << 8) + *r->p++;
        r->range <<= 8;
    }
    m->f[i] = (uint16_t)(f + STEPV);
    m->tot = tot + STEPV;
    if (m->tot > MAXF) {
        uint32_t t = 0;
        for (int k = 0; m->f[k]; k++) { m->f[k] -= m->f[k] >> 1; t += m->f[k]; }
        m->tot = t;
    }
    unsigned sym = m->s[i];
    if (i > 0 && m->f[i] > m->f[i - 1]) {
        uint16_t tf = m->f[i], ts = m->s[i];
        m->f[i] = m->f[i - 1]; m->s[i] = m->s[i - 1];
        m->f[i - 1] = tf; m->s[i - 1] = ts;
    }
    return sym;
}

static int vget(const uint8_t *p, const uint8_t *end, uint32_t *v)            /* varint.h:267-299 */
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

/* the four symbol loops: in[0] = number of byte symbols (0 = 256), then the range coder's bytes */
static int symbols(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t U, int order1, int rle)
{
    int live = in[0] ? in[0] : 256;
    int nctx = order1 ? 256 : 1;
    mdl_t *bm = malloc(sizeof(mdl_t) * (size_t)nctx), *rm = rle ? malloc(sizeof(mdl_t) * NRUN) : NULL;
    int ret = -1;
    if (!bm || (rle && !rm)) goto done;
    for (int i = 0; i < nctx; i++) mdl_init(&bm[i], live);
    if (rle) for (int i = 0; i < NRUN; i++) mdl_init(&rm[i], MAXRUN);
    rc_t rc;
    rc_init(&rc, in + 1, in + n);
    uint32_t last = 0;
    for (uint32_t i = 0; i < U; i++) {
        uint32_t c = mdl_get(&bm[order1 ? last : 0], &rc) & 0xff;
        out[i] = (uint8_t)c;
        last = c;
        if (rle) {
            uint32_t run = 0, r, rctx = last;
            do {
                r = mdl_get(&rm[rctx], &rc);
                if (rctx == last) rctx = 256; else rctx += (rctx < NRUN - 1);
                run += r;
            } while (r == MAXRUN - 1 && run < U);
            while (run-- && i + 1 < U) out[++i] = (uint8_t)last;
        }
    }
    ret = rc.err < 0 ? -1 : 0;
done:
    free(bm); free(rm);
    return ret;
}

/* hts_unpack_meta (pack.c:161-196): symbols per byte and the symbol map; bytes used, 0 on failure */
static int unpack_meta(const uint8_t *d, uint32_t len, uint8_t *map, int *per_byte)
{
    if (!len) return 0;
    uint32_t n = d[0] ? d[0] : 256, j = 1, c = 0;
    if (n <= 1) *per_byte = 0;
    else if (n <= 2) *per_byte = 8;
    else if (n <= 4) *per_byte = 4;
    else if (n <= 16) *per_byte = 2;
    else { *per_byte = 1; return 1; }
    if (len <= 1) return 0;
    do { map[c++] = d[j++]; } while (c < n && j < len);
    return c < n ? 0 : (int)j;
}

int orc_arith_decode(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size);

static int plain(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size)
{
    const uint8_t *end = in + in_size;
    int fmt = *in++; in_size--;
    const int do_pack = fmt & 0x80, do_rle = fmt & 0x40, do_cat = fmt & 0x20, no_size = fmt & 0x10, do_ext = fmt & 0x04;
    const int order = fmt & 3;
    uint32_t osz;
    if (!no_size) { int s = vget(in, end, &osz); in += s; in_size -= (uint32_t)s; } else osz = *out_size;
    if (osz >= 0x7fffffffu || *out_size < osz) return -1;
    *out_size = osz;
    uint32_t t1_size = osz;
    uint8_t *tmp = NULL, *t1 = out;
    uint8_t map[16] = {0};
    int per_byte = 0, ret = -1;
    uint64_t unpacked = 0;
    if (do_pack) {
        tmp = malloc(osz ? osz : 1);
        if (!tmp) return -1;
        t1 = tmp;
        int ml = unpack_meta(in, in_size, map, &per_byte);
        if (!ml) goto done;
        unpacked = osz;
        in += ml; in_size -= (uint32_t)ml;
        uint32_t psz;
        int s = vget(in, end, &psz);
        in += s; in_size -= (uint32_t)s;
        if (psz > t1_size) goto done;
        t1_size = psz;
    }
    if (in_size) {
        if (do_cat) {
            if (t1_size > in_size || t1_size > *out_size) goto done;
            memcpy(t1, in, t1_size);
        } else if (do_ext) goto done;
        else if (symbols(in, in_size, t1, t1_size, order == 1, do_rle)) goto done;
    } else t1_size = 0;
    if (do_pack) {                                 /* hts_unpack (pack.c:207-330) */
        if (per_byte == 1) unpacked = t1_size;
        if (per_byte == 1) memcpy(out, t1, t1_size);
        else if (per_byte == 0) memset(out, map[0], (size_t)unpacked);
        else {
            int bits = per_byte == 8 ? 1 : per_byte == 4 ? 2 : 4;
            if ((unpacked + (uint64_t)per_byte - 1) / (uint64_t)per_byte > t1_size) goto done;
            for (uint64_t i = 0; i < unpacked; i++)
                out[i] = map[(t1[i / (uint64_t)per_byte] >> (bits * (int)(i % (uint64_t)per_byte))) & ((1 << bits) - 1)];
        }
        *out_size = (uint32_t)unpacked;
    } else *out_size = t1_size;
    ret = 0;
done:
    free(tmp);
    return ret;
}

/* arith_uncompress_to with a caller buffer: *out_size is the capacity on entry, the length on return. 0 / -1. */
int orc_arith_decode(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size)
{
    if (in_size == 0) return -1;
    if (!(in[0] & 0x08)) return plain(in, in_size, out, out_size);
    /* byte planes coded separately (:1041-1116) */
    const uint8_t *end = in + in_size;
    uint32_t ulen, off = 1;
    off += (uint32_t)vget(in + off, end, &ulen);
    if (off >= in_size) return -1;
    uint32_t N = in[off++];
    if (N < 1 || ulen != *out_size) return -1;
    uint32_t clen[256], ul[256], idx[256];
    uint64_t ctot = 0;
    for (uint32_t k = 0; k < N; k++) {
        ul[k] = ulen / N + ((ulen % N) > k);
        idx[k] = k ? idx[k - 1] + ul[k - 1] : 0;
        off += (uint32_t)vget(in + off, end, &clen[k]);
        ctot += clen[k];
        if (off > in_size || clen[k] > in_size || clen[k] < 1) return -1;
    }
    if (off + ctot > in_size) return -1;
    in_size = (uint32_t)(off + ctot);
    uint8_t *planes = malloc(ulen ? ulen : 1);
    if (!planes) return -1;
    for (uint32_t k = 0; k < N; k++) {
        uint32_t got = ul[k];
        if (in_size < off || orc_arith_decode(in + off, in_size - off, planes + idx[k], &got) || got != ul[k]) { free(planes); return -1; }
        off += clen[k];
    }
    for (uint32_t j = 0; j < ulen; j++) out[j] = planes[idx[j % N] + j / N];       /* unstripe */
    free(planes);
    *out_size = ulen;
    return 0;
}
