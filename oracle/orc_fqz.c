/* TEST INFRASTRUCTURE — CPU oracle for the fqzcomp quality codec (CRAM 3.1 block method 7), decode side.
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file.
 *
 * Restates htscodecs/htscodecs/fqzcomp_qual.c:
 *   uncompress_block_fqz2f :1456-1613   (stream layout, record loop, reversal pass)
 *   decompress_new_read    :1381-1453   (selector, length, reverse flag, duplicate flag)
 *   fqz_update_ctx         :344-386     (16-bit context from qualities, position, delta, selector)
 *   fqz_read_parameters(1) :1241-1379, read_array :146-190
 * with the range coder of c_range_coder.h:62-164 and the adaptive model of c_simple_model.h:85-169.
 * One model routine serves every alphabet size: a model is {total, slots[n] = (freq, symbol)} kept
 * approximately sorted by a swap with the predecessor, halved when the total passes 65519.
 *
 * Parity pinned: tests/test_oracle_fqz.py — the 16 golden streams of htscodecs/tests/dat/fqzcomp/ decode to
 * column 1 of dat/q* (as tests/fqzcomp.test checks), and oracle == compiled reference on seeded streams of
 * all four strategies and on corrupted streams.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define TOPV (1u << 24)
#define MAXF ((1u << 16) - 17)
#define STEPV 16u
#define NCTX 65536u

typedef struct { const uint8_t *p, *end; uint32_t range, code; int err; } rc_t;
typedef struct { uint32_t tot; int n; uint16_t f[257], s[257]; } mdl_t;     /* f[n] = 0 terminates the halving loop */

static void rc_init(rc_t *r, const uint8_t *p, const uint8_t *end)
{
    r->range = 0xffffffffu; r->code = 0; r->err = 0; r->p = p; r->end = end;
    if (p + 5 > end) { r->p = end; return; }
    for (int i = 0; i < 5; i++) r->code = (r->code << 8) | *r->p++;
}

static void mdl_init(mdl_t *m, int nsym)
{
    m->n = nsym; m->tot = (uint32_t)nsym;
    for (int i = 0; i < nsym; i++) { m->f[i] = 1; m->s[i] = (uint16_t)i; }
    m->f[nsym] = 0; m->s[nsym] = 0;
}

/* SIMPLE_MODEL_decodeSymbol (:135-169).  A cumulative frequency beyond the live slots is the reference's
 * "walked off the end" error: it returns 0 without touching the model or the coder. */
static unsigned mdl_get(mdl_t *m, rc_t *r)
{
    uint32_t tot = m->tot;
    uint32_t freq = (tot && r->range >= tot) ? r->code / (r->range /= tot) : 0;
    if (freq > MAXF) return 0;
    uint32_t acc = 0;
    int i = 0;
    while (i < m->n && acc + m->f[i] <= freq) acc += m->f[i++];
    if (i >= m->n) return 0;
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

/* read_array (:146-190) */
static int rd_array(const uint8_t *in, size_t n, uint32_t *a, int size)
{
    uint8_t R[1024];
    int i, j, z, last = -1;
    if (size > 1024) size = 1024;
    for (i = j = z = 0; z < size && (size_t)i < n; i++) {
        int run = in[i];
        R[j++] = (uint8_t)run;
        z += run;
        if (run == last) {
            if ((size_t)i + 1 >= n) return -1;
            int copy = in[++i];
            z += run * copy;
            while (copy-- && z <= size && j < 1024) R[j++] = (uint8_t)run;
        }
        if (j >= 1024) return -1;
        last = run;
    }
    int used = i, rmax = j;
    for (i = j = z = 0; j < size; i++) {
        int len = 0, part;
        if (z >= rmax) return -1;
        do { part = R[z++]; len += part; } while (part == 255 && z < rmax);
        if (part == 255) return -1;
        while (len && j < size) { len--; a[j++] = (uint32_t)i; }
    }
    return used;
}

typedef struct {
    uint32_t context, qmask, qshift, qloc, sloc, do_sel, fixed_len, do_dedup, max_sym;
    uint32_t qtab[256], ptab[1024], dtab[256];
    uint8_t qmap[256];
} par_t;

static int rd_param(par_t *p, const uint8_t *in, size_t n)                     /* :1241-1322 */
{
    if (n < 7) return -1;
    size_t k = 0;
    p->context = in[0] | in[1] << 8; k = 2;
    uint32_t fl = in[k++];
    p->do_sel = fl & 8; p->fixed_len = fl & 4; p->do_dedup = fl & 2;
    p->max_sym = in[k++];
    uint32_t qbits = in[k] >> 4;
    p->qmask = (1u << qbits) - 1; p->qshift = in[k++] & 15;
    p->qloc = in[k] >> 4; p->sloc = in[k++] & 15;
    uint32_t ploc = in[k] >> 4, dloc = in[k++] & 15;
    if (fl & 16) {
        memset(p->qmap, 0xff, 256);
        if (k + p->max_sym > n) return -1;
        for (uint32_t i = 0; i < p->max_sym; i++) p->qmap[i] = in[k++];
    } else for (int i = 0; i < 256; i++) p->qmap[i] = (uint8_t)i;
    for (int i = 0; i < 256; i++) p->qtab[i] = (uint32_t)i;
    if (qbits && (fl & 128)) { int u = rd_array(in + k, n - k, p->qtab, 256); if (u < 0) return -1; k += u; }
    memset(p->ptab, 0, sizeof(p->ptab));
    if (fl & 32) { int u = rd_array(in + k, n - k, p->ptab, 1024); if (u < 0) return -1; k += u; }
    memset(p->dtab, 0, sizeof(p->dtab));
    if (fl & 64) { int u = rd_array(in + k, n - k, p->dtab, 256); if (u < 0) return -1; k += u; }
    for (int i = 0; i < 1024; i++) p->ptab[i] <<= ploc;
    for (int i = 0; i < 256; i++) p->dtab[i] <<= dloc;
    return (int)k;
}

/* fqz_decompress: returns a malloc'd buffer of *out_size bytes or NULL */
uint8_t *orc_fqz_decode(const uint8_t *in, size_t in_size, size_t *out_size)
{
    uint32_t len;
    size_t k = (size_t)vget(in, in + in_size, &len);
    if (in_size < k || in_size - k < 10) return NULL;
    const uint8_t *q = in + k;
    size_t qn = in_size - k, j = 0;
    if (q[j++] != 5) return NULL;
    uint32_t gflags = q[j++];
    int nparam = (gflags & 1) ? q[j++] : 1;
    if (nparam <= 0) return NULL;
    uint32_t max_sel = nparam > 1 ? (uint32_t)nparam : 0, stab[256];
    if (gflags & 2) {
        max_sel = q[j++];
        int u = rd_array(q + j, qn - j, stab, 256);
        if (u < 0) return NULL;
        j += u;
    } else for (int i = 0; i < 256; i++) stab[i] = i < nparam ? (uint32_t)i : (uint32_t)nparam - 1;
    par_t *P = calloc((size_t)nparam, sizeof(par_t));
    mdl_t *qual = NULL;
    uint8_t *out = NULL, *rev = NULL, *ret = NULL;
    uint32_t *rlen = NULL;
    if (!P) return NULL;
    uint32_t gmax = 0;
    for (int i = 0; i < nparam; i++) {
        int u = j <= qn ? rd_param(&P[i], q + j, qn - j) : -1;
        if (u < 0 || (P[i].do_sel && max_sel == 0)) goto done;
        j += u;
        if (P[i].max_sym > gmax) gmax = P[i].max_sym;
    }
    qual = malloc(sizeof(mdl_t) * NCTX);
    out = malloc(len ? len : 1);
    rev = malloc((size_t)len + 2);
    rlen = malloc(((size_t)len + 2) * sizeof(uint32_t));
    if (!qual || !out || !rev || !rlen) goto done;
    for (uint32_t i = 0; i < NCTX; i++) mdl_init(&qual[i], (int)gmax + 1);
    mdl_t mlen[4], mrev, mdup, msel;
    for (int i = 0; i < 4; i++) mdl_init(&mlen[i], 256);
    mdl_init(&mrev, 2); mdl_init(&mdup, 2);
    mdl_init(&msel, max_sel > 0 ? (int)max_sel + 1 : 1);
    rc_t rc;
    rc_init(&rc, q + j, in + in_size);

    const par_t *pm0 = &P[0];                      /* the record loop keeps block 0 for context and map (:1541: pm by value) */
    uint32_t qctx = 0, p = 0, delta = 0, prevq = 0, sel = 0, first_len = 1, last_len = 0, last = 0, nrec = 0;
    uint32_t i = 0;
    int cur_rev = 0;
    while (i < len) {
        if (p == 0) {
            sel = pm0->do_sel ? mdl_get(&msel, &rc) : 0;
            uint32_t x = (gflags & 2) ? stab[sel < 255 ? sel : 255] : sel;
            if (x >= (uint32_t)nparam) goto done;
            const par_t *pm = &P[x];
            uint32_t rl = last_len;
            if (!pm->fixed_len || first_len) {
                rl = mdl_get(&mlen[0], &rc);
                rl |= mdl_get(&mlen[1], &rc) << 8;
                rl |= mdl_get(&mlen[2], &rc) << 16;
                rl |= mdl_get(&mlen[3], &rc) << 24;
                first_len = 0; last_len = rl;
            }
            if (rl > len - i || rl == 0) goto done;
            if (gflags & 4) { cur_rev = (int)mdl_get(&mrev, &rc); rev[nrec] = (uint8_t)cur_rev; rlen[nrec] = rl; }
            nrec++;
            if (pm->do_dedup && mdl_get(&mdup, &rc)) {
                if (rl > i) goto done;
                memcpy(out + i, out + i - rl, rl);
                i += rl;
                p = 0;
                continue;
            }
            p = rl; delta = 0; prevq = 0; qctx = 0;
            last = pm->context;
        }
        do {
            uint32_t Q = mdl_get(&qual[last], &rc) & 0xff;
            qctx = (qctx << pm0->qshift) + pm0->qtab[Q];
            uint32_t c = (qctx & pm0->qmask) << pm0->qloc;
            c += pm0->ptab[p < 1023 ? p : 1023];
            c += pm0->dtab[delta < 255 ? delta : 255];
            c += sel << pm0->sloc;
            delta += prevq != Q;
            prevq = Q;
            p--;
            last = c & (NCTX - 1);
            out[i++] = pm0->qmap[Q];
        } while (p != 0 && i < len);
    }
    if (gflags & 4) {                              /* stored reversed: undo per record (:1566-1580) */
        uint32_t a = 0;
        for (uint32_t r = 0; r < nrec && a < len; a += rlen[r++]) {
            if (!rev[r]) continue;
            for (uint32_t I = a, J = a + rlen[r] - 1; I < J; I++, J--) { uint8_t t = out[I]; out[I] = out[J]; out[J] = t; }
        }
    }
    if (rc.err < 0) goto done;
    ret = out; out = NULL;
    *out_size = len;
done:
    free(P); free(qual); free(out); free(rev); free(rlen);
    return ret;
}
