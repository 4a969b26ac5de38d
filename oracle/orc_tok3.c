/* TEST INFRASTRUCTURE — CPU oracle for the CRAM 3.1 read-name tokeniser ("tok3", block method 8).
 * Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may use this file.
 *
 * Restates the decode side of htscodecs/htscodecs/tokenise_name3.c:
 *   tok3_decode_names  :1679-1834   (container: header, descriptor walk, name loop)
 *   decode_name        :1023-1210   (one name from the token streams + an earlier name)
 *   decode_token_*     :331-460     (cursor reads on a descriptor)
 *   append_uint32_*    :233-296     (decimal output)
 * The entropy layer under it (rans_decode :1255 / arith_decode :1228) is injected: rANS Nx16 goes
 * to orc_rans_nx16_decode, the adaptive arithmetic coder to a caller-supplied function with the
 * arith_uncompress_to signature (tests pass oracle/_ref's).
 *
 * Parity pinned: tests/test_oracle_tok3.py checks every golden names/tok3/ * file against the
 * plain-text names, and against oracle/_ref's tok3_decode_names on seeded and corrupted inputs.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

int orc_rans_nx16_decode(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size);
typedef unsigned char *(*orc_arith_fn)(unsigned char *in, unsigned int in_size,
                                       unsigned char *out, unsigned int *out_size);

enum { T_TYPE = 0, T_ALPHA, T_CHAR, T_DIGITS0, T_DZLEN, T_DUP, T_DIFF, T_DIGITS, T_DDELTA,
       T_DDELTA0, T_MATCH, T_NOP, T_END };
#define TOK_MAX 128

typedef struct { uint8_t *p; size_t n, pos; } stream_t;
typedef struct { int32_t type, val, aux; } tokrec_t;           /* last_context_tok :131-135 */
typedef struct { size_t name; int ntok; tokrec_t *tok; } namerec_t;

/* varint.h:267-299 */
static int vget(const uint8_t *p, const uint8_t *end, uint32_t *v)
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

static int rd_byte(stream_t *s) { return s->pos < s->n ? s->p[s->pos++] : -1; }
static int rd_u32(stream_t *s, uint32_t *v)
{
    if (s->pos + 4 > s->n) return -1;
    const uint8_t *c = s->p + s->pos;
    *v = c[0] | c[1] << 8 | c[2] << 16 | (uint32_t)c[3] << 24;
    s->pos += 4;
    return 0;
}

/* fixed width: exactly `w` digits (w <= 9), leading zeros kept, high digits dropped (:233-247) */
static int put_fixed(char *o, uint32_t v, uint32_t w)
{
    static const uint32_t p10[10] = {1, 10, 100, 1000, 10000, 100000, 1000000, 10000000, 100000000, 1000000000};
    if (w == 0 || w > 9) return (int)w;       /* no case in the reference's switch: nothing written, l returned */
    o[0] = (char)(v / p10[w - 1] + '0');      /* the leading digit is NOT reduced mod 10 there */
    v %= p10[w - 1];
    for (uint32_t k = w - 1; k >= 1; k--) { o[k] = '0' + v % 10; v /= 10; }
    return (int)w;
}

/* variable width: no leading zeros, and NOTHING for zero (:249-296: the last line is "if (i) *cp++") */
static int put_var(char *o, uint32_t v)
{
    char t[10];
    int n = 0;
    while (v) { t[n++] = '0' + v % 10; v /= 10; }
    for (int k = 0; k < n; k++) o[k] = t[n - 1 - k];
    return n;
}

typedef struct {
    stream_t d[TOK_MAX * 16];
    int max_tok, n_names, max_names;
    namerec_t *nm;
    char *out;
} ctx_t;

/* decode_name :1023-1210.  >0 bytes written, 0 end of data, -1 error */
static long one_name(ctx_t *c, size_t at, long room)
{
    int t0 = rd_byte(&c->d[0]);
    int cnum = c->n_names++;
    if (cnum >= c->max_names) return -1;
    if (t0 < 0 || t0 >= c->max_tok * 16) return 0;
    uint32_t dist;
    if (rd_u32(&c->d[t0], &dist) < 0 || dist > (uint32_t)cnum) return -1;
    int pnum = cnum - (int)dist;
    char *name = c->out + at;
    namerec_t *me = &c->nm[cnum], *pv = &c->nm[pnum];
    me->name = at;

    if (t0 == T_DUP) {
        if (pnum == cnum) return -1;
        size_t l = strlen(c->out + pv->name);
        if ((long)(l + 1) >= room) return -1;
        memcpy(name, c->out + pv->name, l + 1);
        me->ntok = pv->ntok;
        me->tok = malloc((pv->ntok ? pv->ntok : 1) * sizeof(tokrec_t));
        if (!me->tok) return -1;
        memcpy(me->tok, pv->tok, pv->ntok * sizeof(tokrec_t));
        return (long)l + 1;
    }

    me->tok = calloc(TOK_MAX, sizeof(tokrec_t));
    if (!me->tok) return -1;
    me->ntok = 0;                                /* so that pnum == cnum never matches (:1071) */
    long len = 0;
    for (int k = 1; k < TOK_MAX && k < c->max_tok; k++) {
        stream_t *S = &c->d[k << 4];
        int tok = rd_byte(&S[T_TYPE]);
        uint32_t v, w;
        tokrec_t *m = &me->tok[k];
        const tokrec_t *q = (pv->tok && k < pv->ntok) ? &pv->tok[k] : NULL;
        int b;
        switch (tok) {
        case T_CHAR:
            if (len + 1 >= room) return -1;
            if ((b = rd_byte(&S[T_CHAR])) < 0) return -1;
            name[len] = (char)b;
            m->type = T_CHAR; m->val = name[len++];
            break;
        case T_ALPHA: {
            stream_t *a = &S[T_ALPHA];
            long max = room - len, n = 0;
            char ch;
            if (a->pos >= a->n) return -1;
            do { ch = (char)a->p[a->pos++]; name[len + n++] = ch; } while (ch && n < max && a->pos < a->n);
            n--;
            m->type = T_ALPHA; m->aux = (int32_t)len; m->val = (int32_t)n;
            len += n;
            break;
        }
        case T_DIGITS0:
            if ((b = rd_byte(&S[T_DZLEN])) < 0) return -1;
            w = (uint32_t)b;
            if (rd_u32(&S[T_DIGITS0], &v) < 0) return -1;
            if (len + 20 + (long)w >= room) return -1;
            len += put_fixed(name + len, v, w);
            m->type = T_DIGITS0; m->val = (int32_t)v; m->aux = (int32_t)w;
            break;
        case T_DDELTA0:
            if (!q) return -1;
            if ((b = rd_byte(&S[T_DDELTA0])) < 0) return -1;
            v = (uint32_t)b + (uint32_t)q->val;
            if (len + q->aux + 1 >= room) return -1;
            len += put_fixed(name + len, v, (uint32_t)q->aux);
            m->type = T_DIGITS0; m->val = (int32_t)v; m->aux = q->aux;
            break;
        case T_DIGITS:
            if (rd_u32(&S[T_DIGITS], &v) < 0) return -1;
            if (len + 20 >= room) return -1;
            len += put_var(name + len, v);
            m->type = T_DIGITS; m->val = (int32_t)v;
            break;
        case T_DDELTA:
            if (!q) return -1;
            if ((b = rd_byte(&S[T_DDELTA])) < 0) return -1;
            v = (uint32_t)b + (uint32_t)q->val;
            if (len + 20 >= room) return -1;
            len += put_var(name + len, v);
            m->type = T_DIGITS; m->val = (int32_t)v;
            break;
        case T_NOP:
            m->type = T_NOP;
            break;
        case T_MATCH:
            if (!q) return -1;
            switch (q->type) {
            case T_CHAR:
                if (len + 1 >= room) return -1;
                name[len++] = (char)q->val;
                m->type = T_CHAR; m->val = q->val;
                break;
            case T_ALPHA:
                if (q->val < 0 || len + q->val >= room) return -1;
                memcpy(name + len, c->out + pv->name + q->aux, (size_t)q->val);
                m->type = T_ALPHA; m->aux = (int32_t)len; m->val = q->val;
                len += q->val;
                break;
            case T_DIGITS:
                if (len + 20 >= room) return -1;
                len += put_var(name + len, (uint32_t)q->val);
                m->type = T_DIGITS; m->val = q->val;
                break;
            case T_DIGITS0:
                if (len + q->aux >= room) return -1;
                len += put_fixed(name + len, (uint32_t)q->val, (uint32_t)q->aux);
                m->type = T_DIGITS0; m->val = q->val; m->aux = q->aux;
                break;
            default:
                return -1;
            }
            break;
        default:                                   /* T_END, or a type stream that ran dry */
            if (len + 1 >= room) return -1;
            name[len++] = 0;
            m->type = T_END;
            me->ntok = k;
            return len;
        }
    }
    return -1;
}

/* tok3_decode_names :1679-1834.  Returns a malloc'd buffer of *out_len bytes or NULL. */
uint8_t *orc_tok3_decode(const uint8_t *in, uint32_t sz, uint32_t *out_len, orc_arith_fn arith)
{
    if (sz < 9) return NULL;
    int32_t ulen = (int32_t)(in[0] | in[1] << 8 | in[2] << 16 | (uint32_t)in[3] << 24);
    if (ulen < 0 || ulen >= INT32_MAX - 1024) return NULL;
    int32_t nreads = (int32_t)(in[4] | in[5] << 8 | in[6] << 16 | (uint32_t)in[7] << 24);
    int use_arith = in[8];
    if (nreads <= 0 || nreads > 10000000) return NULL;            /* create_context :172-187 */
    if (use_arith && !arith) return NULL;

    ctx_t *c = calloc(1, sizeof(*c));
    if (!c) return NULL;
    c->max_names = nreads + 1;
    c->max_tok = 1;
    c->nm = calloc((size_t)c->max_names, sizeof(namerec_t));
    uint8_t *ret = NULL;
    if (!c->nm) goto done;

    uint32_t o = 9;
    int tnum = -1;
    while (o < sz) {
        uint8_t tt = in[o++];
        int dup = tt & 64, j = 0;
        if (dup) {
            if (o + 2 > sz) goto done;
            j = in[o] << 4; j += in[o + 1]; o += 2;
        }
        if (tt & 128) {
            if (++tnum >= TOK_MAX) goto done;
            c->max_tok = tnum + 1;
            for (int k = 0; k < 16; k++) { free(c->d[(tnum << 4) + k].p); memset(&c->d[(tnum << 4) + k], 0, sizeof(stream_t)); }
        }
        if ((tt & 15) != 0 && (tt & 128)) {                      /* implied type stream: [type, MATCH, MATCH ...] */
            stream_t *t = &c->d[tnum << 4];
            free(t->p);
            t->p = malloc((size_t)nreads);
            if (!t->p) goto done;
            t->n = (size_t)nreads; t->pos = 0;
            memset(t->p, T_MATCH, t->n);
            t->p[0] = tt & 15;
        }
        if (tnum < 0) goto done;
        int i = (tnum << 4) | (tt & 15);
        stream_t *d = &c->d[i];
        if (dup) {
            if (j >= i || !c->d[j].p) goto done;
            uint8_t *np = malloc(c->d[j].n ? c->d[j].n : 1);
            if (!np) goto done;
            memcpy(np, c->d[j].p, c->d[j].n);
            free(d->p);
            d->p = np; d->n = c->d[j].n; d->pos = 0;
            continue;
        }
        /* uncompressed_size :1419-1430 then uncompress :1432-1438 */
        const uint8_t *s = in + o, *e = in + sz;
        uint32_t clen, usz;
        int nb = vget(s, e, &clen);
        vget(s + nb + 1 <= e ? s + nb + 1 : e, e, &usz);
        if ((int32_t)usz < 0 || usz >= INT32_MAX) goto done;
        free(d->p);
        d->p = malloc(usz ? usz : 1);
        if (!d->p) goto done;
        d->n = usz; d->pos = 0;
        uint32_t got = usz;
        if (use_arith) {
            unsigned int g = usz;
            if (!arith((unsigned char *)s + nb, sz - o - nb, d->p, &g)) goto done;
            got = g;
        } else if (orc_rans_nx16_decode(s + nb, sz - o - nb, d->p, &got) != 0) goto done;
        if (got != usz) goto done;
        o += clen + nb;
    }

    long room = (long)ulen + 1024;
    c->out = malloc((size_t)room);
    if (!c->out) goto done;
    size_t at = 0;
    long r;
    while ((r = one_name(c, at, room)) > 0) { at += (size_t)r; room -= r; }
    if (r == 0) { ret = (uint8_t *)c->out; c->out = NULL; *out_len = (uint32_t)at; }

done:
    for (int k = 0; k < TOK_MAX * 16; k++) free(c->d[k].p);
    if (c->nm) for (int k = 0; k < c->max_names; k++) free(c->nm[k].tok);
    free(c->nm);
    free(c->out);
    free(c);
    return ret;
}

void orc_free(void *p) { free(p); }
