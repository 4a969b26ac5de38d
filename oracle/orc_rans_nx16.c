/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Never linked into, imported by or executed from the
 * product path (htslib_b200/).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may use it.
 *
 * CPU restatement of the rANS Nx16 ("RANS_PR", CRAM 3.1 method 5) decoder of htscodecs 1.6.6.
 * Parity status: PINNED — checked against every golden vector in
 * htscodecs/tests/dat/r4x16/ (tests/test_oracle.py) and against the unmodified reference
 * compiled into oracle/_ref/libhts_ref.so on seeded inputs.
 *
 * One routine parametrised by N (4 or 32 interleaved states) replaces the reference's four
 * hand-unrolled decoders.  Citations are relative to /root/reference/htscodecs/htscodecs/.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define L_BOUND (1u << 15)           /* RANS_BYTE_L, rANS_word.h:64 */

/* 7-bit big-endian varint, MSB = continuation (varint.h:267-299).  Reads at most 6 bytes
 * when >= 6 remain, otherwise stops at the end.  Returns bytes consumed (0 at end). */
static int vget(const uint8_t *p, const uint8_t *end, uint32_t *v)
{
    const uint8_t *s = p;
    uint32_t acc = 0;
    uint8_t c;
    if (end - p >= 6) {
        int budget = 5;
        do { c = *p++; acc = (acc << 7) | (c & 0x7f); } while ((c & 0x80) && budget-- > 0);
    } else {
        if (p >= end) { *v = 0; return 0; }
        do { c = *p++; acc = (acc << 7) | (c & 0x7f); } while ((c & 0x80) && p < end);
    }
    *v = acc;
    return (int)(p - s);
}

/* Alphabet: ascending symbols, a symbol equal to previous+1 is followed by a run count,
 * list terminated by 0 (rANS_static16_int.h:191-238).  present[] gets 1 per symbol.
 * Returns bytes consumed, 0 on error. */
static int read_alphabet(const uint8_t *p, const uint8_t *end, uint32_t *present)
{
    const uint8_t *s = p;
    int run = 0, sym;
    if (p >= end) return 0;
    sym = *p++;
    if (sym == 0 && p + 2 >= end)     /* "carefully" entry with j==0: nothing recorded (:201-205) */
        return (int)(p - s);
    for (;;) {
        present[sym] = 1;
        if (p >= end) return 0;
        if (run == 0 && sym + 1 == *p) {
            if (p + 1 >= end) return 0;
            sym = *p++;
            run = *p++;
        } else if (run) {
            run--;
            if (++sym > 255) return 0;
        } else {
            sym = *p++;
        }
        if (sym == 0) break;
        if (p >= end) {           /* reference's careful loop exits when cp == cp_end */
            break;
        }
    }
    return (int)(p - s);
}

/* Order-0 table: alphabet then one varint per present symbol (rANS_static16_int.h:254-272),
 * scaled up to 1<<12 by shifting (normalise_freq_shift :151-162). */
static int read_freq0(const uint8_t *p, const uint8_t *end, uint32_t *F)
{
    const uint8_t *s = p;
    uint32_t tot = 0;
    int n, j;
    if (p == end) return 0;
    n = read_alphabet(p, end, F);
    p += n;                                    /* n == 0 tolerated like the reference */
    for (j = 0; j < 256; j++)
        if (F[j]) { p += vget(p, end, &F[j]); tot += F[j]; }
    if (tot != 0 && tot != 4096) {
        int sh = 0;
        while (tot < 4096) { tot *= 2; sh++; }
        for (j = 0; j < 256; j++) F[j] <<= sh;
    }
    return (int)(p - s);
}

typedef struct { uint16_t f, b; uint8_t s; } slot_t;

/* The core symbol loop shared by every variant.
 *  order 0: symbol i is produced by state i % N (rANS_static32x16pr.c:312-399,
 *           rANS_static4x16pr.c:286-319).
 *  order 1: state z owns out[z*(U/N) .. (z+1)*(U/N)), the last state also owns the tail
 *           (rANS_static32x16pr.c:619-680, rANS_static4x16pr.c:632-689); context = previous
 *           symbol of the same state, 0 at the start.
 * Renormalisation is in state order; a word is consumed only if two bytes remain
 * (RansDecRenormSafe, rANS_word.h:441-449).
 * lut[ctx] points to (1<<shift) slots. */
static int symbol_loop(const uint8_t *p, const uint8_t *end, uint8_t *out, uint32_t U,
                       int N, int order, int shift, slot_t **lut)
{
    uint32_t R[32], pos[32], ctx[32];
    uint32_t mask = (1u << shift) - 1, seg = U / N, step, i;
    int z;
    if (end - p < 4 * N) return -1;
    for (z = 0; z < N; z++) {
        R[z] = p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24;
        p += 4;
        if (R[z] < L_BOUND) return -1;
        pos[z] = order ? (uint32_t)z * seg : (uint32_t)z;
        ctx[z] = 0;
    }
    if (order == 0) {
        for (i = 0; i < U; i++) {
            slot_t e;
            z = i % N;
            e = lut[0][R[z] & mask];
            out[i] = e.s;
            R[z] = e.f * (R[z] >> shift) + e.b;
            if (R[z] < L_BOUND && p + 2 <= end) { R[z] = (R[z] << 16) | p[0] | p[1] << 8; p += 2; }
        }
        return 0;
    }
    for (step = 0; step < seg; step++) {
        for (z = 0; z < N; z++) {
            slot_t e = lut[ctx[z]][R[z] & mask];
            out[pos[z]++] = e.s;
            ctx[z] = e.s;
            R[z] = e.f * (R[z] >> shift) + e.b;
            if (R[z] < L_BOUND && p + 2 <= end) { R[z] = (R[z] << 16) | p[0] | p[1] << 8; p += 2; }
        }
    }
    z = N - 1;
    while (pos[z] < U) {
        slot_t e = lut[ctx[z]][R[z] & mask];
        out[pos[z]++] = e.s;
        ctx[z] = e.s;
        R[z] = e.f * (R[z] >> shift) + e.b;
        if (R[z] < L_BOUND && p + 2 <= end) { R[z] = (R[z] << 16) | p[0] | p[1] << 8; p += 2; }
    }
    return 0;
}

/* slots for one context from frequencies summing to 1<<shift (rans_F_to_s3,
 * rANS_static16_int.h:540-551; decode_freq1 :497-520).  f_wrap reproduces the 32-bit packing
 * of the 32-way order-0 table where F<<(12+8) drops bit 12 (F==4096 -> f==0). */
static int fill_slots(slot_t *t, const uint32_t *F, int shift, int f_wrap)
{
    uint32_t x = 0, y;
    int j;
    for (j = 0; j < 256; j++) {
        if (!F[j]) continue;
        if (F[j] > (1u << shift) - x) return -1;
        for (y = 0; y < F[j]; y++, x++) {
            t[x].s = (uint8_t)j;
            t[x].f = (uint16_t)(f_wrap ? (F[j] & 0xfff) : F[j]);
            t[x].b = (uint16_t)y;
        }
    }
    return x == (1u << shift) ? 0 : -1;
}

static int dec_order0(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t U, int N)
{
    uint32_t F[256] = {0};
    slot_t *tab, *lut[1];
    const uint8_t *p = in, *end = in + in_size;
    int n, rc;
    if (in_size < 16) return -1;
    /* the 4-way decoder parses its table against in+in_size-8 (rANS_static4x16pr.c:228,241),
     * the 32-way one against the true end (rANS_static32x16pr.c:272,285) */
    n = read_freq0(p, N == 4 ? end - 8 : end, F);
    if (!n) return -1;
    p += n;
    tab = malloc(sizeof(slot_t) * 4096);
    if (!tab) return -1;
    if (fill_slots(tab, F, 12, N == 32)) { free(tab); return -1; }
    lut[0] = tab;
    rc = symbol_loop(p, end, out, U, N, 0, 12, lut);
    free(tab);
    return rc;
}

/* Order-1 table (decode_freq1, rANS_static16_int.h:468-536; decode_freq_d :425-456):
 * first the order-0 alphabet A; then for each context in A, for each symbol in A a varint
 * frequency where a zero is followed by a count of further zeros; each row is scaled up to
 * 1<<shift. */
static int dec_order1(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t U, int N)
{
    const uint8_t *p = in, *end = in + in_size, *tend, *after_tab = NULL;
    uint8_t *tbuf = NULL;
    uint32_t A[256] = {0};
    slot_t *store = NULL, *lut[256];
    int shift, n, i, j, rc = -1, nctx = 0;
    if (in_size < (uint32_t)(N == 4 ? 16 : 4 * N)) return -1;
    shift = *p >> 4;
    tend = end;
    if (*p++ & 1) {                  /* table itself is order-0 4-way rANS coded */
        uint32_t usz, csz;
        p += vget(p, end, &usz);
        p += vget(p, end, &csz);
        if (csz > (uint32_t)(end - p)) return -1;
        after_tab = p + csz;
        tbuf = malloc(usz ? usz : 1);
        if (!tbuf) return -1;
        if (dec_order0(p, csz, tbuf, usz, 4)) goto done;
        p = tbuf;
        tend = tbuf + usz;
    }
    if (shift != 10 && shift != 12) {
        /* the reference indexes tables sized for 10 or 12 only; other values are UB there */
        goto done;
    }
    n = read_alphabet(p, tend, A);
    if (!n) goto done;
    p += n;
    if (p >= tend) goto done;
    for (i = 0; i < 256; i++) nctx += A[i] != 0;
    store = calloc((size_t)(nctx + 1) << shift, sizeof(slot_t));
    if (!store) goto done;
    /* Contexts that are absent, or present with no frequencies, are never entered by a valid
     * stream; the reference reads whatever its TLS scratch held (rANS_static4x16pr.c:589-591
     * "continue").  The oracle pins that undefined case to a fixed row: symbol 0, f 1, b = slot
     * index, which is also what the CUDA decoder produces. */
    for (i = 0; i < (1 << shift); i++) { store[i].b = (uint16_t)i; store[i].f = 1; }
    for (i = 0; i < 256; i++) lut[i] = store;
    nctx = 1;
    for (i = 0; i < 256; i++) {
        uint32_t F[256] = {0}, T = 0;
        const uint8_t *q = p;
        int dz = 0;
        if (!A[i]) continue;
        if (q == tend) goto done;
        for (j = 0; j < 256 && q < tend; j++) {
            uint32_t f;
            if (!A[j]) continue;
            if (dz) { f = 0; dz--; }
            else {
                q += vget(q, tend, &f);
                if (f == 0) { if (q >= tend) goto done; dz = *q++; }
            }
            F[j] = f; T += f;
        }
        if (q == p) goto done;
        p = q;
        if (!T) continue;
        if (T != (1u << shift)) {
            int sh = 0; uint32_t t = T;
            while (t < (1u << shift)) { t *= 2; sh++; }
            for (j = 0; j < 256; j++) F[j] <<= sh;
        }
        lut[i] = store + ((size_t)nctx++ << shift);
        if (fill_slots(lut[i], F, shift, 0)) goto done;
    }
    if (after_tab) p = after_tab;
    rc = symbol_loop(p, end, out, U, N, 1, shift, lut);
done:
    free(store);
    free(tbuf);
    return rc;
}

/* hts_unpack_meta / hts_unpack (pack.c:161-196, :207-330) */
static int unpack_meta(const uint8_t *d, uint32_t len, uint8_t *map, int *per_byte)
{
    uint32_t n, j = 1, c = 0;
    if (!len) return 0;
    n = d[0] ? d[0] : 256;
    if (n <= 1) *per_byte = 0;
    else if (n <= 2) *per_byte = 8;
    else if (n <= 4) *per_byte = 4;
    else if (n <= 16) *per_byte = 2;
    else { *per_byte = 1; return 1; }
    if (len <= 1) return 0;
    do { map[c++] = d[j++]; } while (c < n && j < len);
    return c < n ? 0 : (int)j;
}

static int unpack(const uint8_t *d, uint64_t len, uint8_t *out, uint64_t olen, int per_byte,
                  const uint8_t *map)
{
    uint64_t i;
    int bits;
    switch (per_byte) {
    case 1: memcpy(out, d, len); return 0;
    case 0: memset(out, map[0], olen); return 0;
    case 8: bits = 1; break;
    case 4: bits = 2; break;
    case 2: bits = 4; break;
    default: return -1;
    }
    if ((olen + per_byte - 1) / per_byte > len) return -1;
    for (i = 0; i < olen; i++)
        out[i] = map[(d[i / per_byte] >> (bits * (i % per_byte))) & ((1 << bits) - 1)];
    return 0;
}

/* hts_rle_decode (rle.c:142-190) */
static int unrle(const uint8_t *lit, uint64_t nlit, const uint8_t *run, uint64_t nrun,
                 const uint8_t *syms, int nsyms, uint8_t *out, uint64_t *olen)
{
    uint8_t flag[256] = {0};
    const uint8_t *rend = run + nrun;
    uint64_t o = 0, cap = *olen, i;
    for (i = 0; i < (uint64_t)nsyms; i++) flag[syms[i]] = 1;
    for (i = 0; i < nlit; i++) {
        uint8_t b = lit[i];
        if (o >= cap) return -1;
        if (flag[b]) {
            uint32_t r;
            run += vget(run, rend, &r);
            if (r) {
                if (o + r >= cap) return -1;
                memset(out + o, b, (size_t)r + 1);
                o += (uint64_t)r + 1;
                continue;
            }
        }
        out[o++] = b;
    }
    *olen = o;
    return 0;
}

static int core_decode(const uint8_t *in, uint32_t n, uint8_t *out, uint32_t U, int N, int order)
{
    return order ? dec_order1(in, n, out, U, N) : dec_order0(in, n, out, U, N);
}

/*
 * Full container decode == rans_uncompress_to_4x16 (rANS_static4x16pr.c:1586-1873).
 * out must be given; *out_size is capacity in / length out.  Returns 0 or -1.
 */
int orc_rans_nx16_decode(const uint8_t *in, uint32_t in_size, uint8_t *out, uint32_t *out_size)
{
    const uint8_t *end = in + in_size;
    uint8_t fmt, map[16] = {0};
    uint8_t *tmp = NULL, *meta_alloc = NULL, *t1, *t2, *t3;
    const uint8_t *meta = NULL;
    uint32_t osz, t1_size, u_meta = 0;
    uint64_t unpacked = 0;
    int per_byte = 0, rc = -1, N, order;

    if (!in_size) return -1;
    fmt = in[0];

    if (fmt & 0x08) {                                           /* STRIPE :1594-1673 */
        uint32_t ulen, off = 1, k, n, clen[256], ul[256], idx[256];
        uint64_t ctot = 0;
        uint8_t *planes;
        off += vget(in + off, end, &ulen);
        if (off >= in_size) return -1;
        n = in[off++];
        if (n < 1) return -1;
        if (ulen != *out_size) return -1;
        for (k = 0; k < n; k++) {
            ul[k] = ulen / n + ((ulen % n) > k);
            idx[k] = k ? idx[k - 1] + ul[k - 1] : 0;
            off += vget(in + off, end, &clen[k]);
            ctot += clen[k];
            if (off > in_size || clen[k] > in_size || clen[k] < 1) return -1;
        }
        if (off + ctot > in_size) return -1;
        in_size = (uint32_t)(off + ctot);
        planes = malloc(ulen ? ulen : 1);
        if (!planes) return -1;
        for (k = 0; k < n; k++) {
            uint32_t got = ul[k];
            if (orc_rans_nx16_decode(in + off, in_size - off, planes + idx[k], &got) || got != ul[k]) {
                free(planes);
                return -1;
            }
            off += clen[k];
        }
        for (k = 0; k < ulen; k++)                               /* unstripe, utils.h:79-138 */
            out[k] = planes[idx[k % n] + k / n];
        free(planes);
        *out_size = ulen;
        return 0;
    }

    in++; in_size--;
    N = (fmt & 0x04) ? 32 : 4;
    order = fmt & 1;
    if (!(fmt & 0x10)) {                                        /* !NOSZ */
        int s = vget(in, end, &osz);
        in += s; in_size -= s;
    } else
        osz = *out_size;
    if (*out_size < osz) return -1;
    *out_size = osz;
    t1_size = osz;

    if (fmt & 0xc0) {
        tmp = malloc(osz ? osz : 1);
        if (!tmp) return -1;
    }
    if ((fmt & 0xc0) == 0xc0) { t1 = out; t2 = tmp; t3 = out; }
    else if (fmt & 0x80)      { t1 = tmp; t2 = tmp; t3 = out; }
    else if (fmt & 0x40)      { t1 = tmp; t2 = out; t3 = out; }
    else                      { t1 = t2 = t3 = out; }

    if (fmt & 0x80) {                                           /* PACK meta :1748-1767 */
        uint32_t psz;
        int s, m = unpack_meta(in, in_size, map, &per_byte);
        if (!m) goto done;
        unpacked = osz;
        in += m; in_size -= m;
        s = vget(in, end, &psz);
        in += s; in_size -= s;
        if (psz > t1_size) goto done;
        t1_size = psz;
    }
    if (fmt & 0x40) {                                           /* RLE meta :1769-1796 */
        uint32_t rle_len, c_meta, s;
        s = vget(in, end, &u_meta);
        s += vget(in + s, end, &rle_len);
        if (rle_len > t1_size) goto done;
        if (u_meta & 1) {
            meta = in + s;
            u_meta = (u_meta / 2 > (uint32_t)(end - meta)) ? (uint32_t)(end - meta) : u_meta / 2;
            c_meta = u_meta;
        } else {
            s += vget(in + s, end, &c_meta);
            u_meta /= 2;
            meta_alloc = malloc(u_meta ? u_meta : 1);
            if (!meta_alloc) goto done;
            if (dec_order0(in + s, in_size - s, meta_alloc, u_meta, N)) goto done;
            meta = meta_alloc;
        }
        if (c_meta + s > in_size) goto done;
        in += c_meta + s; in_size -= c_meta + s;
        t1_size = rle_len;
    }

    if (in_size) {
        if (fmt & 0x20) {                                       /* CAT */
            if (t1_size > in_size || t1_size > *out_size) goto done;
            memcpy(t1, in, t1_size);
        } else if (core_decode(in, in_size, t1, t1_size, N, order))
            goto done;
    } else
        t1_size = 0;

    {
        uint64_t t2_size = t1_size, t3_size;
        if (fmt & 0x40) {
            int ns;
            uint64_t cap = *out_size;
            if (u_meta == 0) goto done;
            ns = meta[0] ? meta[0] : 256;
            if (u_meta < (uint32_t)(1 + ns)) goto done;
            if (unrle(t1, t1_size, meta + 1 + ns, u_meta - (1 + ns), meta + 1, ns, t2, &cap)) goto done;
            t2_size = cap;
        }
        t3_size = t2_size;
        if (fmt & 0x80) {
            if (per_byte == 1) unpacked = t2_size;
            if (unpack(t2, t2_size, t3, unpacked, per_byte, map)) goto done;
            t3_size = unpacked;
        }
        *out_size = (uint32_t)t3_size;
    }
    rc = 0;
done:
    free(tmp);
    free(meta_alloc);
    return rc;
}
