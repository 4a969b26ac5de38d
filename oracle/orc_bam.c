/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see orc_rans_nx16.c header for the usage rule).
 *
 * CPU restatement of BAM record unpack in htslib 1.23.1:
 *   bam_read1                 sam.c:784-860   (core decode, qname NUL padding, bounds checks,
 *                                              bin recompute, CIGAR/qlen consistency)
 *   fixup_missing_qname_nul   sam.c:763-778
 *   bam_cigar2rqlens          sam.c:508-519 ; bam_cigar_type htslib/sam.h:152 (0x3C1A7)
 *   hts_reg2bin               htslib/hts.h:1516-1522
 *   nibble2base / seq_nt16_str sam_internal.h:63-118 ; hts.c:260
 *   add33                     sam.c:4317-4322 (QUAL + 33 as sam_format1_append does, :4370-4376)
 * Parity status: PINNED — tests/test_oracle_bam.py runs the compiled reference's own bam_read1
 * (oracle/_ref) over the reference's BAM fixtures and seeded synthetic BAM and compares every
 * bam1_core_t field and every data byte.
 *
 * Not restated: the CG-tag long-CIGAR rewrite (bam_tag2cigar, sam.c:680-760).  Records that meet
 * its cheap trigger (first CIGAR op == <l_qseq>S, tid >= 0, pos >= 0) are flagged status 1 so
 * the caller can hand them to the host routine; they never occur in 150 bp short-read data.
 */
#include <stdint.h>
#include <string.h>

typedef struct {
    int64_t  pos;
    int32_t  tid;
    uint16_t bin;
    uint8_t  qual;
    uint8_t  l_extranul;
    uint16_t flag;
    uint16_t l_qname;
    uint32_t n_cigar;
    int32_t  l_qseq;
    int32_t  mtid;
    int64_t  mpos;
    int64_t  isize;
} orc_bam1_core;

static uint32_t le32(const uint8_t *p) { return p[0] | p[1] << 8 | p[2] << 16 | (uint32_t)p[3] << 24; }

static int reg2bin(int64_t beg, int64_t end)
{
    int l, s = 14, t = ((1 << 15) - 1) / 7;
    for (--end, l = 5; l > 0; --l, s += 3, t -= 1 << ((l << 1) + l))
        if (beg >> s == end >> s) return t + (int)(beg >> s);
    return 0;
}

/* Walk the block_size chain.  Returns the record count, or -1-k if record k is malformed
 * (block_size < 32 or running past the end: bam_read1 returns -4 / -3 / -2 there). */
long orc_bam_index(const uint8_t *st, uint64_t len, uint64_t *rec_off, long cap)
{
    uint64_t p = 0;
    long n = 0;
    while (p < len) {
        int32_t bl;
        if (len - p < 4) return -1 - n;
        bl = (int32_t)le32(st + p);
        if (bl < 32 || p + 4 + (uint64_t)bl > len) return -1 - n;
        if (n < cap) rec_off[n] = p;
        n++;
        p += 4 + (uint64_t)bl;
    }
    return n;
}

/* l_data / l_qseq a record will have after bam_read1 (0,0 when the record is invalid). */
int orc_bam_sizes(const uint8_t *rec, uint32_t *l_data, uint32_t *l_qseq)
{
    int32_t bl = (int32_t)le32(rec);
    uint32_t l_qname = rec[4 + 8], n_cigar = le32(rec + 4 + 12) & 0xffff;
    int32_t lq = (int32_t)le32(rec + 4 + 16);
    uint32_t xn = (l_qname % 4) ? 4 - l_qname % 4 : 0;
    uint64_t nl = (uint64_t)(bl - 32) + xn;
    *l_data = *l_qseq = 0;
    if (bl < 32 || nl > 0x7fffffff || lq < 0 || l_qname < 1) return -4;
    if (((uint64_t)n_cigar << 2) + l_qname + xn + (((uint64_t)lq + 1) >> 1) + (uint64_t)lq > nl) return -4;
    if (rec[4 + 32 + l_qname - 1] != 0 && xn == 0) nl += 4;            /* fixup_missing_qname_nul */
    *l_data = (uint32_t)nl;
    *l_qseq = (uint32_t)lq;
    return 0;
}

/*
 * One record -> core, data (bam1_t::data), seq (ASCII), qual (+33).  Any of data/seq/qual may be
 * NULL.  Returns 0, 1 (possible CG long-CIGAR, left untouched) or -4 (bam_read1's error).
 */
int orc_bam_unpack1(const uint8_t *rec, orc_bam1_core *c, uint8_t *data, uint8_t *seq, uint8_t *qual)
{
    static const char nt16[] = "=ACMGRSVTWYHKDBN";
    const uint8_t *x = rec + 4, *body = rec + 36;
    int32_t bl = (int32_t)le32(rec);
    uint32_t x2, x3, l_data, lq, i, qn, rest;
    (void)l_data;
    int status = 0;
    if (orc_bam_sizes(rec, &l_data, &lq)) { memset(c, 0, sizeof(*c)); return -4; }
    c->tid = (int32_t)le32(x);
    c->pos = (int32_t)le32(x + 4);
    x2 = le32(x + 8);
    c->bin = (uint16_t)(x2 >> 16);
    c->qual = (x2 >> 8) & 0xff;
    c->l_qname = x2 & 0xff;
    c->l_extranul = (c->l_qname % 4) ? 4 - c->l_qname % 4 : 0;
    x3 = le32(x + 12);
    c->flag = (uint16_t)(x3 >> 16);
    c->n_cigar = x3 & 0xffff;
    c->l_qseq = (int32_t)lq;
    c->mtid = (int32_t)le32(x + 20);
    c->mpos = (int32_t)le32(x + 24);
    c->isize = (int32_t)le32(x + 28);
    qn = c->l_qname;                                     /* on-disk qname length */
    rest = (uint32_t)(bl - 32) - qn;                     /* bytes after the on-disk qname */
    if (data) memcpy(data, body, qn);
    if (body[qn - 1] != 0) {                             /* fixup_missing_qname_nul */
        if (c->l_extranul > 0) c->l_extranul--;
        else c->l_extranul = 3;
        if (data) data[qn] = 0;
        c->l_qname++;
    }
    if (data) {
        memset(data + c->l_qname, 0, c->l_extranul);
        memcpy(data + c->l_qname + c->l_extranul, body + qn, rest);
    }
    c->l_qname += c->l_extranul;
    {
        const uint8_t *cig = body + (x2 & 0xff);
        const uint8_t *sq = cig + 4 * (size_t)c->n_cigar;
        const uint8_t *ql = sq + (lq + 1) / 2;
        if (c->n_cigar > 0) {
            int64_t rlen = 0, qlen = 0;
            uint32_t k, first = le32(cig);
            if (first == (4u | (lq << 4)) && c->tid >= 0 && c->pos >= 0) status = 1;
            for (k = 0; k < c->n_cigar; k++) {
                uint32_t op = le32(cig + 4 * k);
                int type = 0x3C1A7 >> ((op & 0xf) << 1) & 3;
                if (type & 1) qlen += op >> 4;
                if (type & 2) rlen += op >> 4;
            }
            if ((c->flag & 4) || rlen == 0) rlen = 1;
            c->bin = (uint16_t)reg2bin(c->pos, c->pos + rlen);
            if (lq > 0 && !(c->flag & 4) && qlen != (int64_t)lq) return -4;
        }
        if (seq)
            for (i = 0; i < lq; i++) seq[i] = (uint8_t)nt16[(sq[i >> 1] >> ((~i & 1) << 2)) & 0xf];
        if (qual) {
            if (lq && ql[0] == 0xff) memcpy(qual, ql, lq);
            else for (i = 0; i < lq; i++) qual[i] = (uint8_t)(ql[i] + 33);
        }
    }
    return status;
}

/*
 * bam_write1 (sam.c:862-928) for one record: core + data -> BAM bytes (block_size included).
 * Returns bytes written; -1 for the reference's error conditions (qname > 254 chars, positions
 * beyond INT_MAX); -2 when n_cigar > 0xffff (the CG-tag rewrite, :899-925, is left to the host).
 */
long orc_bam_pack1(const orc_bam1_core *c, const uint8_t *data, uint32_t l_data, uint8_t *out)
{
    uint32_t x[8], block_len = l_data - c->l_extranul + 32, qn = c->l_qname - c->l_extranul;
    int i;
    if (qn > 255) return -1;
    if (c->n_cigar > 0xffff) return -2;
    if (c->pos > 0x7fffffffLL || c->mpos > 0x7fffffffLL || c->isize < -0x80000000LL || c->isize > 0x7fffffffLL) return -1;
    x[0] = (uint32_t)c->tid; x[1] = (uint32_t)c->pos;
    x[2] = (uint32_t)c->bin << 16 | (uint32_t)c->qual << 8 | qn;
    x[3] = (uint32_t)c->flag << 16 | (c->n_cigar & 0xffff);
    x[4] = (uint32_t)c->l_qseq; x[5] = (uint32_t)c->mtid; x[6] = (uint32_t)c->mpos; x[7] = (uint32_t)c->isize;
    out[0] = block_len; out[1] = block_len >> 8; out[2] = block_len >> 16; out[3] = block_len >> 24;
    for (i = 0; i < 8; i++) { out[4 + 4 * i] = x[i]; out[5 + 4 * i] = x[i] >> 8; out[6 + 4 * i] = x[i] >> 16; out[7 + 4 * i] = x[i] >> 24; }
    memcpy(out + 36, data, qn);
    memcpy(out + 36 + qn, data + c->l_qname, l_data - c->l_qname);
    return 4 + (long)block_len;
}
