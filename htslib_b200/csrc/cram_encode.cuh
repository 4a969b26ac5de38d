// CRAM 3.x record ENCODE: bam1_t records -> the data series of a slice (the record loop of cram_encode_slice /
// process_one_read, cram/cram_encode.c:2050-2420, :3490-4010, in the shape the reference writes with no_ref: every base and
// quality explicit, CIGAR as read features, mates detached).
//
// Encoding is parallel where decoding is serial: a record's bytes in every series depend on that record alone, only
// their POSITION depends on the records before it.  So: one thread per record counts what it contributes to each of the
// 30 byte streams (walk<false>), a scan per (slice, stream) turns counts into offsets, one thread per record writes
// (walk<true>).  Same source for both passes, __host__ __device__ (tests/hostsim builds it for the host).
//
// Series layout (all EXTERNAL, ITF8 integers; content id = stream index + 1):
//   BF CF RI RL AP RG | RN (BYTE_ARRAY_STOP 0) | MF NS NP TS | TL | FN, per feature FC FP and DL / RS / HC / PD or
//   BB / SC / IN (BYTE_ARRAY_LEN: length stream + value stream) | BA (unmapped reads) | QS | MQ |
//   tags: every tag value BYTE_ARRAY_LEN over two shared streams (lengths, values), in tag-line order.
// Every record is written detached (MF NS NP TS explicit, no mate cross references to resolve), RG stays an ordinary
// aux tag (RG series = -1), AP is absolute, RI is per record (multi-reference slice, ref_seq_id -2) and RR = 0 (no
// reference needed to decode).  What the reference's decoder returns for such a slice is the input record, except what
// CRAM cannot hold: '=' / 'X' CIGAR ops come back as 'M', the mapping quality of an unmapped read as 0.
#pragma once
#include <stdint.h>
#include <stddef.h>

#ifndef CRAMREC_HD
#ifdef __CUDACC__
#define CRAMREC_HD __host__ __device__
#else
#define CRAMREC_HD
#endif
#endif

namespace cramenc {

enum Stream { S_BF, S_CF, S_RI, S_RL, S_AP, S_RG, S_RN, S_MF, S_NS, S_NP, S_TS, S_TL, S_FN, S_FC, S_FP, S_DL, S_RS, S_HC, S_PD,
              S_BB_LEN, S_BB, S_SC_LEN, S_SC, S_IN_LEN, S_IN, S_BA, S_QS, S_MQ, S_TAG_LEN, S_TAG_VAL, S_BS, S_COUNT };

struct Core { int64_t pos; int32_t tid; uint16_t bin; uint8_t qual, l_extranul; uint16_t flag, l_qname; uint32_t n_cigar; int32_t l_qseq, mtid; int64_t mpos, isize; };

enum { ENC_OK = 0, ENC_UNSUPPORTED = -6, ENC_BAD = -1 };

CRAMREC_HD inline int itf8_size(uint32_t v) { return v < 0x80 ? 1 : v < 0x4000 ? 2 : v < 0x200000 ? 3 : v < 0x10000000 ? 4 : 5; }
CRAMREC_HD inline int itf8_put(uint8_t *p, uint32_t v)                    // itf8_put, cram/cram_io.h
{
    if (v < 0x80) { p[0] = (uint8_t)v; return 1; }
    if (v < 0x4000) { p[0] = (uint8_t)((v >> 8) | 0x80); p[1] = (uint8_t)v; return 2; }
    if (v < 0x200000) { p[0] = (uint8_t)((v >> 16) | 0xc0); p[1] = (uint8_t)(v >> 8); p[2] = (uint8_t)v; return 3; }
    if (v < 0x10000000) { p[0] = (uint8_t)((v >> 24) | 0xe0); p[1] = (uint8_t)(v >> 16); p[2] = (uint8_t)(v >> 8); p[3] = (uint8_t)v; return 4; }
    p[0] = (uint8_t)(0xf0 | ((v >> 28) & 0xff)); p[1] = (uint8_t)(v >> 20); p[2] = (uint8_t)(v >> 12); p[3] = (uint8_t)(v >> 4); p[4] = (uint8_t)(v & 0x0f);
    return 5;
}

// one aux field at p: tag[2], type, value length (bytes after the 3-byte id).  Returns false on a malformed field.
CRAMREC_HD inline bool aux_field(const uint8_t *p, const uint8_t *end, uint32_t &vlen)
{
    if (end - p < 3) return false;
    const uint8_t t = p[2];
    const uint8_t *v = p + 3;
    switch (t) {
    case 'A': case 'c': case 'C': vlen = 1; break;
    case 's': case 'S': vlen = 2; break;
    case 'i': case 'I': case 'f': vlen = 4; break;
    case 'd': vlen = 8; break;
    case 'Z': case 'H': { const uint8_t *q = v; while (q < end && *q) q++; if (q >= end) return false; vlen = (uint32_t)(q - v) + 1; break; }
    case 'B': {
        if (end - v < 5) return false;
        const uint8_t st = v[0];
        const uint32_t n = v[1] | v[2] << 8 | v[3] << 16 | (uint32_t)v[4] << 24;
        const uint32_t es = (st == 'c' || st == 'C') ? 1 : (st == 's' || st == 'S') ? 2 : (st == 'i' || st == 'I' || st == 'f') ? 4 : 0;
        if (!es || n > 0x10000000u) return false;
        vlen = 5 + n * es;
        break; }
    default: return false;
    }
    return (uint64_t)(end - v) >= vlen;
}

// WRITE = false: cnt[s] += bytes this record adds to stream s.  WRITE = true: the bytes go to base[s] + off[s] (off advances).
template <bool WRITE>
struct Emit {
    uint32_t *n;                     // counts or running offsets, S_COUNT entries
    uint8_t *const *base;
    CRAMREC_HD void put_int(int s, int32_t v)
    {
        if (WRITE) n[s] += (uint32_t)itf8_put(base[s] + n[s], (uint32_t)v);
        else n[s] += (uint32_t)itf8_size((uint32_t)v);
    }
    CRAMREC_HD void put_byte(int s, uint8_t b) { if (WRITE) base[s][n[s]] = b; n[s] += 1; }
    CRAMREC_HD void put_bytes(int s, const uint8_t *p, uint32_t len) { if (WRITE) for (uint32_t i = 0; i < len; i++) base[s][n[s] + i] = p[i]; n[s] += len; }
    CRAMREC_HD void put_fill(int s, uint8_t b, uint32_t len) { if (WRITE) for (uint32_t i = 0; i < len; i++) base[s][n[s] + i] = b; n[s] += len; }
    CRAMREC_HD void put_bases(int s, const uint8_t *seq4, uint32_t from, uint32_t len)          // 4-bit SEQ -> ASCII
    {
        if (WRITE) for (uint32_t i = 0; i < len; i++) { const uint32_t q = from + i; base[s][n[s] + i] = (uint8_t)"=ACMGRSVTWYHKDBN"[(seq4[q >> 1] >> ((~q & 1) << 2)) & 15]; }
        n[s] += len;
    }
};

CRAMREC_HD inline uint8_t base_at(const uint8_t *seq4, uint32_t q) { return (uint8_t)"=ACMGRSVTWYHKDBN"[(seq4[q >> 1] >> ((~q & 1) << 2)) & 15]; }
CRAMREC_HD inline int l1_row(uint8_t c) { return c == 'A' ? 0 : c == 'C' ? 1 : c == 'G' ? 2 : c == 'T' ? 3 : 4; }
// substitution code of read base sb against reference base rb under the default matrix (rows CGTN AGTN ACTN ACGN ACGT): -1 = not expressible
CRAMREC_HD inline int bs_code(uint8_t rb, uint8_t sb)
{
    const char *row = l1_row(rb) == 0 ? "CGTN" : l1_row(rb) == 1 ? "AGTN" : l1_row(rb) == 2 ? "ACTN" : l1_row(rb) == 3 ? "ACGN" : "ACGT";
    for (int k = 0; k < 4; k++) if ((uint8_t)row[k] == sb) return k;
    return -1;
}

// CIGAR (+ SEQ, + reference) -> read features.  EMIT = false only counts them (FN is written before the features).
// ref != nullptr: match operations are compared with the reference (ref[0] = base 1 of the record's reference sequence):
// equal bases leave no feature, a different A/C/G/T/N is a substitution 'X' (BS = its code in the reference base's row),
// anything else a one-base 'b'.  ref == nullptr: every match operation is a 'b' run with its bases.
template <bool WRITE, bool EMIT>
CRAMREC_HD inline int features(const Core &c, const uint8_t *cig, uint32_t nc, const uint8_t *seq4, int32_t ls, bool noseq,
                               const uint8_t *ref, int64_t ref_len, Emit<WRITE> &E, uint32_t &nf)
{
    uint32_t spos = 1, prev = 0, qlen = 0;
    int64_t rpos = c.pos;
    nf = 0;
    for (uint32_t k = 0; k < nc; k++) {
        const uint32_t w = cig[4 * k] | cig[4 * k + 1] << 8 | cig[4 * k + 2] << 16 | (uint32_t)cig[4 * k + 3] << 24;
        const uint32_t op = w & 15, len = w >> 4;
        if (len == 0) return ENC_UNSUPPORTED;                                  // zero-length ops do not survive the feature form
        uint8_t code;
        switch (op) {
        case 0: case 7: case 8: code = 'b'; break;
        case 1: code = 'I'; break;
        case 2: code = 'D'; break;
        case 3: code = 'N'; break;
        case 4: code = 'S'; break;
        case 5: code = 'H'; break;
        default: code = 'P'; break;
        }
        if (code == 'b' && noseq) { spos += len; qlen += len; rpos += len; continue; }       // implicit match
        if (code == 'b' && ref) {
            if ((uint64_t)spos - 1 + len > (uint64_t)ls) return ENC_BAD;
            for (uint32_t i = 0; i < len; i++) {
                const uint8_t rb = rpos + i < ref_len ? ref[rpos + i] : (uint8_t)'N', sb = base_at(seq4, spos - 1 + i);
                if (rb == sb) continue;
                const int bs = bs_code(rb, sb);
                nf++;
                if (EMIT) {
                    E.put_byte(S_FC, bs >= 0 ? 'X' : 'b');
                    E.put_int(S_FP, (int32_t)(spos + i - prev));
                    prev = spos + i;
                    if (bs >= 0) E.put_byte(S_BS, (uint8_t)bs);
                    else { E.put_int(S_BB_LEN, 1); E.put_byte(S_BB, sb); }
                }
            }
            spos += len; qlen += len; rpos += len;
            continue;
        }
        nf++;
        if (EMIT) { E.put_byte(S_FC, code); E.put_int(S_FP, (int32_t)(spos - prev)); prev = spos; }
        if (code == 'b' || code == 'I' || code == 'S') {
            if (!noseq && (uint64_t)spos - 1 + len > (uint64_t)ls) return ENC_BAD;              // CIGAR longer than SEQ
            const int ln = code == 'b' ? S_BB_LEN : code == 'I' ? S_IN_LEN : S_SC_LEN;
            if (EMIT) { E.put_int(ln, (int32_t)len); if (noseq) E.put_fill(ln + 1, 'N', len); else E.put_bases(ln + 1, seq4, spos - 1, len); }
            spos += len; qlen += len;
            if (code == 'b') rpos += len;
        } else {
            if (EMIT) E.put_int(code == 'D' ? S_DL : code == 'N' ? S_RS : code == 'H' ? S_HC : S_PD, (int32_t)len);
            if (code == 'D' || code == 'N') rpos += len;
        }
    }
    if (!noseq && qlen != (uint32_t)ls) return ENC_BAD;                        // bam_set1 would refuse what the decoder rebuilds
    return ENC_OK;
}

// One record.  tl = its tag-line index (the host built the dictionary).  ref / ref_len: the record's reference sequence
// (nullptr: none).  Returns ENC_OK or why the slice cannot be written here.
template <bool WRITE>
CRAMREC_HD inline int walk(const Core &c, const uint8_t *data, uint32_t l_data, int32_t tl, const uint8_t *ref, int64_t ref_len, Emit<WRITE> &E)
{
    const uint32_t lq = c.l_qname, nc = c.n_cigar;
    const int32_t ls = c.l_qseq;
    if (lq == 0 || ls < 0 || (uint64_t)lq + 4ull * nc + ((uint64_t)ls + 1) / 2 + (uint64_t)ls > l_data) return ENC_BAD;
    const uint8_t *cig = data + lq, *seq4 = cig + 4 * nc, *qual = seq4 + (ls + 1) / 2, *aux = qual + ls, *end = data + l_data;
    const bool unmapped = (c.flag & 4) != 0;
    const bool has_qual = ls > 0 && qual[0] != 0xff;
    if (!unmapped && c.pos < 0) return ENC_UNSUPPORTED;                       // the reference's decoder refuses a mapped read at position 0
    if (c.flag >= 0x1000) return ENC_BAD;
    // a mapped read stored without its sequence ("*"): CRAM_FLAG_NO_SEQ, the read length comes from the CIGAR, match
    // operations are implicit (no feature), inserted and clipped bases are placeholders (process_one_read :3845-3870)
    const bool noseq = !unmapped && ls == 0;
    uint32_t cig_q = 0;
    for (uint32_t k = 0; k < nc; k++) {
        const uint32_t w = cig[4 * k] | cig[4 * k + 1] << 8 | cig[4 * k + 2] << 16 | (uint32_t)cig[4 * k + 3] << 24;
        const uint32_t op = w & 15;
        if (op > 8) return ENC_BAD;
        if (op == 0 || op == 1 || op == 4 || op == 7 || op == 8) cig_q += w >> 4;
    }
    if (noseq && nc == 0) return ENC_UNSUPPORTED;
    E.put_int(S_BF, c.flag);
    E.put_int(S_CF, 2 | (has_qual ? 1 : 0) | (noseq ? 8 : 0));               // DETACHED | PRESERVE_QUAL_SCORES | NO_SEQ
    E.put_int(S_RI, c.tid);
    E.put_int(S_RL, noseq ? (int32_t)cig_q : ls);
    E.put_int(S_AP, (int32_t)(c.pos + 1));
    E.put_int(S_RG, -1);
    {   // the name up to its first NUL, then the stop byte
        uint32_t nl = 0;
        while (nl < lq && data[nl]) nl++;
        E.put_bytes(S_RN, data, nl);
        E.put_byte(S_RN, 0);
    }
    E.put_int(S_MF, 0);
    E.put_int(S_NS, c.mtid);
    E.put_int(S_NP, (int32_t)(c.mpos + 1));
    E.put_int(S_TS, (int32_t)c.isize);
    E.put_int(S_TL, tl);
    // tags: lengths + values, in order
    for (const uint8_t *p = aux; p < end;) {
        uint32_t vlen = 0;
        if (!aux_field(p, end, vlen)) return ENC_BAD;
        E.put_int(S_TAG_LEN, (int32_t)vlen);
        E.put_bytes(S_TAG_VAL, p + 3, vlen);
        p += 3 + vlen;
    }
    if (!unmapped) {
        uint32_t nf = 0;
        int rc = features<WRITE, false>(c, cig, nc, seq4, ls, noseq, ref, ref_len, E, nf);
        if (rc != ENC_OK) return rc;
        E.put_int(S_FN, (int32_t)nf);
        rc = features<WRITE, true>(c, cig, nc, seq4, ls, noseq, ref, ref_len, E, nf);
        if (rc != ENC_OK) return rc;
        E.put_int(S_MQ, c.qual);
    } else E.put_bases(S_BA, seq4, 0, (uint32_t)ls);
    if (has_qual) E.put_bytes(S_QS, qual, (uint32_t)ls);
    return ENC_OK;
}

}  // namespace cramenc
