// Branch-free Huffman token walk shared by the two BGZF inflate kernels (included into the anonymous
// namespace of bgzf_inflate.cu).  One loop iteration decodes ONE Huffman symbol — a literal/length
// symbol or, right after a length, a distance symbol — for every lane of the warp in lock step; lanes
// that are done idle (their table entry is masked to zero, so nothing moves).  The 32 bits at the lane's
// bit position come straight from the stream; long codes (second-level table) take a warp-uniform
// branch.  inflate_fast of zlib / the loops of libdeflate behind bgzf.c:730-804 are what this replaces.
constexpr uint32_t REC_STRIDE = 256u * 4u;   // token records of the CTA kernel: word j of thread t at rec[j * 256 + t]

__device__ __forceinline__ uint32_t sld8(uint32_t a) { uint32_t v; asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ uint32_t sld32(uint32_t a) { uint32_t v; asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v; }
__device__ __forceinline__ void sst8(uint32_t a, uint32_t v) { asm volatile("st.shared.u8 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sst32(uint32_t a, uint32_t v) { asm volatile("st.shared.u32 [%0], %1;" :: "r"(a), "r"(v) : "memory"); }
// predicated forms: straight-line code (the compiler turns `if (c) asm(...)` into divergent branches)
__device__ __forceinline__ uint32_t sld8_if(uint32_t a, uint32_t c)
{ uint32_t v; asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.shared.u8 %0, [%1];\n\t}" : "=r"(v) : "r"(a), "r"(c) : "memory"); return v; }
__device__ __forceinline__ uint32_t sld32_if(uint32_t a, uint32_t c)
{ uint32_t v; asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\tmov.u32 %0, 0;\n\t@q ld.shared.u32 %0, [%1];\n\t}" : "=r"(v) : "r"(a), "r"(c) : "memory"); return v; }
__device__ __forceinline__ void sst8_if(uint32_t a, uint32_t v, uint32_t c)
{ asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u8 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"(c) : "memory"); }
__device__ __forceinline__ void sst32_if(uint32_t a, uint32_t v, uint32_t c)
{ asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u32 [%0], %1;\n\t}" :: "r"(a), "r"(v), "r"(c) : "memory"); }


#ifdef HGPU_PROFILE
__device__ uint32_t g_dbg[6 * 256];
__device__ int g_dbg_job = -1;
#endif
constexpr uint32_t PREROLL = 128;            // bits walked in front of a cut before counting starts
constexpr uint32_t REC_MAX = 48;             // recorded tokens per thread: word j of thread t at rec[j * 256 + t]
                                             // (48 tokens: bytes in front < 48 x 258 < 2^14, matches in front < 2^6)

// the 32 bits at bit position `pos` (wbase coordinates)
template <bool SMEM>
__device__ __forceinline__ uint32_t stream_bits(uint32_t sbase, const uint32_t *wbase, const uint32_t *wend, uint32_t pos)
{
    uint32_t w0, w1;
    if (SMEM) {
        const uint32_t a = sbase + ((pos >> 5) << 2);
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w0) : "r"(a));
        asm volatile("ld.shared.u32 %0, [%1+4];" : "=r"(w1) : "r"(a));
    } else {
        const uint32_t *p = wbase + (pos >> 5);
        w0 = p < wend ? *p : 0u;
        w1 = p + 1 < wend ? p[1] : 0u;
    }
    return __funnelshift_r(w0, w1, pos & 31u);
}

// One Huffman symbol for every lane: table entry `e` (sub-table links resolved; 0 for lanes whose
// mask `actm` is 0, so that nothing below moves for them), bits consumed by the code itself in nb.
// raw = the entry before masking (to tell an invalid code from an idle lane).
__device__ __forceinline__ uint32_t huff_entry(uint32_t lita, uint32_t dsta, uint32_t isd, uint32_t w, uint32_t actm,
                                               uint32_t &nb, uint32_t &raw)
{
    const uint32_t ta = isd ? dsta : lita, msk = isd ? (1u << DST_ROOT) - 1u : (1u << LIT_ROOT) - 1u;
    uint32_t e = sld32(ta + ((w & msk) << 2));
    uint32_t rootc = 0;
    const bool sub = (e & 0x30u) == 0x30u;
    if (__any_sync(0xffffffffu, sub && actm)) {                  // a code longer than the root index: rare
        if (sub) {
            const uint32_t root = isd ? DST_ROOT : LIT_ROOT;
            e = sld32(ta + (((e >> 16) + ((w >> root) & ~(0xffffffffu << ((e >> 8) & 0xffu)))) << 2));
            rootc = root;
            if ((e & 0x30u) == 0x30u) e = 0;                     // never: a link behind a link
        }
    }
    raw = e;
    e &= actm;
    nb = (e & 15u) + (rootc & actm);
    return e;
}

// MODE 0: count.  MODE 1: count + record token starts.  MODE 2: emit into the shared window (CTA kernel).
// MODE 3: emit into global memory + {dst, len | dist << 16} match records (warp kernel).
// Walks [start, end) for the lanes with act set; all 32 lanes of the warp must call it together.
template <int MODE, bool SMEM, bool DGLOBAL>
__device__ __forceinline__ void huff_walk(const InflateSmem &s, uint32_t sbase, const uint32_t *wbase, const uint32_t *wend,
                                          uint32_t cut, uint32_t start, uint32_t end, bool act, uint32_t reca,
                                          uint32_t &exitp, uint32_t &nout, uint32_t &nmatch, uint32_t &st, uint32_t &rcnt,
                                          uint32_t wa, uint32_t obase, uint32_t mbase, uint16_t *dlist, bool &bad_dist,
                                          uint32_t from = 0, uint32_t *first = nullptr, uint8_t *gout = nullptr, uint2 *mrec = nullptr)
{
    // `from` (MODE 0/1): tokens that start in front of it are walked but not counted — a pre-roll, so that
    // the walk has usually found the true token grid by the time it reaches the thread's own range;
    // *first = where the first counted token starts.
    const uint32_t lita = (uint32_t)__cvta_generic_to_shared(s.lit), dsta = (uint32_t)__cvta_generic_to_shared(s.dst);
    const uint32_t da = DGLOBAL ? 0u : (uint32_t)__cvta_generic_to_shared(dlist) + 2u * mbase;
    const bool live = act && start < end;
    // idle lanes run the loop too (their loads are real): park them on the first word of the stream
    uint32_t pos = live ? start : 0u, n = 0, m = 0, lenp = 0, status = ST_RUN, j = 0, isd = 0, recp = reca, p0 = 0xffffffffu;
    uint32_t actm = live ? 0xffffffffu : 0u;
    uint32_t onm = (MODE >= 2 || start >= from) ? 0xffffffffu : 0u;          // counting yet?
    const bool recok = from >= cut;        // (a predecessor that stopped at an invalid code hands over a start in front of the cut)
    while (__any_sync(0xffffffffu, actm != 0u)) {
        const uint32_t w = stream_bits<SMEM>(sbase, wbase, wend, pos);
        if (MODE < 2) {
            if (!isd && pos >= from) onm = 0xffffffffu;
            if (onm && actm && p0 == 0xffffffffu) p0 = pos;
        }
        if (MODE == 1) {
            // a token starts here: relative position (12 bits) | bytes in front (14 bits) | matches in front (6 bits)
            const uint32_t r = (actm && onm && !isd && recok && j < REC_MAX) ? 1u : 0u;
            sst32_if(recp, (pos - cut) | n << 12 | m << 26, r);
            recp += r * (REC_STRIDE);
            j += r;
        }
        uint32_t nb, raw;
        uint32_t e = huff_entry(lita, dsta, isd, w, actm, nb, raw);
        const uint32_t xb = (e >> 8) & 0xffu;
        const uint32_t val = (e >> 16) + ((w >> nb) & ~(0xffffffffu << xb));
        if (MODE == 2) {
            const uint32_t o = obase + n;
            sst8_if(wa + o, val, e & 0x10u);                     // literal
            const uint32_t dm = e & 0x40u;                       // distance: the match is complete
            if (dm && val > o) bad_dist = true;
            sst8_if(wa + o, lenp - 3u, dm);
            sst8_if(wa + o + 1, val - 1u, dm);
            sst8_if(wa + o + 2, (val - 1u) >> 8, dm);
            if (DGLOBAL) { if (dm) dlist[mbase + m] = (uint16_t)o; }
            else asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %2, 0;\n\t@q st.shared.u16 [%0], %1;\n\t}"
                              :: "r"(da + 2u * m), "h"((uint16_t)o), "r"(dm) : "memory");
        }
        if (MODE == 3) {
            const uint32_t o = obase + n;
            gst8_if(gout + o, val, e & 0x10u);                   // literal
            if (e & 0x40u) {                                     // distance: the match is complete
                if (val > o) bad_dist = true;
                mrec[mbase + m] = make_uint2(o, lenp | (val << 16));
            }
        }
        pos += nb + xb;
        const uint32_t ec = MODE >= 2 ? e : e & onm;             // what counts
        n += ((ec >> 4) & 1u) + ((ec & 0x40u) ? lenp : 0u);
        m += (ec >> 6) & 1u;
        lenp = (e & 0x20u) ? val : lenp;
        isd = (e >> 5) & 1u;
        const bool eob = (e & 0x80u) != 0u, bad = actm && (raw & 0xf0u) == 0u;
        status = eob ? ST_EOB : bad ? ST_BAD : status;
        actm = (actm && !eob && !bad && (isd || pos < end)) ? 0xffffffffu : 0u;
    }
    exitp = live ? pos : start; nout = n; nmatch = m; st = status;
    if (MODE == 1) rcnt = j;
    if (MODE < 2 && first) *first = p0;       // 0xffffffff: the walk ended (a false end-of-block / invalid code) in front of `from`
}

// The thread's start moved to `start`: walk from there until a recorded token start is hit.
// (n0, m0) are the totals of the recorded walk.  landed = false: the walk never hit one.
// All 32 lanes call together; act selects the lanes that have something to do.
template <bool SMEM>
__device__ __forceinline__ void huff_fixup(const InflateSmem &s, uint32_t sbase, const uint32_t *wbase, const uint32_t *wend,
                                           uint32_t cut, uint32_t start, uint32_t end, bool act, uint32_t reca, uint32_t rcnt,
                                           uint32_t n0, uint32_t m0, uint32_t &nout, uint32_t &nmatch, bool &landed)
{
    const uint32_t lita = (uint32_t)__cvta_generic_to_shared(s.lit), dsta = (uint32_t)__cvta_generic_to_shared(s.dst);
    act = act && rcnt > 0 && start < end && start >= cut;
    uint32_t pos = act ? start : 0u, n = 0, m = 0, lenp = 0, jj = 0, isd = 0;        // idle lanes: parked on the first word
    bool hit = false;
    uint32_t actm = act ? 0xffffffffu : 0u;
    uint32_t rw = sld32_if(reca, actm);                          // recorded word jj
    if (!act) rw = 0xffffffffu;
    while (__any_sync(0xffffffffu, actm != 0u)) {
        // at a token start: advance the record pointer to the first recorded start >= pos (the two walks
        // step differently), then see whether this is one.  Every lane runs the loop; chk selects.
        const bool chk = actm && !isd;
        const uint32_t rel = pos - cut;
        for (;;) {
            const bool adv = chk && (rw & 0xfffu) < rel;         // the sentinel 0xffffffff never advances (rel < 4096)
            if (!__any_sync(0xffffffffu, adv)) break;
            jj += adv;
            const bool in = adv && jj < rcnt;
            const uint32_t v = sld32_if(reca + jj * (REC_STRIDE), in);
            rw = in ? v : adv ? 0xffffffffu : rw;
        }
        const bool land = chk && rw != 0xffffffffu && (rw & 0xfffu) == rel;
        if (land) { hit = true; nout = n + n0 - ((rw >> 12) & 0x3fffu); nmatch = m + m0 - (rw >> 26); }
        if (chk && (land || rw == 0xffffffffu || pos >= end)) actm = 0u;
        const uint32_t w = stream_bits<SMEM>(sbase, wbase, wend, pos);
        uint32_t nb, raw;
        const uint32_t e = huff_entry(lita, dsta, isd, w, actm, nb, raw);
        const uint32_t xb = (e >> 8) & 0xffu;
        const uint32_t val = (e >> 16) + ((w >> nb) & ~(0xffffffffu << xb));
        pos += nb + xb;
        n += ((e >> 4) & 1u) + ((e & 0x40u) ? lenp : 0u);
        m += (e >> 6) & 1u;
        lenp = (e & 0x20u) ? val : lenp;
        isd = (e >> 5) & 1u;
        if ((e & 0x70u) == 0u) actm = 0u;                        // end-of-block or an invalid code on the way: take the full walk
    }
    landed = hit;
}

