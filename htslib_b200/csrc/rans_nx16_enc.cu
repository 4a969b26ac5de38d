// rANS Nx16 ("RANS_PR", CRAM 3.1 method 5) ENCODER for sm_100a — order 0 / order 1, 4-way or
// 32-way (X32), with the CAT fallback; one warp per stream.
//
// Stands where rans_compress_to_4x16 -> rans_compress_O0/O1_4x16 / _32x16 stand in the reference
// (htscodecs rANS_static4x16pr.c:1203-1579, :112-211, :402-497; rANS_static32x16pr.c:67-252,
// :412-525; table writer rANS_static16_int.h:165-421).  Compressed bytes need not equal the
// reference's (encoder byte parity is unpinned) — what is pinned is that the
// reference's decoder, the oracle and our own decoder all return the input (tests/test_gpu_rans_enc.py).
//
// The N states live one per lane, mirroring the decoder: symbols are encoded backwards, the lanes
// that must renormalise in a step are found with a ballot, and their 16-bit words are laid down
// below the previous ones in lane order — the exact reverse of the decoder's
// `base + popc(ballot & lanemask_lt)` read order.  Frequencies: warp-aggregated shared/global
// histograms (__match_any_sync), normalised to 2^12 (order 0) or 2^10 / 2^12 per context
// (order 1) so that every present symbol keeps f >= 1.
#include "hgpu_internal.h"

namespace {

constexpr uint32_t RANS_L = 1u << 15;
constexpr uint32_t ENC_TAB_SMEM = 16 * 1024;          // {f,start} table: smem when A*A*4 fits
constexpr uint32_t ENC_SCRATCH = 256 * 256 * 4 + 256 * 1024;   // per-warp: full table + table text

struct EncSmem {
    uint32_t cnt[256];        // order-0 histogram / row scratch
    uint32_t cum[257];
    uint8_t  idxof[256];      // byte -> compact index
    uint8_t  symof[256];      // compact index -> byte
    uint32_t tab[ENC_TAB_SMEM / 4];
};

__device__ int vput(uint8_t *p, uint32_t v)          // var_put_u32, varint.h:206
{
    int n = 0;
    if (v >= (1u << 28)) p[n++] = ((v >> 28) & 0x7f) | 128;
    if (v >= (1u << 21)) p[n++] = ((v >> 21) & 0x7f) | 128;
    if (v >= (1u << 14)) p[n++] = ((v >> 14) & 0x7f) | 128;
    if (v >= (1u << 7))  p[n++] = ((v >> 7) & 0x7f) | 128;
    p[n++] = v & 0x7f;
    return n;
}

// Warp: normalise cnt[0..n) (n <= 256, total > 0) in place to sum exactly M with every non-zero
// count >= 1.  Returns false if impossible (more symbols than M).
__device__ bool normalise_row(uint32_t *cnt, int n, uint32_t M)
{
    const uint32_t lane = hgpu_lane();
    uint64_t tot = 0;
    uint32_t present = 0;
    for (int j = lane; j < n; j += 32) { tot += cnt[j]; present += cnt[j] != 0; }
    for (int d = 16; d > 0; d >>= 1) { tot += __shfl_xor_sync(0xffffffffu, tot, d); present += __shfl_xor_sync(0xffffffffu, present, d); }
    if (tot == 0) return true;
    if (present > M) return false;
    __syncwarp();
    int64_t sum = 0;
    uint32_t best = 0, besti = 0;
    for (int j = lane; j < n; j += 32) {
        uint32_t c = cnt[j];
        if (!c) continue;
        uint32_t f = (uint32_t)(((uint64_t)c * M) / tot);
        if (f == 0) f = 1;
        cnt[j] = f;
        sum += f;
        if (f > best) { best = f; besti = j; }
    }
    for (int d = 16; d > 0; d >>= 1) {
        sum += __shfl_xor_sync(0xffffffffu, sum, d);
        uint32_t ob = __shfl_xor_sync(0xffffffffu, best, d), oi = __shfl_xor_sync(0xffffffffu, besti, d);
        if (ob > best || (ob == best && oi < besti)) { best = ob; besti = oi; }
    }
    __syncwarp();
    int64_t diff = (int64_t)M - sum;
    if (diff > 0) { if (lane == 0) cnt[besti] += (uint32_t)diff; __syncwarp(); return true; }
    // too much (the f>=1 bumps): take it back from the largest entries
    while (diff < 0) {
        uint32_t b = 0, bi = 0;
        for (int j = lane; j < n; j += 32) { uint32_t f = cnt[j]; if (f > b) { b = f; bi = j; } }
        for (int d = 16; d > 0; d >>= 1) {
            uint32_t ob = __shfl_xor_sync(0xffffffffu, b, d), oi = __shfl_xor_sync(0xffffffffu, bi, d);
            if (ob > b || (ob == b && oi < bi)) { b = ob; bi = oi; }
        }
        if (b <= 1) return false;
        uint32_t take = (uint32_t)min((int64_t)(b - 1), -diff);
        __syncwarp();
        if (lane == 0) cnt[bi] -= take;
        __syncwarp();
        diff += take;
    }
    return true;
}

// alphabet writer (encode_alphabet, rANS_static16_int.h:165-189); present[] != 0 marks symbols
__device__ int put_alphabet(uint8_t *cp, const uint32_t *present)
{
    uint8_t *op = cp;
    int rle = 0;
    for (int j = 0; j < 256; j++) {
        if (!present[j]) continue;
        if (rle) { rle--; continue; }
        *cp++ = (uint8_t)j;
        if (j && present[j - 1]) {
            int r = j + 1;
            while (r < 256 && present[r]) r++;
            rle = r - (j + 1);
            *cp++ = (uint8_t)rle;
        }
    }
    *cp++ = 0;
    return (int)(cp - op);
}

__device__ __forceinline__ uint32_t enc_put(uint32_t x, uint32_t f, uint32_t start, uint32_t shift)
{
    return ((x / f) << shift) + (x % f) + start;
}

// One stream.  Returns bytes written, or 0 on failure (capacity).
__device__ uint32_t encode_stream(EncSmem &s, uint8_t *scratch, const uint8_t *in, uint32_t U, uint32_t want,
                                  uint8_t *out, uint32_t cap)
{
    const uint32_t lane = hgpu_lane();
    const uint32_t N = (want & 4) ? 32 : 4;
    const uint32_t order = want & 1;
    if (cap < 16) return 0;
    uint32_t hdr = 0;
    if (lane == 0) { out[0] = (uint8_t)((want & 5)); hdr = 1 + vput(out + 1, U); }
    hdr = __shfl_sync(0xffffffffu, hdr, 0);
    if (U == 0) return hdr;                                         // empty stream: format byte + size only
    // ---- alphabet (order-0 histogram, warp-aggregated) ----
    __syncwarp();
    for (int j = lane; j < 256; j += 32) s.cnt[j] = 0;
    __syncwarp();
    for (uint32_t base = 0; base < U; base += 32) {
        uint32_t p = base + lane;
        bool act = p < U;
        uint32_t b = act ? in[p] : 0x100;
        uint32_t peers = __match_any_sync(0xffffffffu, b);
        if (act && (peers & hgpu_lanemask_lt()) == 0) atomicAdd(&s.cnt[b], __popc(peers));
    }
    __syncwarp();
    uint8_t *tbl = scratch + 256 * 256 * 4;                          // table text
    uint32_t tlen = 0;
    uint32_t shift = 12;
    uint32_t *tab = nullptr;
    uint32_t A = 0;
    const uint32_t seg = U / N;
    if (order == 0) {
        if (!normalise_row(s.cnt, 256, 4096)) return 0;
        if (lane == 0) {
            int n = put_alphabet(tbl, s.cnt);
            for (int j = 0; j < 256; j++) if (s.cnt[j]) n += vput(tbl + n, s.cnt[j]);
            tlen = (uint32_t)n;
            uint32_t x = 0;
            for (int j = 0; j < 256; j++) { s.cum[j] = x; x += s.cnt[j]; }
        }
        tlen = __shfl_sync(0xffffffffu, tlen, 0);
        __syncwarp();
    } else {
        // compact alphabet = every byte that occurs, plus 0 (the start context)
        if (lane == 0) {
            uint32_t a = 0;
            for (int j = 0; j < 256; j++) {
                if (s.cnt[j] || j == 0) { s.idxof[j] = (uint8_t)a; s.symof[a] = (uint8_t)j; a++; }
            }
            A = a;
        }
        A = __shfl_sync(0xffffffffu, A, 0);
        shift = A <= 128 ? 10 : 12;
        tab = A * A * 4 <= ENC_TAB_SMEM ? s.tab : reinterpret_cast<uint32_t *>(scratch);
        __syncwarp();
        for (uint32_t j = lane; j < A * A; j += 32) tab[j] = 0;
        __syncwarp();
        // order-1 histogram: context = previous byte of the same segment, 0 at a segment start
        for (uint32_t base = 0; base < U; base += 32) {
            uint32_t p = base + lane;
            bool act = p < U;
            uint32_t key = 0xffffffffu;
            if (act) {
                uint32_t z = seg ? p / seg : N - 1;
                if (z > N - 1) z = N - 1;
                uint32_t c = (p == z * seg) ? 0u : in[p - 1];
                key = (uint32_t)s.idxof[c] * A + s.idxof[in[p]];
            }
            uint32_t peers = __match_any_sync(0xffffffffu, key);
            if (act && (peers & hgpu_lanemask_lt()) == 0) atomicAdd(&tab[key], __popc(peers));
        }
        __syncwarp();
        __threadfence_block();
        // normalise every row; serialise the table (decode_freq1's format)
        if (lane == 0) {
            tbl[0] = (uint8_t)(shift << 4);
            uint32_t pres[256];
            for (int j = 0; j < 256; j++) pres[j] = (s.cnt[j] || j == 0) ? 1 : 0;
            tlen = 1 + put_alphabet(tbl + 1, pres);
        }
        tlen = __shfl_sync(0xffffffffu, tlen, 0);
        for (uint32_t r = 0; r < A; r++) {
            uint32_t *row = tab + r * A;
            if (!normalise_row(row, (int)A, 1u << shift)) return 0;
            __syncwarp();
            if (lane == 0) {
                uint8_t *cp = tbl + tlen;
                int dz = 0;
                for (uint32_t k = 0; k < A; k++) {                    // encode_freq_d (:278-307)
                    if (row[k]) {
                        if (dz) { cp -= dz - 1; *cp++ = (uint8_t)(dz - 1); }
                        dz = 0;
                        cp += vput(cp, row[k]);
                    } else { dz++; *cp++ = 0; }
                }
                if (dz) { cp -= dz - 1; *cp++ = (uint8_t)(dz - 1); }
                tlen = (uint32_t)(cp - tbl);
                // row -> {f | start<<16}
                uint32_t x = 0;
                for (uint32_t k = 0; k < A; k++) { uint32_t f = row[k]; row[k] = f | (x << 16); x += f; }
            }
            tlen = __shfl_sync(0xffffffffu, tlen, 0);
            __syncwarp();
        }
        __threadfence_block();
    }
    const uint32_t body = hdr + tlen + 4 * N;                         // first byte after the states
    if ((uint64_t)body + 2ull * U + 64 > cap) {
        // not enough room to encode safely in place: the caller's bound is too small
        if ((uint64_t)hdr + U > cap) return 0;
    }
    // ---- encode backwards; words are laid downwards from the end of the slot ----
    uint32_t wp = cap & ~1u;                                          // byte offset of the lowest word written so far
    uint32_t x = RANS_L;
    const bool mine = lane < N;
    bool overflow = false;
#define ENC_EMIT(ACTIVE, F, START)                                                                 \
    {                                                                                              \
        bool need = (ACTIVE) && x >= ((F) << (31 - shift));                                        \
        uint32_t bal = __ballot_sync(0xffffffffu, need);                                           \
        if (bal) {                                                                                 \
            uint32_t cntw = __popc(bal);                                                           \
            if (wp < body + 2 * cntw) overflow = true;                                             \
            else {                                                                                 \
                wp -= 2 * cntw;                                                                    \
                if (need) {                                                                        \
                    uint32_t o = wp + 2 * __popc(bal & hgpu_lanemask_lt());                        \
                    out[o] = (uint8_t)x; out[o + 1] = (uint8_t)(x >> 8);                           \
                    x >>= 16;                                                                      \
                }                                                                                  \
            }                                                                                      \
        }                                                                                          \
        if (ACTIVE) x = enc_put(x, (F), (START), shift);                                           \
    }
    if (order == 0) {
        // symbol i belongs to state i % N; decode order is i ascending, so encode i descending.
        uint32_t rows = (U + N - 1) / N;
        for (uint32_t rr = rows; rr-- > 0;) {
            uint32_t i = rr * N + lane;
            bool act = mine && i < U;
            uint32_t b = act ? in[i] : 0;
            uint32_t f = act ? s.cnt[b] : 1, st = act ? s.cum[b] : 0;
            ENC_EMIT(act, f, st)
            if (overflow) break;
        }
    } else {
        const bool last = lane == N - 1;
        // the last state's tail first (it is decoded last)
        for (uint32_t p = U; p-- > seg * N;) {
            uint32_t c = (p == (N - 1) * seg) ? 0u : in[p - 1];
            uint32_t e = tab[(uint32_t)s.idxof[c] * A + s.idxof[in[p]]];
            ENC_EMIT(last, e & 0xffffu, e >> 16)
            if (overflow) break;
        }
        const uint8_t *ip = in + (size_t)(mine ? lane : 0) * seg;
        for (uint32_t q = seg; q-- > 0 && !overflow;) {
            uint32_t c = q ? ip[q - 1] : 0u;
            uint32_t e = tab[(uint32_t)s.idxof[c] * A + s.idxof[ip[q]]];
            ENC_EMIT(mine, e & 0xffffu, e >> 16)
        }
    }
#undef ENC_EMIT
    uint32_t total = 0;
    if (!overflow) total = body + ((cap & ~1u) - wp);
    if (overflow || total >= hdr + U) {
        // CAT fallback (rANS_static4x16pr.c:1539-1553): format byte | 0x20, raw payload
        if ((uint64_t)hdr + U > cap) return 0;
        __syncwarp();
        if (lane == 0) out[0] = (uint8_t)((want & 4) | 0x20);
        for (uint32_t i = lane; i < U; i += 32) out[hdr + i] = in[i];
        return hdr + U;
    }
    // table, states, then the words moved up behind them
    __syncwarp();
    for (uint32_t i = lane; i < tlen; i += 32) out[hdr + i] = tbl[i];
    if (mine) {
        uint8_t *q = out + hdr + tlen + 4 * lane;
        q[0] = (uint8_t)x; q[1] = (uint8_t)(x >> 8); q[2] = (uint8_t)(x >> 16); q[3] = (uint8_t)(x >> 24);
    }
    uint32_t nw = (cap & ~1u) - wp;
    // forward copy in 32-byte strides: destination is below the source, chunks never overtake
    for (uint32_t i0 = 0; i0 < nw; i0 += 32) {
        uint32_t i = i0 + lane;
        uint8_t v = i < nw ? out[wp + i] : 0;
        __syncwarp();
        if (i < nw) out[body + i] = v;
        __syncwarp();
    }
    return total;
}

__global__ void __launch_bounds__(32)
rans_nx16_encode_kernel(const uint8_t *__restrict__ in, const uint64_t *__restrict__ in_off,
                        const uint32_t *__restrict__ in_len, const uint32_t *__restrict__ order, uint32_t n,
                        uint8_t *out, const uint64_t *__restrict__ out_off, const uint32_t *__restrict__ out_cap,
                        uint32_t *out_len, int32_t *status, uint8_t *scratch, uint32_t *counter)
{
    __shared__ EncSmem s;
    uint8_t *my = scratch + (size_t)blockIdx.x * ENC_SCRATCH;
    for (;;) {
        uint32_t job = 0;
        if (hgpu_lane() == 0) job = atomicAdd(counter, 1u);
        job = __shfl_sync(0xffffffffu, job, 0);
        if (job >= n) break;
        uint32_t got = encode_stream(s, my, in + in_off[job], in_len[job], order[job], out + out_off[job], out_cap[job]);
        __syncwarp();
        if (hgpu_lane() == 0) { out_len[job] = got; status[job] = got ? HGPU_OK : HGPU_RANS_ERR; }
    }
}

} // namespace

extern "C" uint32_t hgpu_rans_nx16_compress_bound(uint32_t size, int order)
{
    // same shape as rans_compress_bound_4x16 (rANS_static4x16pr.c:1203): payload slack + table + states
    uint64_t b = (uint64_t)(1.05 * size) + 257 * 3 + 4 + 64;
    if (order & 1) b += 257 * 257 * 3;
    b += (order & 4) ? 32 * 4 : 4 * 4;
    b += 20;
    return b > 0xffffffffull ? 0xffffffffu : (uint32_t)b;
}

extern "C" int hgpu_rans_nx16_encode_batch_dev(hgpu_ctx *ctx, const uint8_t *d_in, const uint64_t *d_in_off,
        const uint32_t *d_in_len, const uint32_t *d_order, uint32_t n, uint8_t *d_out, const uint64_t *d_out_off,
        const uint32_t *d_out_cap, uint32_t *d_out_len, int32_t *d_status, void *stream)
{
    if (!ctx) { hgpu_set_error("null context"); return HGPU_ERR_ARG; }
    if (n == 0) return HGPU_OK;
    cudaStream_t st = stream ? (cudaStream_t)stream : ctx->stream;
    int per_sm = 0;
    if (hgpu_check(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, rans_nx16_encode_kernel, 32, 0), "enc occupancy"))
        return HGPU_ERR_CUDA;
    if (per_sm < 1) per_sm = 1;
    uint32_t grid = (uint32_t)ctx->sm_count * (uint32_t)per_sm;
    if (grid > n) grid = n;
    int rc = hgpu_ensure_scratch(ctx, (size_t)grid * ENC_SCRATCH);
    if (rc) return rc;
    uint32_t *counter = hgpu_take_counter(ctx, st);
    if (!counter) return HGPU_ERR_CUDA;
    rans_nx16_encode_kernel<<<grid, 32, 0, st>>>(d_in, d_in_off, d_in_len, d_order, n, d_out, d_out_off, d_out_cap,
                                                d_out_len, d_status, ctx->d_scratch, counter);
    hgpu_count_launch();
    return hgpu_check(cudaGetLastError(), "rans encode launch");
}
