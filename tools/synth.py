"""Deterministic synthetic 150 bp short-read data for bench.py and the large-size tests
(SURVEY.md §8d "Concrete synthetic inputs").  Not the oracle and not the product: it only
manufactures inputs.  BGZF blocks are produced the way htslib's writer does on its zlib arm —
bam_write1's bgzf_flush_try rule (records are not split across blocks, payload <= 0xff00,
sam.c:888, bgzf.c:1996) and bgzf_compress = deflateInit2(level, Z_DEFLATED, -15, 8)
(bgzf.c:624-683) — using the very same system zlib through Python's zlib module.
"""
import struct
import zlib
import numpy as np

BGZF_BLOCK_SIZE = 0xff00
BGZF_EOF = bytes.fromhex("1f8b08040000000000ff0600424302001b0003000000000000000000")
NT16 = np.zeros(256, dtype=np.uint8)
for _i, _c in enumerate(b"=ACMGRSVTWYHKDBN"):
    NT16[_c] = _i


def novaseq_quals(rng, n, change=0.06):
    """NovaSeq-like 4-bin qualities {2,12,23,37}: a run-structured Markov chain (cf. htscodecs
    tests/dat/q4).  Returns raw Phred bytes (not +33)."""
    bins = np.array([2, 12, 23, 37], dtype=np.uint8)
    flips = rng.random(n) < change
    draw = rng.choice(4, size=n, p=[0.04, 0.10, 0.22, 0.64])
    idx = np.maximum.accumulate(np.where(flips, np.arange(n), 0))
    state = np.where(idx > 0, draw[idx], 3)
    return bins[state]


def hiseq_quals(rng, n):
    """HiSeq-like ~40-value qualities with positional decay (cf. tests/dat/q40+dir)."""
    base = 38 - (np.arange(n) % 150) * (10.0 / 150)
    q = np.clip(np.rint(base + rng.normal(0, 4, size=n)), 2, 41).astype(np.uint8)
    return q


def reg2bin(beg, end):
    end -= 1
    if beg >> 14 == end >> 14: return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17: return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20: return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23: return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26: return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def bam_records(seed, n_reads, read_len=150, quals="novaseq", tid=0, pos0=10000):
    """n_reads BAM records (block_size + core + data), coordinate-sorted 30x-like (step ~5 bp).
    Returns (bytes, list of record byte offsets)."""
    rng = np.random.default_rng(seed)
    step = rng.integers(1, 10, size=n_reads)
    pos = pos0 + np.cumsum(step)
    span = int(pos[-1] - pos0) + read_len + 16
    ref = rng.integers(0, 4, size=span, dtype=np.uint8)
    ref_c = np.frombuffer(b"ACGT", dtype=np.uint8)[ref]
    q = (novaseq_quals if quals == "novaseq" else hiseq_quals)(rng, n_reads * read_len).reshape(n_reads, read_len)
    subs = rng.random((n_reads, read_len)) < 0.001
    sub_base = rng.integers(0, 4, size=(n_reads, read_len), dtype=np.uint8)
    tile = 1101 + rng.integers(0, 78, size=n_reads).cumsum() // 4000
    xs = rng.integers(1000, 32000, size=n_reads)
    ys = rng.integers(1000, 50000, size=n_reads)
    lane = 1 + (seed % 4)
    mapq = rng.choice(np.array([0, 20, 40, 60]), size=n_reads, p=[0.03, 0.03, 0.06, 0.88])
    flags = rng.choice(np.array([99, 147, 83, 163]), size=n_reads)
    indel = rng.random(n_reads) < 0.015
    tlen = rng.integers(300, 600, size=n_reads)
    out = bytearray()
    offs = []
    for i in range(n_reads):
        p = int(pos[i])
        o = p - pos0
        bases = ref_c[o:o + read_len].copy()
        if subs[i].any():
            m = subs[i]
            bases[m] = np.frombuffer(b"ACGT", dtype=np.uint8)[sub_base[i][m]]
        nm = int(subs[i].sum())
        if indel[i]:
            k = 20 + (xs[i] % 100)
            cigar = [(k << 4) | 0, (1 << 4) | 1, ((read_len - k - 1) << 4) | 0]       # kM 1I (L-k-1)M
            rlen = read_len - 1
            nm += 1
            md = b"%d" % rlen
        else:
            cigar = [(read_len << 4) | 0]
            rlen = read_len
            md = b"%d" % read_len
        nib = NT16[bases]
        if read_len & 1:
            nib = np.append(nib, 0)
        packed = ((nib[0::2] << 4) | nib[1::2]).astype(np.uint8).tobytes()
        name = b"A00123:45:HXXXXDSXX:%d:%d:%d:%d\0" % (lane, tile[i], xs[i], ys[i])
        fl = int(flags[i])
        mate = p + int(tlen[i]) - read_len if fl in (99, 163) else max(0, p - int(tlen[i]) + read_len)
        tl = int(tlen[i]) if fl in (99, 163) else -int(tlen[i])
        aux = b"NMC" + bytes([min(nm, 255)]) + b"MDZ" + md + b"\0" + b"RGZgrp1\0" + b"ASC" + bytes([min(255, read_len - 2 * nm)])
        body = struct.pack("<iiBBHHHiiii", tid, p, len(name), int(mapq[i]), reg2bin(p, p + rlen), len(cigar), fl,
                           read_len, tid, mate, tl)
        body += name + struct.pack("<%dI" % len(cigar), *cigar) + packed + q[i].tobytes() + aux
        offs.append(len(out))
        out += struct.pack("<i", len(body)) + body
    return bytes(out), offs


def bgzf_block(payload, level=6):
    if level == 0:
        raw = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xffff) + payload
    else:
        c = zlib.compressobj(level, zlib.DEFLATED, -15, 8)
        raw = c.compress(payload) + c.flush()
        if 18 + len(raw) + 8 > 65536:      # zlib arm falls back to a stored block (bgzf.c:653-667)
            raw = b"\x01" + struct.pack("<HH", len(payload), len(payload) ^ 0xffff) + payload
    hdr = b"\x1f\x8b\x08\x04\0\0\0\0\0\xff\x06\0BC\x02\0" + struct.pack("<H", 18 + len(raw) + 8 - 1)
    return hdr + raw + struct.pack("<II", zlib.crc32(payload) & 0xffffffff, len(payload))


def bgzf_pack_records(stream, offs, level=6):
    """Cut a record stream into BGZF blocks like bam_write1 + bgzf_flush_try: a record that does
    not fit in the space left starts a new block.  Returns list of (compressed block, payload len)."""
    blocks = []
    start = 0
    ends = offs[1:] + [len(stream)]
    cur = 0
    for e in ends:
        if e - start > BGZF_BLOCK_SIZE and cur > start:
            blocks.append((bgzf_block(stream[start:cur], level), cur - start))
            start = cur
        cur = e
        while cur - start > BGZF_BLOCK_SIZE:          # one record larger than a block: split like bgzf_write
            blocks.append((bgzf_block(stream[start:start + BGZF_BLOCK_SIZE], level), BGZF_BLOCK_SIZE))
            start += BGZF_BLOCK_SIZE
    if cur > start:
        blocks.append((bgzf_block(stream[start:cur], level), cur - start))
    return blocks


def bam_shard(args):
    """Worker: (seed, n_reads, level, quals) -> (compressed bytes, block lengths, payload lengths, crc of payload)."""
    seed, n_reads, level, quals = args
    stream, offs = bam_records(seed, n_reads, quals=quals, pos0=10000 + seed * 7)
    blocks = bgzf_pack_records(stream, offs, level)
    comp = b"".join(b for b, _ in blocks)
    return comp, [len(b) for b, _ in blocks], [u for _, u in blocks], zlib.crc32(stream), n_reads


def bam_bgzf_corpus(total_uncompressed, level=6, quals="novaseq", seed=42, procs=None, reads_per_shard=60000):
    """Synthetic BGZF-compressed BAM record stream of ~total_uncompressed bytes.
    Returns dict(comp=np.uint8[], clen=u32[], ulen=u32[], crcs=[...], n_reads)."""
    import multiprocessing as mp
    import os
    per_read = 336
    n_shards = max(1, int(total_uncompressed / (per_read * reads_per_shard) + 0.5))
    if total_uncompressed < per_read * reads_per_shard:
        reads_per_shard = max(1, int(total_uncompressed // per_read)); n_shards = 1
    jobs = [(seed * 100003 + s, reads_per_shard, level, quals) for s in range(n_shards)]
    procs = procs or min(len(jobs), len(os.sched_getaffinity(0)))
    if procs > 1:
        with mp.get_context("fork").Pool(procs) as pool:
            res = pool.map(bam_shard, jobs, chunksize=1)
    else:
        res = [bam_shard(j) for j in jobs]
    comp = np.frombuffer(b"".join(r[0] for r in res), dtype=np.uint8)
    clen = np.array([x for r in res for x in r[1]], dtype=np.uint32)
    ulen = np.array([x for r in res for x in r[2]], dtype=np.uint32)
    return dict(comp=comp, clen=clen, ulen=ulen, shard_crcs=[r[3] for r in res], shard_blocks=[len(r[1]) for r in res],
                n_reads=sum(r[4] for r in res))
