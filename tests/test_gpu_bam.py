"""GPU parity: BAM record index + unpack (hgpu_bam_*_dev) against the oracle (itself pinned to the
reference's bam_read1): every bam1_core_t field, every data byte, ASCII SEQ and QUAL+33."""
import ctypes as C
import glob, os, random, struct, sys, zlib
import numpy as np
import pytest
import htslib_b200 as H
from _libs import GOLD, ROOT, BamCore, bam_header, bam_header_len, bgzf_file, orc_bam_unpack_all, orc_bgzf_scan
sys.path.insert(0, ROOT)
from tools import synth

pytestmark = pytest.mark.gpu
BAMS = sorted(glob.glob(os.path.join(GOLD, "htslib", "bgzf_boundaries", "*.bam"))) + \
    [os.path.join(GOLD, "htslib", "range.bam"), os.path.join(GOLD, "htslib", "colons.bam")]


@pytest.fixture(scope="module")
def ctx():
    c = H.Context(0)
    yield c
    c.close()


def gpu_unpack(ctx, stream, hints=None):
    import torch
    dev = torch.device("cuda:0")
    d = torch.from_numpy(np.frombuffer(stream + b"\0" * 8, dtype=np.uint8).copy()).to(dev)
    dh = None
    if hints is not None:
        dh = torch.from_numpy(np.array(hints, dtype=np.int64)).to(dev)
    r = ctx.bam_unpack_dev(d, len(stream), dh)
    torch.cuda.synchronize()
    n = r["n"]
    core = r["core"].cpu().numpy()
    data = r["data"].cpu().numpy(); doff = r["data_off"].cpu().numpy()
    seq = r["seq"].cpu().numpy(); qual = r["qual"].cpu().numpy(); soff = r["seq_off"].cpu().numpy()
    st = r["status"].cpu().numpy(); roff = r["rec_off"].cpu().numpy()
    out = []
    for i in range(n):
        c = BamCore.from_buffer_copy(core[i].tobytes())
        out.append((int(st[i]), c.astuple(), data[doff[i]:doff[i + 1]].tobytes(), seq[soff[i]:soff[i + 1]].tobytes(),
                    qual[soff[i]:soff[i + 1]].tobytes()))
    return out, [int(x) for x in roff[:n]]


def compare(ctx, stream, hints=None):
    want, woffs = orc_bam_unpack_all(stream)
    got, goffs = gpu_unpack(ctx, stream, hints)
    assert goffs == woffs
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert g[0] == w[0]
        if w[0] >= 0:
            assert g[1:] == w[1:]
    return len(got)


@pytest.mark.parametrize("path", BAMS, ids=os.path.basename)
def test_fixture_bams_with_block_hints(ctx, path):
    img = open(path, "rb").read()
    _, blocks = orc_bgzf_scan(img)
    parts = [zlib.decompress(img[o + 18:o + l - 8], -15) for o, l in blocks]
    stream = b"".join(parts)
    h = bam_header_len(stream)
    body = stream[h:]
    # hints = inflated block starts relative to the record stream (many are NOT record starts in
    # the bgzf_boundaries files: records straddle blocks there)
    starts, p = [0], 0
    for part in parts:
        p += len(part)
        if p > h and p - h < len(body):
            starts.append(p - h)
    assert compare(ctx, body, sorted(set(starts))) > 0
    assert compare(ctx, body, None) > 0                   # no hints: serial walk, same answer


def test_synthetic_and_odd_records(ctx):
    stream, offs = synth.bam_records(3, 5000)
    rng = random.Random(4)
    extra = bytearray()
    for k in range(64):
        name = bytes(rng.choice(b"abcXYZ019") for _ in range(k % 8)) + (b"\0" if k % 5 else b"Q")
        lq = [0, 1, 7, 150, 151, 300][k % 6]
        cig = [] if k % 4 == 0 else ([(lq << 4) | 0] if lq else [])
        if k == 13: cig = [(lq << 4) | 4]                 # <l_qseq>S : bam_tag2cigar trigger -> status 1
        flag = 4 if not cig else 0
        seq = bytes(rng.randrange(256) for _ in range((lq + 1) // 2))
        qual = bytes([0xff] * lq) if k % 7 == 0 else bytes(rng.randrange(42) for _ in range(lq))
        aux = b"XAZ" + b"hi\0" if k % 2 else b""
        body = struct.pack("<iiBBHHHiiii", max(0, k % 3 - 1), 1000 + k * 50, len(name), 30, 4681, len(cig), flag, lq, -1, -1, 0)
        body += name + struct.pack("<%dI" % len(cig), *cig) + seq + qual + aux
        extra += struct.pack("<i", len(body)) + body
    full = stream + bytes(extra)
    hints = [0] + [offs[i] for i in range(200, 5000, 200)] + [offs[777] + 3, len(stream) + 1]   # two wrong hints
    n = compare(ctx, full, sorted(hints))
    assert n == 5064


def test_invalid_records_and_broken_chain(ctx):
    stream, offs = synth.bam_records(5, 20)
    good = bytearray(stream[:offs[5]])
    b = bytearray(good); struct.pack_into("<i", b, offs[4] + 4 + 16, -5)
    got, _ = gpu_unpack(ctx, bytes(b))
    assert [g[0] for g in got] == [0, 0, 0, 0, -4]
    b = bytearray(good); struct.pack_into("<I", b, offs[4] + 36 + good[offs[4] + 12], (149 << 4))
    got, _ = gpu_unpack(ctx, bytes(b))
    assert [g[0] for g in got] == [0, 0, 0, 0, -4]
    with pytest.raises(H.HgpuError):
        gpu_unpack(ctx, bytes(good[:-7]))                  # truncated last record
    b = bytearray(good); struct.pack_into("<i", b, offs[2], 5)
    with pytest.raises(H.HgpuError):
        gpu_unpack(ctx, bytes(b))                          # block_size < 32


def test_inflate_then_unpack_pipeline(ctx):
    """BGZF inflate output feeds the unpacker directly on the device (BAM->SAM front half)."""
    import torch
    dev = torch.device("cuda:0")
    corpus = synth.bam_bgzf_corpus(30e6, procs=1, reads_per_shard=30000)
    comp, clen, ulen = corpus["comp"], corpus["clen"], corpus["ulen"]
    nb = len(clen)
    in_off = np.concatenate([[0], np.cumsum(clen.astype(np.int64))[:-1]]).astype(np.int64)
    out_off = np.concatenate([[0], np.cumsum(ulen.astype(np.int64))[:-1]]).astype(np.int64)
    U = int(ulen.astype(np.int64).sum())
    d_in = torch.zeros(comp.size + 64, dtype=torch.uint8, device=dev); d_in[:comp.size].copy_(torch.from_numpy(comp.copy()))
    d_out = torch.zeros(U + 64, dtype=torch.uint8, device=dev)
    t = lambda a: torch.from_numpy(a).to(dev)
    d_len = torch.zeros(nb, dtype=torch.int32, device=dev); d_st = torch.zeros(nb, dtype=torch.int32, device=dev)
    d_oo = t(out_off)
    ctx.bgzf_inflate_dev(d_in, t(in_off), t(clen.view(np.int32)), d_out, d_oo, t(ulen.view(np.int32)), d_len, d_st)
    torch.cuda.synchronize()
    assert int(d_st.abs().sum()) == 0
    r = ctx.bam_unpack_dev(d_out, U, d_oo)
    torch.cuda.synchronize()
    assert r["n"] == corpus["n_reads"]
    assert int(r["status"].abs().sum()) == 0
    # size-independent properties: every read is 150 bp, SEQ is ACGT only, QUAL+33 is one of the four NovaSeq bins
    assert int(r["seq_off"][-1]) == 150 * r["n"]
    sq = r["seq"].cpu().numpy()
    assert set(np.unique(sq).tolist()) <= set(b"ACGT")
    assert set(np.unique(r["qual"].cpu().numpy()).tolist()) <= {2 + 33, 12 + 33, 23 + 33, 37 + 33}
    # and the first 500 records agree with the oracle byte for byte
    stream = d_out[:U].cpu().numpy().tobytes()
    cut = int(r["rec_off"][500].item())
    want, woffs = orc_bam_unpack_all(stream[:cut])
    core = r["core"].cpu().numpy(); data = r["data"].cpu().numpy(); doff = r["data_off"].cpu().numpy()
    for i in range(500):
        assert BamCore.from_buffer_copy(core[i].tobytes()).astuple() == want[i][1]
        assert data[doff[i]:doff[i + 1]].tobytes() == want[i][2]
