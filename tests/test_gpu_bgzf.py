"""GPU parity: BGZF inflate + CRC32 (hgpu_bgzf_inflate_*, hts_crc32) against the oracle, zlib
and the reference's fixtures — bit exact, per-block status codes included."""
import glob, os, random, struct, zlib
import numpy as np
import pytest
import htslib_b200 as H
from _libs import (GOLD, BGZF_EOF, bgzf_block, bgzf_file, orc_bgzf_inflate_block, orc_bgzf_scan)
from test_oracle_bgzf import payloads

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = H.Context(0)
    yield c
    c.close()


def gpu_blocks(ctx, blocks, caps=None):
    """Inflate a list of whole BGZF blocks through the device-pointer API; returns [(status, bytes)]."""
    import torch
    dev = torch.device("cuda:0")
    n = len(blocks)
    in_len = np.array([len(b) for b in blocks], dtype=np.uint32)
    in_off = np.zeros(n, dtype=np.uint64)
    p = 0
    for i in range(n):
        in_off[i] = p; p += len(blocks[i]) + (i * 5) % 7
    blob = np.zeros(p + 8, dtype=np.uint8)
    for i, b in enumerate(blocks):
        blob[int(in_off[i]):int(in_off[i]) + len(b)] = np.frombuffer(b, dtype=np.uint8)
    cap = np.array(caps if caps is not None else [65536] * n, dtype=np.uint32)
    out_off = (np.arange(n, dtype=np.uint64) * 65536)
    t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).to(dev)
    d_in = torch.from_numpy(blob).to(dev)
    d_out = torch.full((n * 65536 + 8,), 0x55, dtype=torch.uint8, device=dev)
    d_len = torch.zeros(n, dtype=torch.int32, device=dev)
    d_st = torch.full((n,), 9, dtype=torch.int32, device=dev)
    a, b, c, d = t(in_off), t(in_len), t(out_off), t(cap)
    ctx.bgzf_inflate_dev(d_in, a, b, d_out, c, d, d_len, d_st, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    out = d_out.cpu().numpy(); ln = d_len.cpu().numpy(); st = d_st.cpu().numpy()
    return [(int(st[i]), out[i * 65536:i * 65536 + int(ln[i])].tobytes()) for i in range(n)]


def test_levels_and_payload_shapes(ctx):
    blocks, want = [], []
    for level in (0, 1, 2, 4, 6, 9):
        rng = random.Random(level)
        for p in payloads(rng):
            p = p[:65280]
            blocks.append(bgzf_block(p, level)); want.append(p)
    res = gpu_blocks(ctx, blocks)
    for i, ((st, data), w) in enumerate(zip(res, want)):
        assert st == 0, i
        assert data == w, i


def test_fixed_stored_and_multiblock_members(ctx):
    rng = random.Random(3)
    p = bytes(rng.choice(b"abcdefgh ") for _ in range(30000))
    blocks = []
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 8, zlib.Z_FIXED)
    blocks.append(bgzf_block(p, raw_deflate=c.compress(p) + c.flush()))
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 8)
    raw = b""
    for i in range(0, len(p), 1000):
        raw += c.compress(p[i:i + 1000]) + c.flush(zlib.Z_FULL_FLUSH)
    raw += c.flush()
    blocks.append(bgzf_block(p, raw_deflate=raw))
    c = zlib.compressobj(9, zlib.DEFLATED, -15, 9, zlib.Z_HUFFMAN_ONLY)
    blocks.append(bgzf_block(p, raw_deflate=c.compress(p) + c.flush()))
    c = zlib.compressobj(6, zlib.DEFLATED, -15, 1)                 # memLevel 1: many small deflate blocks
    blocks.append(bgzf_block(p, raw_deflate=c.compress(p) + c.flush()))
    blocks.append(BGZF_EOF)
    blocks.append(bgzf_block(b""))
    res = gpu_blocks(ctx, blocks)
    for (st, data), w in zip(res, [p, p, p, p, b"", b""]):
        assert st == 0 and data == w


def test_status_codes_match_oracle(ctx):
    rng = random.Random(17)
    p = bytes(rng.choice(b"ACGTN\n") for _ in range(20000))
    good = bgzf_block(p)
    blocks = [good]
    for t in range(300):
        b = bytearray(good)
        k = rng.randrange(len(b)) if t >= 40 else rng.randrange(18)      # the first 40 hit the 18-byte header
        b[k] ^= 1 << rng.randrange(8)
        if k in (16, 17):
            continue                                   # BSIZE: a length error, caught by the scan
        blocks.append(bytes(b))
    res = gpu_blocks(ctx, blocks)
    seen = set()
    for blk, (st, data) in zip(blocks, res):
        rc, want = orc_bgzf_inflate_block(blk)
        if rc >= 0:
            assert st == 0 and data == want
        else:
            assert st == rc, (st, rc)
        seen.add(st)
    assert {0, -1, -2, -3} <= seen


def test_output_slot_too_small(ctx):
    p = b"0123456789" * 1000
    res = gpu_blocks(ctx, [bgzf_block(p), bgzf_block(p)], caps=[len(p), len(p) - 1])
    assert res[0] == (0, p)
    assert res[1][0] == H.BGZF_ERR_SPACE


@pytest.mark.parametrize("name", ["bgziptest.txt.gz", "range.bam", "colons.bam", "bgzf_boundaries/bgzf_boundaries1.bam",
                                  "bgzf_boundaries/bgzf_boundaries2.bam", "bgzf_boundaries/bgzf_boundaries3.bam"])
def test_reference_fixtures_file_api(ctx, name):
    img = np.fromfile(os.path.join(GOLD, "htslib", name), dtype=np.uint8)
    _, blocks = orc_bgzf_scan(img.tobytes())
    want = b"".join(orc_bgzf_inflate_block(img.tobytes()[o:o + l])[1] for o, l in blocks)
    out = np.zeros(len(want) + 64, dtype=np.uint8)
    rc, n, bad = ctx.bgzf_inflate_file_host(img, out)
    assert rc == 0 and n == len(want)
    assert out[:n].tobytes() == want
    if name == "bgziptest.txt.gz":
        assert want == open(os.path.join(GOLD, "htslib", "bgziptest.txt"), "rb").read()


def test_file_api_error_reporting(ctx):
    rng = random.Random(23)
    data = bytes(rng.choice(b"ACGT") for _ in range(300000))
    img = bytearray(bgzf_file(data, 6))
    off, ln, _ = H.bgzf_scan(np.frombuffer(bytes(img), dtype=np.uint8))
    img[int(off[2]) + int(ln[2]) - 8] ^= 0xff              # CRC of block 2
    out = np.zeros(len(data), dtype=np.uint8)
    rc, n, bad = ctx.bgzf_inflate_file_host(np.frombuffer(bytes(img), dtype=np.uint8), out)
    assert rc == H.BGZF_ERR_CRC and bad == 2
    rc, n, bad = ctx.bgzf_inflate_file_host(np.frombuffer(bytes(img[:-40]), dtype=np.uint8), out)
    assert rc == H.BGZF_ERR_HEADER


def test_crc32_device(ctx):
    rng = random.Random(5)
    for n in (1, 2, 3, 31, 32, 33, 1000, 65536, (1 << 20) + 12345, 3 * (1 << 20)):
        b = bytes(rng.randrange(256) for _ in range(min(n, 4096))) * (n // min(n, 4096) + 1)
        b = b[:n]
        assert ctx.crc32(b) == zlib.crc32(b)
        assert ctx.crc32(b, 0x12345678) == zlib.crc32(b, 0x12345678)
    a = np.frombuffer(b, dtype=np.uint8)
    assert H.lib().hts_crc32(0, a.ctypes.data, a.size) == zlib.crc32(b)


def test_large_file_roundtrip_property(ctx):
    """Full-size style check without a per-byte oracle: ~256 MB of output, every block's CRC is
    verified on the device and the whole stream's CRC32 must equal the generator's."""
    rng = np.random.default_rng(7)
    base = rng.integers(0, 4, size=1 << 22, dtype=np.uint8)
    text = np.frombuffer(b"ACGT", dtype=np.uint8)[base]
    text[::151] = 10
    unit = bgzf_file(text.tobytes(), 6, eof=False)
    reps = 64
    img = np.frombuffer(unit * reps + BGZF_EOF, dtype=np.uint8)
    out = np.zeros(text.size * reps, dtype=np.uint8)
    rc, n, bad = ctx.bgzf_inflate_file_host(img, out)
    assert rc == 0 and n == text.size * reps
    assert zlib.crc32(out[:text.size].tobytes()) == zlib.crc32(text.tobytes())
    v = out.reshape(reps, text.size)
    assert (v == v[0]).all()


@pytest.mark.parametrize("name", ["fail_block_3331.npy", "fail_block_6735.npy"])
def test_false_end_of_block_in_the_preroll_regression(ctx, name):
    """Two blocks of the 0.5 GB synthetic corpus that once failed their CRC: a sub-range decoder's 128-bit pre-roll ran into
    a false end-of-block code whose exit coincided with the true token start, so the chain looked consistent.  The failure
    depended on where the block sat in memory: every 16-byte phase of the input is tried, through the file path."""
    import zlib
    blk = np.load(os.path.join(GOLD, "bgzf_regress", name)).tobytes()
    want = zlib.decompress(blk[18:-8], -15)
    rng = random.Random(3)
    for phase in range(16):
        # a leading block whose compressed length puts `blk` at the wanted phase
        for _ in range(400):
            lead_payload = bytes(rng.randrange(256) for _ in range(rng.randrange(1, 200)))
            lead = bgzf_block(lead_payload)
            if len(lead) % 16 == phase:
                break
        else:
            pytest.skip("no leading block of phase %d found" % phase)
        img = np.frombuffer(lead + blk + BGZF_EOF, dtype=np.uint8).copy()
        out = np.zeros(len(lead_payload) + len(want) + 64, dtype=np.uint8)
        rc, n, bad = ctx.bgzf_inflate_file_host(img, out)
        assert rc == 0 and n == len(lead_payload) + len(want), (phase, rc, n, bad)
        assert out[:n].tobytes() == lead_payload + want, phase
