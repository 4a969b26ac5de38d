"""CRAM record decode (htslib_b200/csrc/cram_records.cuh: the record loop of cram_decode_slice, cram_decode_seq, cram_decode_aux,
cram_decode_slice_xref, cram_to_bam) against the compiled reference's sam_read1 on every CRAM fixture.

Two runs of the same source: `-m gpu` calls hgpu_cram_decode_records_host (the kernels); without a GPU the logic is checked through
tests/hostsim (the same __host__ __device__ code built for the host by g++, kernels replaced by loops — test infrastructure,
never loaded by htslib_b200)."""
import ctypes as C
import os
import subprocess
import numpy as np
import pytest
import htslib_b200 as H
from _libs import GOLD, ref, ref_cram_read_all, ref_write_cram
from test_cram_blocks import _expect

HT = os.path.join(GOLD, "htslib")
HERE = os.path.dirname(os.path.abspath(__file__))
CASES = [("ce#1000.v31.cram", "ce.fa"), ("ce#1000.v30.cram", "ce.fa"), ("ce#1000.v31arith.cram", "ce.fa"), ("ce#1000.v31fqz.cram", "ce.fa"),
         ("ce#5b_java.cram", "ce.fa"), ("auxf#values_java.cram", "auxf.fa"), ("xx#large_aux_java.cram", "xx.fa"), ("range.cram", "ce.fa")]


def hostsim():
    so = os.path.join(HERE, "hostsim", "_build", "libcramrec_hostsim.so")
    src = os.path.join(HERE, "..", "htslib_b200", "csrc", "cram_records.cu")
    hdr = os.path.join(HERE, "..", "htslib_b200", "csrc", "cram_records.cuh")
    if not os.path.exists(so) or os.path.getmtime(so) < max(os.path.getmtime(src), os.path.getmtime(hdr)):
        subprocess.check_call(["bash", os.path.join(HERE, "hostsim", "build.sh")], stdout=subprocess.DEVNULL)
    l = C.CDLL(so)
    l.hostsim_cram_decode_records.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_char_p, C.c_int, C.c_void_p]
    l.hostsim_last_error.restype = C.c_char_p
    return (l.hostsim_cram_decode_records, l.hostsim_cram_records_free, lambda: l.hostsim_last_error().decode())


def cpu_blocks(img):
    """The blocks uncompressed by the per-codec checkers (what hgpu_cram_uncompress_blocks_host returns on a GPU box)."""
    blocks, _ = H.cram_scan_blocks(img)
    sizes = blocks["uncomp_size"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum((sizes + 15) // 16 * 16)]).astype(np.uint64)
    udata = np.zeros(int(off[-1]) + 16, dtype=np.uint8)
    for i, b in enumerate(blocks):
        want = _expect(img, b)
        assert want is not None and len(want) == int(b["uncomp_size"]), (i, int(b["method"]))
        udata[int(off[i]):int(off[i]) + len(want)] = np.frombuffer(want, dtype=np.uint8)
    return blocks, udata, off[:-1].copy()


def compare(name, fa, got, decode_md, path=None):
    want = ref_cram_read_all(path or os.path.join(HT, name), os.path.join(HT, fa), decode_md)
    assert got["slice_status"].tolist() == [0] * len(got["slice_status"]), (name, got["slice_status"].tolist())
    assert len(got["data"]) == len(want), (name, len(got["data"]), len(want))
    names = [f for f, _ in H.BAM1_CORE_DT]
    for i, (wc, wd) in enumerate(want):
        gc = tuple(int(got["core"][i][f]) for f in names)
        assert gc == wc, (name, decode_md, i, dict(zip(names, gc)), dict(zip(names, wc)))
        assert got["data"][i] == wd, (name, decode_md, i, got["data"][i][:80], wd[:80])
        assert got["rec_status"][i] == 0


@pytest.mark.skipif(ref() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("decode_md", [0, 1])
@pytest.mark.parametrize("name,fa", CASES)
def test_hostsim_records_equal_reference(name, fa, decode_md):
    img = np.fromfile(os.path.join(HT, name), dtype=np.uint8)
    blocks, udata, off = cpu_blocks(img)
    fasta = H.load_fasta_upper(os.path.join(HT, fa), H.cram_sq_names(blocks, udata, off))
    got = H.cram_decode_records(None, img, blocks, udata, off, fasta, name.encode(), decode_md, _entry=hostsim())
    compare(name, fa, got, decode_md)


@pytest.mark.gpu
@pytest.mark.skipif(ref() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("decode_md", [0, 1])
@pytest.mark.parametrize("name,fa", CASES)
def test_gpu_records_equal_reference(name, fa, decode_md):
    img = np.fromfile(os.path.join(HT, name), dtype=np.uint8)
    ctx = H.Context(0)
    blocks, res = H.cram_uncompress_blocks(ctx, img)
    sizes = blocks["uncomp_size"].astype(np.int64)
    off = np.concatenate([[0], np.cumsum((sizes + 15) // 16 * 16)]).astype(np.uint64)
    udata = np.zeros(int(off[-1]) + 16, dtype=np.uint8)
    for i, (st, data) in enumerate(res):
        assert st == 0, (i, st)
        udata[int(off[i]):int(off[i]) + len(data)] = np.frombuffer(data, dtype=np.uint8)
    fasta = H.load_fasta_upper(os.path.join(HT, fa), H.cram_sq_names(blocks, udata, off))
    got = H.cram_decode_records(ctx, img, blocks, udata, off[:-1].copy(), fasta, name.encode(), decode_md)
    compare(name, fa, got, decode_md)
    ctx.close()


# the SAM files the reference's own test/test.pl round-trips through CRAM (test/*.sam; the part before '#' names the reference),
# written here by the compiled reference in several shapes: CRAM 3.0 / 3.1, no reference, embedded reference, several references
# per slice with tiny slices, generated read names, stored MD/NM, several slices per container
SAMS = sorted(f[:-4] for f in os.listdir(os.path.join(HT, "sam")) if f.endswith(".sam"))
NO_REF, EMBED_REF, MULTI, SEQS, SLICES, LOSSY, STORE_MD, STORE_NM = 11, 7, 10, 3, 4, 19, 21, 22
SHAPES = {"v30": ("3.0", []), "v31": ("3.1", []), "noref": ("3.0", [(NO_REF, 1)]), "embed": ("3.1", [(EMBED_REF, 1)]),
          "multi": ("3.0", [(MULTI, 1), (SEQS, 3)]), "lossy": ("3.1", [(LOSSY, 1)]), "storemd": ("3.0", [(STORE_MD, 1), (STORE_NM, 1)]),
          "slices": ("3.1", [(SLICES, 2), (SEQS, 2)])}


def _written(tmp_path, sam, shape):
    fa = sam.split("#")[0] + ".fa"
    out = str(tmp_path / ("%s.%s.cram" % (sam.replace("#", "_"), shape)))
    version, opts = SHAPES[shape]
    n = ref_write_cram(os.path.join(HT, "sam", sam + ".sam"), os.path.join(HT, fa), out, version, opts)
    return out, fa, n


@pytest.mark.skipif(ref() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", sorted(SHAPES))
@pytest.mark.parametrize("sam", SAMS)
def test_hostsim_written_by_reference(tmp_path, sam, shape):
    out, fa, n = _written(tmp_path, sam, shape)
    img = np.fromfile(out, dtype=np.uint8)
    blocks, udata, off = cpu_blocks(img)
    fasta = H.load_fasta_upper(os.path.join(HT, fa), H.cram_sq_names(blocks, udata, off))
    for decode_md in (0, 1):
        got = H.cram_decode_records(None, img, blocks, udata, off, fasta, os.path.basename(out).encode(), decode_md, _entry=hostsim())
        assert len(got["data"]) == n
        compare(sam, fa, got, decode_md, path=out)


@pytest.mark.gpu
@pytest.mark.skipif(ref() is None, reason="oracle/_ref not built")
@pytest.mark.parametrize("shape", sorted(SHAPES))
def test_gpu_written_by_reference(tmp_path, shape):
    ctx = H.Context(0)
    for sam in SAMS:
        out, fa, n = _written(tmp_path, sam, shape)
        img = np.fromfile(out, dtype=np.uint8)
        blocks, udata, off = cpu_blocks(img)
        fasta = H.load_fasta_upper(os.path.join(HT, fa), H.cram_sq_names(blocks, udata, off))
        for decode_md in (0, 1):
            got = H.cram_decode_records(ctx, img, blocks, udata, off, fasta, os.path.basename(out).encode(), decode_md)
            assert len(got["data"]) == n
            compare(sam, fa, got, decode_md, path=out)
    ctx.close()
