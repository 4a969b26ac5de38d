"""CPU-side checks of the drop-in boundary: libhtsgpu.so loads without a GPU, exports every
symbol include/htsgpu.h declares, and fails loudly (no CPU fallback) when no device exists."""
import ctypes as C
import os
import re
import numpy as np
import pytest
import htslib_b200 as H


def declared_symbols():
    src = open(H.HEADER_PATH).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b([A-Za-z_][A-Za-z0-9_]*)\s*\(", src)
    skip = {"defined", "sizeof"}
    return sorted({n for n in names if n not in skip and (n.startswith("hgpu_") or n.startswith("rans_") or n.startswith("hts_"))})


def test_library_exports_every_declared_symbol():
    L = H.lib()
    syms = declared_symbols()
    assert len(syms) >= 15
    for s in syms:
        assert hasattr(L, s), "libhtsgpu.so does not export %s" % s


def test_no_cpu_fallback_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(H.HgpuError):
        H.Context(0)
    # reference-named shim must return NULL, not silently decode on the CPU
    comp = open(os.path.join(os.path.dirname(__file__), "golden", "htscodecs", "dat", "r4x16", "q4.1"), "rb").read()
    buf = (C.c_uint8 * len(comp)).from_buffer_copy(comp)
    n = C.c_uint(0)
    assert not H.lib().rans_uncompress_4x16(buf, len(comp), C.byref(n))


def test_host_scan_matches_oracle():
    from _libs import bgzf_file, orc_bgzf_scan
    import random
    rng = random.Random(2)
    data = bytes(rng.randrange(256) for _ in range(200000))
    img = np.frombuffer(bgzf_file(data, 1, block=30000), dtype=np.uint8)
    off, ln, isz = H.bgzf_scan(img)
    n, blocks = orc_bgzf_scan(img.tobytes())
    assert n == len(off) and [(int(a), int(b)) for a, b in zip(off, ln)] == blocks
    assert int(isz.sum()) == len(data)
    bad = img.copy(); bad[int(off[2]) + 1] = 0
    with pytest.raises(H.HgpuError):
        H.bgzf_scan(bad)
    assert H.lib().hgpu_bgzf_scan(bad.ctypes.data, bad.size, None, None, None, 0) == orc_bgzf_scan(bad.tobytes())[0]
