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


def test_every_shim_declines_without_a_device():
    """No CPU fallback anywhere on the libhtscodecs seam: with no CUDA device every reference-named entry point
    returns NULL (decoders and encoders alike) instead of computing on the host."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = H.lib()
    for name in ("rans_uncompress", "rans_compress", "arith_compress", "arith_uncompress", "rans_compress_4x16",
                 "tok3_decode_names", "tok3_encode_names", "fqz_decompress", "fqz_compress"):
        getattr(L, name).restype = C.c_void_p
    g = os.path.join(os.path.dirname(__file__), "golden", "htscodecs")
    n = C.c_uint(0); m = C.c_size_t(0); k = C.c_int(0)
    raw = b"ACGT" * 100
    r4x8 = open(os.path.join(g, "dat", "r4x8", "q4.0"), "rb").read()
    arith = open(os.path.join(g, "dat", "arith", "q4.0"), "rb").read()
    tok3 = open(os.path.join(g, "names", "tok3", "01.names.3"), "rb").read()
    fqz = open(os.path.join(g, "dat", "fqzcomp", "q4.0"), "rb").read()
    assert not L.rans_uncompress(r4x8, len(r4x8), C.byref(n))
    assert not L.rans_compress(raw, len(raw), C.byref(n), 0)
    assert not L.arith_uncompress(arith, len(arith), C.byref(n))
    assert not L.arith_compress(raw, len(raw), C.byref(n), 0)
    assert not L.rans_compress_4x16(raw, len(raw), C.byref(n), 0)
    assert not L.tok3_decode_names(tok3, len(tok3), C.byref(n))
    names = C.create_string_buffer(b"read1\0read2\0", 13)
    assert not L.tok3_encode_names(names, 12, 3, 0, C.byref(k), None)
    assert not L.fqz_decompress(fqz, C.c_size_t(len(fqz)), C.byref(m), None, 0)
    lens = (C.c_uint32 * 1)(400); flags = (C.c_uint32 * 1)(0)

    class Slice(C.Structure):
        _fields_ = [("n", C.c_int), ("len", C.POINTER(C.c_uint32)), ("flags", C.POINTER(C.c_uint32))]
    assert not L.fqz_compress(4, C.byref(Slice(1, lens, flags)), raw, C.c_size_t(len(raw)), C.byref(m), 0, None)


def test_header_is_plain_c_and_the_seam_links(tmp_path):
    """include/htsgpu.h must compile as C99, and a C program that names every reference entry point must link
    against libhtsgpu.so the way `-lhtsgpu -lhtscodecs` would (seam B1)."""
    import subprocess
    src = tmp_path / "seam.c"
    src.write_text('''
#include <stdio.h>
#include "htsgpu.h"
int main(void) {
    void *f[] = { (void *)rans_uncompress_to_4x16, (void *)rans_uncompress_4x16, (void *)rans_compress_to_4x16,
                  (void *)rans_compress_4x16, (void *)rans_compress_bound_4x16, (void *)rans_set_cpu,
                  (void *)rans_compress, (void *)rans_uncompress,
                  (void *)arith_compress, (void *)arith_compress_to, (void *)arith_uncompress, (void *)arith_uncompress_to,
                  (void *)arith_compress_bound, (void *)tok3_encode_names, (void *)tok3_decode_names,
                  (void *)fqz_compress, (void *)fqz_decompress, (void *)hts_crc32, (void *)bgzf_compress };
    unsigned n = 0, i;
    for (i = 0; i < sizeof f / sizeof *f; i++) n += f[i] != 0;
    printf("%u entry points, bound(1000,1)=%u\\n", n, arith_compress_bound(1000, 1));
    return 0;
}
''')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    libdir = os.path.join(root, "htslib_b200")
    exe = tmp_path / "seam"
    subprocess.check_call(["gcc", "-std=c99", "-Wall", "-Werror", "-Wno-pedantic", "-I", os.path.join(root, "include"), str(src),
                           "-L", libdir, "-lhtsgpu", "-Wl,-rpath," + libdir, "-Wl,--unresolved-symbols=ignore-in-shared-libs",
                           "-o", str(exe)])
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = "/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.check_output([str(exe)], env=env).decode()
    assert out.startswith("19 entry points") and "bound(1000,1)=" in out
