"""ctypes loaders for the three libraries the tests compare.

  orc()  -> oracle/liborc.so          CPU restatement (test infrastructure)
  ref()  -> oracle/_ref/libhts_ref.so unmodified reference, compiled from /root/reference (may be absent)
  The product library is loaded through htslib_b200 itself.
"""
import ctypes as C
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")
u8p = C.POINTER(C.c_uint8)
_orc = _ref = None


def orc():
    global _orc
    if _orc is None:
        so = os.path.join(ROOT, "oracle", "liborc.so")
        srcs = [os.path.join(ROOT, "oracle", f) for f in os.listdir(os.path.join(ROOT, "oracle")) if f.endswith(".c")]
        if not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
            subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liborc.so"])
        _orc = C.CDLL(so)
    return _orc


def ref():
    """The compiled reference, or None when it was never built (needs /root/reference once)."""
    global _ref
    if _ref is None:
        so = os.path.join(ROOT, "oracle", "_ref", "libhts_ref.so")
        if not os.path.exists(so):
            return None
        _ref = C.CDLL(so)
        _ref.rans_uncompress_to_4x16.restype = C.c_void_p
        _ref.rans_compress_to_4x16.restype = C.c_void_p
        _ref.rans_compress_bound_4x16.restype = C.c_uint
    return _ref


def buf(b):
    return (C.c_uint8 * max(1, len(b))).from_buffer_copy(bytes(b) + (b"\0" if len(b) == 0 else b""))


def orc_rans_nx16_decode(data, out_size):
    o = orc()
    out = (C.c_uint8 * max(1, out_size))()
    n = C.c_uint32(out_size)
    rc = o.orc_rans_nx16_decode(buf(data), C.c_uint32(len(data)), out, C.byref(n))
    if rc != 0:
        return None
    return bytes(out[: n.value])


def ref_rans_nx16_decode(data, out_size):
    r = ref()
    out = (C.c_uint8 * max(1, out_size))()
    n = C.c_uint(out_size)
    p = r.rans_uncompress_to_4x16(buf(data), C.c_uint(len(data)), out, C.byref(n))
    if not p:
        return None
    return bytes(out[: n.value])


def ref_rans_nx16_encode(data, order):
    r = ref()
    cap = r.rans_compress_bound_4x16(C.c_uint(len(data)), C.c_int(order))
    out = (C.c_uint8 * max(1, cap))()
    n = C.c_uint(cap)
    p = r.rans_compress_to_4x16(buf(data), C.c_uint(len(data)), out, C.byref(n), C.c_int(order))
    assert p, "reference encoder failed"
    return bytes(out[: n.value])


def golden_raw(name):
    """The raw input the htscodecs tests feed: first column, newlines removed (rans4x16.test:12)."""
    path = os.path.join(GOLD, "htscodecs", "dat", name)
    out = bytearray()
    with open(path, "rb") as f:
        for line in f:
            out += line.rstrip(b"\n").split(b"\t")[0]
    return bytes(out)
