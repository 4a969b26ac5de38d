#!/usr/bin/env python
"""Regenerates tests/golden/htslib/ce#1000.v31.cram / .v30.cram with the compiled reference
(oracle/_ref/libhts_ref.so): ce#1000.sam + ce.fa -> CRAM 3.1 ('normal' profile: RANS_PR blocks incl.
32-way order-1 QS, tok3 names), CRAM 3.0 (rANS 4x8) and a CRAM 3.1 written with use_arith (method 6
blocks, arith-coded tok3 streams, 4 slices).  Run from the repo root in the build container."""
import ctypes as C
import os, sys
import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
G = os.path.join(ROOT, "tests", "golden", "htslib")
r = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhts_ref.so"))
r.hts_open.restype = C.c_void_p; r.hts_open.argtypes = [C.c_char_p, C.c_char_p]
r.sam_hdr_read.restype = C.c_void_p; r.sam_hdr_read.argtypes = [C.c_void_p]
r.sam_hdr_write.argtypes = [C.c_void_p, C.c_void_p]
r.bam_init1.restype = C.c_void_p
r.sam_read1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
r.sam_write1.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
r.hts_close.argtypes = [C.c_void_p]
r.hts_set_opt.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
CRAM_OPT_VERSION, CRAM_OPT_REFERENCE = 6, 9


CRAM_OPT_SEQS_PER_SLICE, CRAM_OPT_USE_FQZ, CRAM_OPT_USE_ARITH, HTS_OPT_COMPRESSION_LEVEL = 3, 25, 26, 100


def write(version, name, reps=1, int_opts=()):
    out = os.path.join(G, name).encode()
    fo = r.hts_open(out, b"wc")
    assert fo
    assert r.hts_set_opt(fo, CRAM_OPT_VERSION, version.encode()) == 0
    r.hts_set_opt.argtypes = [C.c_void_p, C.c_int, C.c_int]
    for opt, val in int_opts:
        assert r.hts_set_opt(fo, opt, val) == 0
    r.hts_set_opt.argtypes = [C.c_void_p, C.c_int, C.c_char_p]
    assert r.hts_set_opt(fo, CRAM_OPT_REFERENCE, os.path.join(G, "ce.fa").encode()) == 0
    first = True
    for _ in range(reps):
        fi = r.hts_open(os.path.join(G, "ce#1000.sam").encode(), b"r")
        h = r.sam_hdr_read(fi)
        if first:
            assert r.sam_hdr_write(fo, h) == 0
            first = False
        b = r.bam_init1()
        n = 0
        while r.sam_read1(fi, h, b) >= 0:
            assert r.sam_write1(fo, h, b) >= 0
            n += 1
        r.hts_close(fi)
    r.hts_close(fo)
    print(name, os.path.getsize(out), "bytes,", n * reps, "records")


if __name__ == "__main__":
    write("3.1", "ce#1000.v31.cram")
    write("3.0", "ce#1000.v30.cram")
    # adaptive-arithmetic blocks (method 6) and arith-coded tok3 streams, several slices; level 3 keeps the
    # bzip2-backed X_EXT methods out (oracle/_ref is built without libbz2)
    write("3.1", "ce#1000.v31fqz.cram", int_opts=[(CRAM_OPT_USE_FQZ, 1), (HTS_OPT_COMPRESSION_LEVEL, 7)])   # fqzcomp quality blocks (method 7)
    write("3.1", "ce#1000.v31arith.cram", int_opts=[(CRAM_OPT_USE_ARITH, 1), (HTS_OPT_COMPRESSION_LEVEL, 3), (CRAM_OPT_SEQS_PER_SLICE, 300)])
