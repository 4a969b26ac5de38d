"""Writes tests/golden/tok3_slices/slice{0..7}.tok3 (+ .names.gz): read-name blocks shaped like the RN
block of a CRAM 3.1 slice — 10 000 Illumina-style names (the generator tools/synth.py uses for the BAM
corpus) in coordinate-sorted, i.e. flowcell-random, order — encoded by the UNMODIFIED reference
(oracle/_ref tok3_encode_names, level 3 = what cram_compress_by_method passes for TOK3,
cram/cram_io.c:1885-1891).  Run once here; the fixtures travel to the GPU box, the reference does not.

    python tests/golden/make_tok3_slices.py
"""
import gzip
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import _libs as L                                                  # noqa: E402

out_dir = os.path.join(HERE, "tok3_slices")
os.makedirs(out_dir, exist_ok=True)
for s in range(8):
    rng = np.random.default_rng(900 + s)
    n = 10000
    lane = rng.integers(1, 5, n)
    tile = rng.choice(np.array([1101 + 100 * a + b for a in range(4) for b in range(1, 79)]), n)
    xs = rng.integers(1000, 33000, n)
    ys = rng.integers(1000, 75000, n)
    names = [b"A00123:45:HXXXXDSXX:%d:%d:%d:%d" % (lane[i], tile[i], xs[i], ys[i]) for i in range(n)]
    if s >= 6:                                                     # name-collated flavour: mates adjacent (DUP tokens)
        names = [nm for nm in names[: n // 2] for _ in (0, 1)]
    blob = b"\0".join(names) + b"\0"
    comp = L.ref_tok3_encode(blob, 3, 0)
    assert L.ref_tok3_decode(comp) == blob
    open(os.path.join(out_dir, "slice%d.tok3" % s), "wb").write(comp)
    with gzip.GzipFile(os.path.join(out_dir, "slice%d.names.gz" % s), "wb", mtime=0) as f:
        f.write(blob)
    print(s, len(blob), len(comp))
