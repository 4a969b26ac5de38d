"""GPU rANS Nx16 encoder (hgpu_rans_nx16_encode_batch_dev): its streams must be decoded to the
input by the unmodified reference decoder, by the oracle and by our own decoder (round trip);
size must stay within a stated ratio of the reference encoder's output."""
import random
import numpy as np
import pytest
import htslib_b200 as H
from _libs import orc_rans_nx16_decode, ref, ref_rans_nx16_decode, ref_rans_nx16_encode
from test_oracle_rans import _synth
from test_gpu_rans import run_batch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = H.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("order", [0, 1, 4, 5])
def test_roundtrip_all_decoders(ctx, order):
    rng = random.Random(40 + order)
    raws = []
    for kind in ("q4", "q40", "runs", "one", "u32", "rand"):
        for n in (0, 1, 3, 31, 32, 33, 100, 1000, 1023, 4099, 70001):
            raws.append(_synth(rng, n, kind) if n else b"")
    comps = ctx.rans_nx16_encode(raws, [order] * len(raws))
    assert all(c is not None for c in comps)
    nz = [(c, r) for c, r in zip(comps, raws) if len(r)]
    for c, r in nz:
        assert orc_rans_nx16_decode(c, len(r)) == r, (order, len(r), c[:8].hex())
        if ref() is not None:
            assert ref_rans_nx16_decode(c, len(r)) == r, (order, len(r), c[:8].hex())
    res = run_batch(ctx, [c for c, _ in nz], [len(r) for _, r in nz])
    for (st, data), (_, r) in zip(res, nz):
        assert st == 0 and data == r


@pytest.mark.skipif(ref() is None, reason="oracle/_ref not built")
def test_size_ratio_vs_reference(ctx):
    """Stated ratio: within 3 % of the reference encoder on quality-like data (>= 64 KiB)."""
    from tools import synth
    rng = np.random.default_rng(5)
    q4 = (synth.novaseq_quals(rng, 1_500_000) + 33).astype(np.uint8).tobytes()
    q40 = (synth.hiseq_quals(rng, 1_500_000) + 33).astype(np.uint8).tobytes()
    for raw in (q4, q40):
        for order in (4, 5, 0, 1):
            mine = ctx.rans_nx16_encode([raw], [order])[0]
            theirs = ref_rans_nx16_encode(raw, order)
            assert ref_rans_nx16_decode(mine, len(raw)) == raw
            assert len(mine) <= 1.03 * len(theirs) + 64, (order, len(mine), len(theirs))
