"""BASELINE.json configs[2]: CRAM 3.1 rANS-Nx16 decode of 30x-WGS-shaped slices.

Inputs are written by the UNMODIFIED reference encoder (rans_compress_to_4x16 of oracle/_ref), one
slice = the RANS_PR blocks a 'normal'-profile 10 000 x 150 bp slice holds (SURVEY.md §8a'):
  QS  1 500 000 B  order 5  (X32 | order-1)       NovaSeq 4-bin or HiSeq ~40-value qualities
  BF     15 000 B  order 1  (4-way order-1)       BAM flags, two bytes each
  5 x    10 000 B  order 0  (4-way order-0)       CF / AP / NF / FN / BS-like small series
Every decoded block is compared on the device with the generator's input.  Not the oracle and not
the product: manufactures inputs and times the product's entry points."""
import ctypes as C
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIQ = 16          # unique slices per alphabet, tiled to the batch at distinct addresses


def _ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libhts_ref.so")
    if not os.path.exists(so):
        return None
    r = C.CDLL(so)
    r.rans_compress_to_4x16.restype = C.c_void_p
    r.rans_compress_bound_4x16.restype = C.c_uint
    r.rans_uncompress_to_4x16.restype = C.c_void_p
    r.rans_uncompress_to_4x16.argtypes = [C.c_char_p, C.c_uint, C.c_void_p, C.POINTER(C.c_uint)]
    return r


def ref_encode(r, raw, order):
    cap = r.rans_compress_bound_4x16(C.c_uint(len(raw)), C.c_int(order))
    out = (C.c_uint8 * cap)()
    n = C.c_uint(cap)
    p = r.rans_compress_to_4x16(raw, C.c_uint(len(raw)), out, C.byref(n), C.c_int(order))
    assert p, "reference encoder failed"
    return bytes(out[: n.value])


def make_slices(alphabet, seed=4242):
    """UNIQ slices -> list over stream TYPES (size-descending) of (raws[UNIQ], comps[UNIQ], order)."""
    from tools import synth
    r = _ref()
    assert r is not None, "oracle/_ref/libhts_ref.so is needed to write the input streams"
    rng = np.random.default_rng(seed + (0 if alphabet == "novaseq" else 1))
    qfn = synth.novaseq_quals if alphabet == "novaseq" else synth.hiseq_quals
    types = [("QS", 1_500_000, 5), ("BF", 15_000, 1)] + [("S%d" % k, 10_000, 0) for k in range(5)]
    out = []
    for name, size, order in types:
        raws = []
        for u in range(UNIQ):
            if name == "QS":
                raw = qfn(rng, size).astype(np.uint8).tobytes()           # CRAM stores raw Phred values
            elif name == "BF":
                raw = rng.choice(np.array([99, 147, 83, 163], dtype=np.uint16), size=size // 2).astype("<u2").tobytes()
            else:
                raw = np.clip(rng.normal(60, 25, size=size), 0, 255).astype(np.uint8).tobytes()
            raws.append(raw)
        comps = [ref_encode(r, x, order) for x in raws]
        out.append({"name": name, "order": order, "raws": raws, "comps": comps})
    return out


class Batch:
    """nsl slices laid out type-major (all QS blocks, then all BF blocks, ...): largest first for the
    persistent grids, and the copies of one unique stream sit at a fixed stride so all outputs can be
    compared with the expected bytes in a handful of tensor operations."""

    def __init__(self, types, nsl, torch, dev, only_qs=False):
        assert nsl % UNIQ == 0
        self.torch, self.dev, self.nsl = torch, dev, nsl
        self.types = types[:1] if only_qs else types
        in_len, out_len, src = [], [], []
        for t in self.types:
            for s in range(nsl):
                u = s % UNIQ
                in_len.append(len(t["comps"][u])); out_len.append(len(t["raws"][u])); src.append((t, u))
        self.in_len = np.array(in_len, dtype=np.uint32)
        self.out_len = np.array(out_len, dtype=np.uint32)
        al = lambda a: (a.astype(np.int64) + 15) // 16 * 16
        self.in_off = np.concatenate([[0], np.cumsum(al(self.in_len))[:-1]]).astype(np.uint64)
        self.out_off = np.concatenate([[0], np.cumsum(al(self.out_len))[:-1]]).astype(np.uint64)
        self.n = len(in_len)
        self.U = int(self.out_len.astype(np.int64).sum())
        self.C = int(self.in_len.astype(np.int64).sum())
        self.in_bytes = int(self.in_off[-1]) + int(self.in_len[-1]) + 64
        self.out_bytes = int(self.out_off[-1]) + int(al(self.out_len[-1:])[0]) + 64
        blob = np.zeros(self.in_bytes, dtype=np.uint8)
        # one period (UNIQ streams) per type, then tile
        k = 0
        for t in self.types:
            for s in range(nsl):
                c = t["comps"][s % UNIQ]
                o = int(self.in_off[k]); blob[o:o + len(c)] = np.frombuffer(c, dtype=np.uint8); k += 1
        self.h_in = blob

    def to_device(self):
        torch, dev = self.torch, self.dev
        t = lambda a: torch.from_numpy(a.view(np.int64) if a.dtype == np.uint64 else a.view(np.int32)).to(dev)
        self.d_in = torch.from_numpy(self.h_in).to(dev)
        self.d_out = torch.empty(self.out_bytes, dtype=torch.uint8, device=dev)
        self.d_io, self.d_il, self.d_oo, self.d_ol = t(self.in_off), t(self.in_len), t(self.out_off), t(self.out_len)
        self.d_got = torch.zeros(self.n, dtype=torch.int32, device=dev)
        self.d_st = torch.zeros(self.n, dtype=torch.int32, device=dev)

    def check_device(self, d_out=None):
        """every block == the generator's input"""
        torch = self.torch
        d_out = self.d_out if d_out is None else d_out
        assert int(self.d_st.abs().sum().item()) == 0, "rANS decode reported errors"
        assert bool((self.d_got.cpu().numpy().astype(np.uint32) == self.out_len).all()), "decoded lengths differ"
        k = 0
        for t in self.types:
            L = len(t["raws"][0])
            La = (L + 15) // 16 * 16
            exp = np.zeros((UNIQ, La), dtype=np.uint8)
            for u in range(UNIQ):
                exp[u, :L] = np.frombuffer(t["raws"][u], dtype=np.uint8)
            d_exp = torch.from_numpy(exp).to(self.dev)
            o = int(self.out_off[k])
            got = d_out[o:o + self.nsl * La].view(self.nsl // UNIQ, UNIQ, La)[:, :, :L]
            assert bool(torch.equal(got, d_exp[None, :, :L].expand(self.nsl // UNIQ, UNIQ, L))), "decoded bytes differ from the input (%s)" % t["name"]
            k += self.nsl


def ref_cpu_rate(types, seconds=2.0):
    """rans_uncompress_to_4x16 of the unmodified reference over the slices' blocks: one core, then
    one thread per host core (ctypes releases the GIL)."""
    from concurrent.futures import ThreadPoolExecutor
    r = _ref()
    if r is None:
        return {}
    items = [(c, len(x)) for t in types for c, x in zip(t["comps"], t["raws"])]
    # visit in slice order so every thread sees the real mix of block sizes
    order = [ti * UNIQ + u for u in range(UNIQ) for ti in range(len(types))]
    items = [items[i] for i in order]

    def work(k, until):
        out = (C.c_uint8 * max(n for _, n in items))()
        done = 0
        i = k * 7
        while time.perf_counter() < until:
            c, n = items[i % len(items)]
            m = C.c_uint(n)
            if not r.rans_uncompress_to_4x16(c, len(c), out, C.byref(m)):
                raise RuntimeError("reference decoder failed")
            done += m.value
            i += 1
        return done
    t0 = time.perf_counter()
    one = work(0, t0 + seconds) / (time.perf_counter() - t0)
    cores = len(os.sched_getaffinity(0))
    allc = 0.0
    for _ in range(2):
        t0 = time.perf_counter()
        with ThreadPoolExecutor(cores) as ex:
            tot = sum(ex.map(lambda k: work(k, t0 + seconds), range(cores)))
        allc = max(allc, tot / (time.perf_counter() - t0))
    return {"one_core_GBps": one / 1e9, "all_cores_GBps": allc / 1e9, "cores": cores,
            "sample": "the %d unique slices' blocks in slice order, %.0f s per arm, rans_uncompress_to_4x16 of oracle/_ref (auto SIMD dispatch)" % (UNIQ, seconds)}


def run(ctx, torch, dev, peak, alphabet="novaseq", waves=2, reps=5, world=1, dist=None, e2e=True, cpu=True, only_qs=False):
    import htslib_b200 as H
    types = make_slices(alphabet)
    wave = int(H.lib().hgpu_rans_nx16_wave_size(ctx.h))
    nsl = max(UNIQ, (waves * wave) // UNIQ * UNIQ)
    b = Batch(types, nsl, torch, dev, only_qs=only_qs)
    b.to_device()
    st = torch.cuda.current_stream().cuda_stream
    assert st != 0
    mx = int(b.out_len.max())
    times = []
    L0 = H.lib().hgpu_launch_count()
    for it in range(2 + reps):
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        ctx.rans_nx16_decode_dev(b.d_in, b.d_io, b.d_il, b.d_out, b.d_oo, b.d_ol, b.d_got, b.d_st, mx, st)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tt = torch.tensor([ms], device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            ms = float(tt.item())
        if it >= 2:
            times.append(ms)
    launches = (H.lib().hgpu_launch_count() - L0) // (2 + reps)
    b.check_device()
    ms = float(np.mean(times))
    fmt_hist = {}
    for t in b.types:
        for c in t["comps"]:
            fmt_hist["0x%02x" % c[0]] = fmt_hist.get("0x%02x" % c[0], 0) + 1
    res = {"workload": "CRAM 3.1 rANS-Nx16 decode: %d slices per GPU x (%s), %s qualities, %d unique slices tiled at distinct addresses; streams written by the reference's rans_compress_to_4x16"
                       % (nsl, "QS 1.5MB only" if only_qs else "QS 1.5MB o5 + BF 15kB o1 + 5 x 10kB o0", alphabet, UNIQ),
           "n_gpus": world, "streams_per_gpu": b.n, "format_bytes_of_unique_streams": fmt_hist,
           "resident_32way_streams": wave, "uncompressed_GB_per_gpu": b.U / 1e9, "compressed_GB_per_gpu": b.C / 1e9,
           "ms": ms, "value": world * b.U / ms / 1e6, "unit": "GB/s (uncompressed, all GPUs)", "launches_per_step": int(launches),
           "checked": "every output block equals the generator's input (device compare)",
           "roofline": {"bound": "hbm", "achieved": (b.U + b.C) / ms / 1e6, "peak": peak[0], "unit": "GB/s",
                        "frac": (b.U + b.C) / ms / 1e6 / peak[0], "traffic": None, "peak_source": peak[1],
                        "algorithmic_bytes_per_launch_set": b.U + b.C}}
    if e2e:
        # the same batch through the host-pointer entry point with pinned buffers: H2D of the streams,
        # the five launches, D2H of every decoded byte
        h_in = torch.empty(b.in_bytes, dtype=torch.uint8).pin_memory()
        h_in.copy_(torch.from_numpy(b.h_in))
        h_out = torch.empty(b.out_bytes, dtype=torch.uint8).pin_memory()
        got = np.zeros(b.n, dtype=np.uint32); stt = np.zeros(b.n, dtype=np.int32)
        del b.d_out, b.d_in
        torch.cuda.empty_cache()
        Lh = H.lib()
        fn, on = h_in.numpy(), h_out.numpy()
        wall = []
        for it in range(3):
            if dist is not None:
                dist.barrier()
            t0 = time.perf_counter()
            rc = Lh.hgpu_rans_nx16_decode_batch_host(ctx.h, fn.ctypes.data, b.in_off.ctypes.data, b.in_len.ctypes.data, b.n,
                                                     on.ctypes.data, b.out_off.ctypes.data, b.out_len.ctypes.data,
                                                     got.ctypes.data, stt.ctypes.data)
            sec = time.perf_counter() - t0
            assert rc == 0, H.last_error()
            if dist is not None:
                tt = torch.tensor([sec], device=dev)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                sec = float(tt.item())
            if it >= 1:
                wall.append(sec)
        assert int(np.abs(stt).sum()) == 0 and bool((got == b.out_len).all())
        t0_ = b.types[0]
        L = len(t0_["raws"][0])
        assert on[:L].tobytes() == t0_["raws"][0], "e2e output differs from the input"
        sec = float(np.mean(wall))
        res["e2e"] = {"value": world * b.U / sec / 1e9, "unit": "GB/s", "h2d_bytes_per_step": b.in_bytes + b.n * 24,
                      "d2h_bytes_per_step": int(b.out_off[-1]) + int(b.out_len[-1]) + b.n * 8,
                      "api": "hgpu_rans_nx16_decode_batch_host (pinned host buffers)"}
        del h_in, h_out
    if cpu:
        try:
            res["cpu_baseline"] = dict(ref_cpu_rate(b.types), kind="reference")
        except Exception as ex:
            res["cpu_baseline"] = {"error": repr(ex)}
    return res


if __name__ == "__main__":
    # quick standalone run:  python tools/rans_bench.py [novaseq|hiseq] [waves] [qs]
    import json
    import torch
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import htslib_b200 as H
    torch.cuda.set_device(0)
    dev = torch.device("cuda", 0)
    ctx = H.Context(0)
    s = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(s)
    alphabet = sys.argv[1] if len(sys.argv) > 1 else "novaseq"
    waves = int(sys.argv[2]) if len(sys.argv) > 2 else 1
    only_qs = len(sys.argv) > 3 and sys.argv[3] == "qs"
    pk = (6580.3, "measured")
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        pk = (float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)")
    print(json.dumps(run(ctx, torch, dev, pk, alphabet=alphabet, waves=waves, e2e=False, cpu=False, only_qs=only_qs)))
